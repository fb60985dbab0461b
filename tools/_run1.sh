set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2a.log
for qk in 1 8 24; do for x in 1 0; do UWR_ATTN_X3=$x python tests/tools/attn_x3_experiment.py 128 $qk 2>&1 | grep -v Warn; done; done > gpurun_out/x3_exp.log 2>&1; cat gpurun_out/x3_exp.log
UWR_ATTN_X3=1 python tools/kernel_bench.py attnb attnf 2>&1 | tail -4
UWR_ATTN_X3=0 python tools/kernel_bench.py attnb attnf 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/prof_r2a_b16.json > gpurun_out/bench_r2a.log 2>gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.log; tail -3 gpurun_out/bench_r2a.err
