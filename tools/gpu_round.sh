#!/bin/bash
# One GPU-box pass: parity tests, headline bench (+ reference arm), ncu launch list of the same bench
# command, `--set full` captures of the dominant kernels (one launch each, no source import: the
# whole gpurun_out/ must stay below 64 MiB), and the config-2/3 step tables.  Outputs -> gpurun_out/.
set -x
TAG=${1:-v11}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/prof_r1_b16_$TAG.json > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log
# launch list: the eager (non-graph) step issues the same kernels as the replayed graph
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_launch_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/kernel_bench.py t5nn t5nt t5tn dwf dwb attnf attnb lnb fft > gpurun_out/plain_kb_$TAG.log 2>&1 &&
UWR_KB_REPS=1 ncu --set full --clock-control none -k regex:"gemm_tcgen05|dwconv_.wd|attn_|ln_.wd|fft_pass2" -c 18 \
    -o gpurun_out/ncu_top_$TAG python tools/kernel_bench.py t5nn dwf dwb attnf attnb lnb fft > gpurun_out/ncu_kb_$TAG.log 2>&1
cat gpurun_out/plain_kb_$TAG.log
UWR_PROFILE_OUT=gpurun_out/prof_spectral_$TAG.json python tools/train_bench.py SpectralTransformer L1withColor 8 > gpurun_out/train_spectral_$TAG.log 2>&1; tail -1 gpurun_out/train_spectral_$TAG.log
UWR_PROFILE_OUT=gpurun_out/prof_newbig_$TAG.json python tools/train_bench.py NewBigFRFNModel fflMix 16 > gpurun_out/train_newbig_$TAG.log 2>&1; tail -1 gpurun_out/train_newbig_$TAG.log
python tests/tools/infer_sweep.py > gpurun_out/infer_sweep_$TAG.log 2>&1; tail -3 gpurun_out/infer_sweep_$TAG.log
ls -la gpurun_out | tail -30; du -sh gpurun_out
