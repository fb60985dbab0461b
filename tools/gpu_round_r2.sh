#!/bin/bash
# Round-2 evidence pass on one B200 (outputs -> gpurun_out/, summarised into profiles/ by tools/ncu_summarize.py):
# headline bench (+ both reference arms), ncu launch list of the same step, `--set full` captures of the dominant
# kernels (one launch each), configs 2/3 with their PyTorch-eager legs, inference sweep.
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/prof_b16_$TAG.json > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph > gpurun_out/plain_launch_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/kernel_bench.py t5h t5nn t5nt t5tn t5big dwh attnf attnb lnb mdta oproj conv > gpurun_out/plain_kb_$TAG.log 2>&1 &&
UWR_KB_REPS=1 ncu --set full --clock-control none -k regex:"gemm_tcgen05|dwconv_|attn_|ln_.wd|mdta_|proj_.*mma" -c 26 \
    -o gpurun_out/ncu_top_$TAG python tools/kernel_bench.py t5h t5nn t5nt t5tn dwh attnf attnb lnb mdta oproj > gpurun_out/ncu_kb_$TAG.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw-page CSV of the report (what tools/ncu_summarize.py reads), drop the report
ncu -i gpurun_out/ncu_top_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_top_$TAG.raw.csv 2>/dev/null && rm -f gpurun_out/ncu_top_$TAG.ncu-rep
cat gpurun_out/plain_kb_$TAG.log
UWR_EAGER_REF=1 UWR_PROFILE_OUT=gpurun_out/prof_spectral_$TAG.json python tools/train_bench.py SpectralTransformer L1withColor 8 > gpurun_out/train_spectral_$TAG.log 2>&1; tail -1 gpurun_out/train_spectral_$TAG.log | cut -c1-600
UWR_EAGER_REF=1 UWR_PROFILE_OUT=gpurun_out/prof_newbig_$TAG.json python tools/train_bench.py NewBigFRFNModel fflMix 16 > gpurun_out/train_newbig_$TAG.log 2>&1; tail -1 gpurun_out/train_newbig_$TAG.log | cut -c1-600
# programmatic dependent launch (opt-in): the launch-bound configs with it, and the headline config with it
UWR_PDL=1 python tools/train_bench.py SpectralTransformer L1withColor 8 > gpurun_out/train_spectral_pdl_$TAG.log 2>&1; tail -1 gpurun_out/train_spectral_pdl_$TAG.log | cut -c1-330
UWR_PDL=1 python tools/train_bench.py NewBigFRFNModel fflMix 16 > gpurun_out/train_newbig_pdl_$TAG.log 2>&1; tail -1 gpurun_out/train_newbig_pdl_$TAG.log | cut -c1-330
UWR_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_pdl_$TAG.log 2>/dev/null; cut -c1-200 gpurun_out/bench_pdl_$TAG.log
python tests/tools/infer_sweep.py > gpurun_out/infer_sweep_$TAG.log 2>&1; tail -3 gpurun_out/infer_sweep_$TAG.log
ls -la gpurun_out | tail -20; du -sh gpurun_out
# micro-benchmarks behind DESIGN.md §9: HBM ceiling per read:write mix, graph node cost with / without PDL, L2-sized chunks
for m in rw_mix pdl_gap ffma2_bench; do [ -x tools/micro/$m ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/$m tools/micro/$m.cu; done
(tools/micro/rw_mix; tools/micro/pdl_gap; python tools/micro/l2_chunk_ffn.py) > gpurun_out/micro_$TAG.txt 2>&1; cat gpurun_out/micro_$TAG.txt
tools/micro/ffma2_bench > gpurun_out/ffma2_$TAG.txt 2>&1; cat gpurun_out/ffma2_$TAG.txt | tail -3
