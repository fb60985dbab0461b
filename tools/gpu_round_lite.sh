#!/bin/bash
# Short GPU-box pass: parity tests, headline bench (+ reference arm) and the per-kernel event table.
set -x
TAG=${1:-v14}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/prof_r1_b16_$TAG.json > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_launch_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_$TAG.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
