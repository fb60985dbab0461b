tools/micro/pdl_gap
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pdl.log 2>&1; tail -3 gpurun_out/pytest_gpu_pdl.log
for v in 0 1 0 1; do
  echo "== UWR_PDL=$v"
  UWR_PDL=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2>gpurun_out/bench_pdl$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
UWR_PDL=0 python tools/train_bench.py SpectralTransformer L1withColor 8 2>&1 | tail -1 | cut -c1-330
UWR_PDL=1 python tools/train_bench.py SpectralTransformer L1withColor 8 2>&1 | tail -1 | cut -c1-330
