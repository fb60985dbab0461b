for v in base stages base stages; do
  echo "== $v"
  export UWR_B200_LIB=$PWD/ab/lib_$v.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  python tools/train_bench.py SpectralTransformer L1withColor 8 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spectral', d['images_per_s_graph'], d['uwr_kernel_ms'].get('uwr_gemm_tcgen05'))"
done
python -m pytest tests -m gpu -x -q -k "gemm or linear" 2>&1 | tail -2
