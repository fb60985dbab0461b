for v in base prefetch base prefetch; do
  echo "== $v"
  UWR_B200_LIB=$PWD/ab/lib_$v.so python tools/kernel_bench.py t5nt t5nn t5h 2>&1 | tail -8
  UWR_B200_LIB=$PWD/ab/lib_$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
