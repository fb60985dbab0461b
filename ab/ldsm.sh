python -m pytest tests -m gpu -x -q -k "attention or attn or mdassa or newbig" 2>&1 | tail -2
for v in base ldsm base ldsm; do
  echo "== $v"
  export UWR_B200_LIB=$PWD/ab/lib_$v.so
  python tools/kernel_bench.py attnb attnf 2>&1 | grep -v "^$" | tail -6
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
