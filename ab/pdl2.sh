python -m pytest tests/test_ast_gpu.py -x -q -k "programmatic or direct_gradient" 2>&1 | tail -3
for v in base new new1 base new new1; do
  echo "== $v"
  if [ $v = base ]; then export UWR_B200_LIB=$PWD/ab/lib_base.so; unset UWR_PDL; fi
  if [ $v = new ]; then unset UWR_B200_LIB; unset UWR_PDL; fi
  if [ $v = new1 ]; then unset UWR_B200_LIB; export UWR_PDL=1; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
