for v in base NOSTORE NOSTAGE NOSTORE_NOSTAGE BONCE; do
  echo "== $v"
  UWR_B200_LIB=$PWD/ab/lib_$v.so python tools/kernel_bench.py t5nt t5nn t5h 2>&1 | tail -6
done
