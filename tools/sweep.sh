run() { echo "== $*"; env "$@" timeout 100 python tools/perf_chol.py hetero --stats --reps 5 2>&1 | tail -3; }
for v in g0f0 g1f1; do
  echo "##### $v"; export GSUM_B200_LIB=$PWD/build/lib_$v.so
  GSUM_B200_DIAG_DELAY=0 timeout 100 python tools/ht_check.py 4 | tail -4
  run GSUM_B200_DIAG_DELAY=0
  run GSUM_B200_DIAG_DELAY=0 GSUM_B200_FACTOR_CTAS=12
done
