run() { echo "== $*"; env "$@" timeout 40 python tools/perf_chol.py hetero --stats --reps 5 2>&1 | tail -3; }
export GSUM_B200_LIB=$PWD/build/lib_ng3pw2.so
timeout 40 python tools/ht_check.py 2 | tail -2 || exit 1
run GSUM_B200_DIAG_DELAY=64
run GSUM_B200_DIAG_DELAY=64 GSUM_B200_FACTOR_CTAS=20
timeout 40 python tools/solve_bench.py 2>&1 | tail -2
