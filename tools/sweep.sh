run() { echo "== $*"; env "$@" timeout 60 python tools/perf_chol.py hetero_tma --stats --reps 5 2>&1 | tail -3; }
HT_MODE=hetero_tma timeout 40 python tools/ht_small.py 256 64 2 2>&1 | tail -2
HT_MODE=hetero_tma timeout 40 python tools/ht_small.py 64 300 1 2>&1 | tail -1
HT_MODE=hetero timeout 40 python tools/ht_small.py 256 64 1 2>&1 | tail -1
HT_MODE=hetero_tma timeout 40 python tools/ht_check.py 2 2>&1 | tail -2
run A=1
run GSUM_B200_FACTOR_CTAS=12
run GSUM_B200_FACTOR_CTAS=10
