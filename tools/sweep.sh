run() { echo "== $*"; env "$@" timeout 60 python tools/perf_chol.py $M --stats --reps 5 2>&1 | tail -3; }
for M in hetero_tma; do export M
HT_MODE=$M timeout 40 python tools/ht_small.py 256 64 2 2>&1 | tail -2
HT_MODE=$M timeout 40 python tools/ht_small.py 64 100 1 2>&1 | tail -1
HT_MODE=$M timeout 40 python tools/ht_check.py 2 2>&1 | tail -2
run GSUM_B200_FACTOR_CTAS=16
run GSUM_B200_FACTOR_CTAS=12
run GSUM_B200_FACTOR_CTAS=14 GSUM_B200_DIAG_DELAY=0
done
