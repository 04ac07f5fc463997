#!/bin/bash
# more factor CTAs than 12 (four workers each) at 128 and 64 length scales
O=gpurun_out/r02_sweep6.txt; : > $O
run() { echo "== nls $1 $2" >> $O; env $2 python tools/perf_chol.py --nls $1 --reps 6 2>&1 | grep -E "grid ms" >> $O; }
run 128 GSUM_B200_FACTOR_CTAS=14
run 128 GSUM_B200_FACTOR_CTAS=16
run 64 GSUM_B200_FACTOR_CTAS=16
cat $O
