#!/bin/bash
O=gpurun_out/r02_sweep3.txt; : > $O
run() { echo "== $*" >> $O; env "$@" python tools/perf_chol.py --nls 128 --reps 4 --stats 2>&1 | grep -E "grid ms|sha1|GEMM CTAs" | tail -3 >> $O; }
run GSUM_B200_NGROUPS=3
run GSUM_B200_NGROUPS=2
run GSUM_B200_NGROUPS=1
cat $O
