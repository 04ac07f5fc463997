for v in 4 17 34; do echo "== tiles/CTA $v"; GSUM_B200_LIB=$PWD/build/lib_cov$v.so timeout 60 python tools/perf_chol.py hetero_tma --reps 6 2>&1 | tail -1; done
