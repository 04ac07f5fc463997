#!/bin/bash
# One `ncu --set full` capture per secondary kernel (K5-K9) at the BASELINE.json sizes; raw metric pages land in gpurun_out/side_*.csv, the digest in gpurun_out/side_summary.txt.
# usage: bash tools/ncu_side.sh   (on the GPU box, after `python tools/ncu_targets.py` ran clean without ncu)
set -u
cap() {  # name, target section, kernel regex, launches to skip
  ncu --set full --clock-control none -k "regex:$3" -s "$4" -c 1 -f -o "/tmp/side_$1" \
      python tools/ncu_targets.py "$2" > "gpurun_out/side_$1.log" 2>&1 || echo "capture $1 failed"
  ncu -i "/tmp/side_$1.ncu-rep" --page raw --csv > "gpurun_out/side_$1.csv" 2>/dev/null || echo "export $1 failed"
}
python tools/ncu_targets.py > gpurun_out/side_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/side_plain.log; exit 1; }
cap cov_cross      c3   '^cov_cross_kernel'        0
cap predict_solve  c3   '^chol_hetero_tma_kernel'  1
cap rows_sqnorm    c3   '^rows_sqnorm_kernel'      0
cap schur_cov      c3   '^schur_kernel'            0
cap pstrf_panel    c5   '^pstrf_panel_kernel'      16
cap pstrf_schur    c5   '^schur_kernel'            16
cap chol4096       c5   '^chol_hetero_tma_kernel'  0
cap draws          c5   '^draws_kernel'            1
cap coverage       c5   '^coverage_rows_kernel'    0
cap normal_rows    c5   '^normal_rows_kernel'      0
cap grad_rows      grad '^grad_rows_kernel'        0
cap grad_reduce    grad '^grad_reduce_kernel'      0
python tools/ncu_summary.py gpurun_out/side_*.csv | tee gpurun_out/side_summary.txt
