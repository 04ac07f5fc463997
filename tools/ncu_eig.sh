#!/bin/bash
# `ncu --set full` captures of the 'eig'-route kernels (gsum_eigh's Jacobi round mid-iteration, the DMMA GEMM of the
# solves, the column quadratic form) at N = 1024; raw pages in gpurun_out/eig_*.csv, digest in gpurun_out/eig_summary.txt.
# usage: bash tools/ncu_eig.sh   (on the GPU box)
set -u
cap() {  # name, kernel regex, launches to skip
  ncu --set full --clock-control none --graph-profiling node -k "regex:$2" -s "$3" -c 1 -f -o "/tmp/eig_$1" \
      python tools/ncu_targets.py eig > "gpurun_out/eig_$1.log" 2>&1 || echo "capture $1 failed"
  ncu -i "/tmp/eig_$1.ncu-rep" --page raw --csv > "gpurun_out/eig_$1.csv" 2>/dev/null || echo "export $1 failed"
}
python tools/ncu_targets.py eig > gpurun_out/eig_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/eig_plain.log; exit 1; }
cap jacobi_early  '^jacobi_round'  500
cap jacobi_late   '^jacobi_round'  15000
cap gemm_vt_y     '^eig_gemm_kernel'      0
cap gemm_v_t      '^eig_gemm_kernel'      1
cap gemm_cov      '^eig_gemm_kernel'      5
cap colquad       '^eig_colquad_kernel'   0
python tools/ncu_summary.py gpurun_out/eig_*.csv | tee gpurun_out/eig_summary.txt
