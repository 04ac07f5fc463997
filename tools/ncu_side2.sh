set -u
timeout 600 python -m pytest tests/test_gpu_diagnostics.py tests/test_gpu_gradient.py -x -q -m gpu 2>&1 | tail -4
cap() {
  ncu --set full --clock-control none -k "regex:$3" -s "$4" -c 1 -f -o "/tmp/side_$1" python tools/ncu_targets.py "$2" > "gpurun_out/side_$1.log" 2>&1 || echo "capture $1 failed"
  ncu -i "/tmp/side_$1.ncu-rep" --page raw --csv > "gpurun_out/side_$1.csv" 2>/dev/null || echo "export $1 failed"
}
cap coverage       c5   '^coverage_rows_kernel'    0
cap grad_rows      grad '^grad_rows_kernel'        0
cap grad_reduce    grad '^grad_reduce_kernel'      0
python tools/ncu_summary.py gpurun_out/side_coverage.csv gpurun_out/side_grad_rows.csv gpurun_out/side_grad_reduce.csv | grep -v "^$"
timeout 300 python tools/perf_configs.py c5 2>&1 | tail -3
