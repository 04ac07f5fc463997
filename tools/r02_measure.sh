#!/bin/bash
# Round-2 measurement pass on ONE B200 (run under gpurun): GPU tests, smoke, both bench arms, the ncu launch list of the
# bench command and one `ncu --set full` capture of each headline kernel.  Everything lands in gpurun_out/r02_*; the
# digests are copied into profiles/ by hand afterwards.   usage: bash tools/r02_measure.sh [tests] [bench] [ncu]
set -u
what="${*:-tests bench ncu}"
O=gpurun_out
mkdir -p $O
if [[ $what == *tests* ]]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02_gpu_tests.txt
  python -c "import __graft_entry__ as g; g.smoke()" >> $O/r02_gpu_tests.txt 2>&1
  tail -3 $O/r02_gpu_tests.txt
fi
if [[ $what == *bench* ]]; then
  python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
  python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err || { tail -5 $O/r02_bench_1gpu.err; exit 1; }
  cat $O/r02_bench_1gpu.json
  python tools/perf_chol.py --nls 16 > $O/r02_perf_chol_16.txt 2>&1
  python tools/perf_chol.py --nls 128 --stats > $O/r02_perf_chol_128.txt 2>&1
  python tools/perf_configs.py > $O/r02_configs.txt 2>&1
fi
if [[ $what == *ncu* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
      python bench.py --steps 2 --warmup 1 > $O/r02_ncu_launches.log 2>&1
  cap() {  # name, kernel regex, launches to skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c 1 -f -o "$O/r02_$name" "$@" > "$O/r02_ncu_$name.log" 2>&1 || echo "capture $name failed"
    ncu -i "$O/r02_$name.ncu-rep" --page raw --csv > "$O/r02_$name.csv" 2>/dev/null || echo "export $name failed"
  }
  cap hetero_tma_128 '^chol_hetero_tma_kernel' 2 python tools/perf_chol.py --nls 128 --reps 1
  cap chain_16       '^chol_hetero_tma_kernel' 2 python tools/perf_chol.py --nls 16 --reps 1
  cap cov_sym        '^cov_sym_kernel'         2 python tools/perf_chol.py --nls 128 --reps 1
  cap smalln         '^smalln_lml_kernel'      1 python tools/smalln_ncu.py
  python tools/ncu_summary.py $O/r02_hetero_tma_128.csv $O/r02_chain_16.csv $O/r02_cov_sym.csv $O/r02_smalln.csv | tee $O/r02_ncu_summary.txt | grep -E "Kernel|duration|DMMA|DRAM read|DRAM write"
fi
