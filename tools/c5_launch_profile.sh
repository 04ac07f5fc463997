ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c5.csv python tools/perf_configs.py c5 > gpurun_out/ncu_c5.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c5.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u=='ns' else (v if u in ('us','usecond') else v*1e3 if u=='ms' else v)
    a=agg.setdefault(r[ki].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{t/1e3:10.2f} ms {n:6d}  {k}")
PY
