#!/bin/bash
# SASS census of the shipped library: per kernel, the count of the mnemonics that prove the Blackwell-native path
# (DMMA = FP64 tensor core, UTMALDG = TMA tensor load, SYNCS = mbarrier, LDGSTS = cp.async, UTCxMMA/LDTM/STTM = tcgen05: absent, no f64 kind).
# usage: bash tools/sass_census.sh > profiles/r02_sass_census.txt
so=${1:-gsum_b200/libgsum_b200.so}
echo "# cuobjdump -sass $so  ($(date -u +%F))"
cuobjdump -sass "$so" | awk '
  /Function :/ { k=$3; order[++n]=k }
  /DMMA/ {d[k]++} /UTMALDG/ {t[k]++} /SYNCS/ {s[k]++} /LDGSTS/ {l[k]++} /UTC[A-Z]*MMA|LDTM|STTM/ {u[k]++} /DFMA/ {f[k]++} /MUFU/ {m[k]++}
  END { printf "%-64s %6s %8s %6s %7s %6s %6s %8s\n","kernel","DMMA","UTMALDG","SYNCS","LDGSTS","DFMA","MUFU","tcgen05";
        for(i=1;i<=n;i++){k=order[i]; printf "%-64s %6d %8d %6d %7d %6d %6d %8d\n", substr(k,1,64), d[k],t[k],s[k],l[k],f[k],m[k],u[k]; D+=d[k];T+=t[k];S+=s[k];L+=l[k]}
        printf "%-64s %6d %8d %6d %7d\n","TOTAL",D,T,S,L }'
