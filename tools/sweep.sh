for v in COV_EXP_OFF COV_STORE_OFF; do echo "== $v"; GSUM_B200_LIB=$PWD/build/lib_$v.so timeout 60 python tools/perf_chol.py hetero_tma --reps 6 2>&1 | tail -1; done
echo "== default"; timeout 60 python tools/perf_chol.py hetero_tma --reps 6 2>&1 | tail -1
python - <<'PY'
import torch, time
x = torch.empty(1070*1024*1024//8, dtype=torch.float64, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.fill_(1.5)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print("fill 1.07 GiB: %.3f ms -> %.2f TB/s" % (ms, x.numel()*8/ms*1e-9))
PY
