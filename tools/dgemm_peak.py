"""cuBLAS DGEMM throughput (the FP64 roofline denominator; MEASURED_PEAKS.json has none).
Same protocol as MEASURED_PEAKS.json's `how`: torch.matmul f64 n^3, best of 10 (burst) and 4 s back to back (sustained)."""
import json, sys, time, torch
def dgemm_peak(n=8192, sustained_s=4.0):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n**3 / best * 1e-9
    t0 = time.time(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 0; e0.record()
    while time.time() - t0 < sustained_s:
        for _ in range(4): torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sus = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
    return {"n": n, "fp64_tflops": burst, "fp64_tflops_sustained": sus}
if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    print(json.dumps(dgemm_peak(n)))
