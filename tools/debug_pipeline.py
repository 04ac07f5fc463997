"""Dev: compare the pipeline schedule against the dataflow schedule on plain batched factorizations."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsum_b200 import _lib, ops

def run(mode, A):
    os.environ["GSUM_B200_SCHEDULE"] = mode
    ctx = _lib.Context(0)
    L, info, _ = ops.cholesky(A.copy(), return_info=True, ctx=ctx)
    ctx.close()
    return L, info

rs = np.random.RandomState(0)
for n, batch in [(200, 8), (200, 64), (256, 200), (512, 2), (1024, 1), (1024, 2), (1024, 16)]:
    x = np.linspace(0, 1, n)
    A = np.stack([np.exp(-0.5 * ((x[:, None] - x[None, :]) / (0.02 + 0.01 * b)) ** 2) + 1e-6 * np.eye(n) for b in range(batch)])
    Ld, i_d = run("dataflow", A)
    Lp, i_p = run("pipeline", A)
    bad = [b for b in range(batch) if not np.allclose(Ld[b], Lp[b], rtol=1e-9, atol=1e-12, equal_nan=False)]
    print(f"n={n} batch={batch}: info df {np.count_nonzero(i_d)} pl {np.count_nonzero(i_p)}  mismatching matrices {len(bad)} {bad[:10]}")
    if bad:
        b = bad[0]
        T = (n + 63) // 64
        D = np.abs(Ld[b] - Lp[b]); D[np.isnan(D)] = 1e300
        tiles = [(i, k) for i in range(T) for k in range(i + 1) if D[i*64:(i+1)*64, k*64:(k+1)*64].max() > 1e-9]
        print("   first bad matrix", b, "bad tiles (i,k):", tiles[:24])
