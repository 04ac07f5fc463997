"""Dev probe: single large factorisations (N = 8192, 12288) through the public ops — residual and timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.gaussian_process.kernels import RBF
from gsum_b200 import ops, _lib
ctx = _lib.default_context()
for n in (8192, 12288):
    X = np.sort(np.random.RandomState(n).rand(n))[:, None]
    A = ops.kernel_matrix(X, None, 0.01, constant=1.0, noise=1e-3)
    ref = RBF(0.01)(X[:200], X[:300]); ref[np.arange(200), np.arange(200)] += 1e-3
    print(n, "kernel_matrix block err", np.max(np.abs(A[:200, :300] - ref)))
    t0 = time.perf_counter(); L = ops.cholesky(A); t1 = time.perf_counter()
    rs = np.random.RandomState(0); v = rs.randn(n, 3)
    res = np.max(np.abs(L @ (L.T @ v) - A @ v)) / np.max(np.abs(A @ v))
    xs = ops.cho_solve(L, A @ v)
    print(f"N={n}: cholesky {1e3 * (t1 - t0):.0f} ms e2e (numpy in/out), |L L^T v - A v| rel {res:.2e}, cho_solve err {np.max(np.abs(xs - v)):.2e}, upper zero {np.all(np.triu(L, 1) == 0)}")
    ctx.profile(True); ops.cholesky(A); ms, fl, nb = ctx.profile_read(); ctx.profile(False)
    print(f"   device factorisation {ms:.2f} ms = {fl / ms * 1e-9:.2f} TFLOP/s")
    if n == 8192:
        G, Lp, piv, rank, status = ops.pivoted_cholesky(A)
        print("   pivoted cholesky status", status, "rank", rank, "G G^T residual", np.max(np.abs((G @ (G.T @ v)) - A @ v)) / np.max(np.abs(A @ v)))
