"""Dev probe: dependency-free workload for the factorisation kernel's main loop — a forward solve of many right-hand
sides against an existing factor (every border tile row is an independent chain).  Prints the kernel's own cycle split
(GSUM_B200_DF_STATS=1) and the event-timed rate."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GSUM_B200_DF_STATS", "1")
from sklearn.gaussian_process.kernels import RBF
from gsum_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 64
rs = np.random.RandomState(0)
X = np.sort(rs.rand(n))[:, None]
A = RBF(0.01)(X) + 1e-2 * np.eye(n)
L = np.linalg.cholesky(A)
B = rs.randn(n, m)
for rep in range(3):
    t0 = time.perf_counter()
    W = ops.cho_solve(L, B, forward_only=True)
    dt = time.perf_counter() - t0
from scipy.linalg import solve_triangular
ref = solve_triangular(L, B[:, :64], lower=True)
print("n", n, "rhs", m, "wall %.1f ms (host copies included)" % (dt * 1e3), "rel err", np.abs(W[:, :64] - ref).max() / np.abs(ref).max(),
      "| ideal DMMA cycles/CTA %.0f" % ((m / 64) * sum(2 * k for k in range(n // 64)) * 2048 / 148))
