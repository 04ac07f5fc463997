"""Device-time of the eig-route GEMMs through gsum_eig_conditional (U = V^T R_on; U^T diag(1/w) U) with device-resident
(w, V): wall clock around the call minus nothing — host arrays for R_on / outputs, so copies are included; the kernel-only
figure comes from ncu (tools/ncu_eig.sh).  Usage: python tools/eig_gemm_probe.py"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from gsum_b200 import ops  # noqa: E402
from gsum_b200._lib import default_context  # noqa: E402

ctx = default_context()
rs = np.random.RandomState(0)
for n, m in ((1024, 4096), (2048, 4096)):
    Q, _ = np.linalg.qr(rs.randn(n, n))
    w = np.linspace(1.0, 2.0, n)
    A = (Q * w[None, :]) @ Q.T
    res = ops.ResidentEigen(0.5 * (A + A.T))
    R_on = rs.randn(n, m)
    Ainv = np.linalg.inv(A)
    lin, var, cov = res.conditional(R_on, rs.randn(n, 4), want_var=True, want_cov=True)
    err = np.max(np.abs(cov - R_on.T @ Ainv @ R_on)) / np.max(np.abs(cov))
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(3):
        res.conditional(R_on, None, want_cov=True)
    dt = (time.perf_counter() - t0) / 3
    flops = 2.0 * n * n * m + 2.0 * n * m * m
    print(f"n={n} m={m}: conditional(cov) {1e3 * dt:.2f} ms wall incl. {8e-6 * (n * m + m * m):.0f} MB of copies, "
          f"{flops / 1e9:.1f} GFLOP, rel err {err:.1e}, launches/call {(ctx.launch_count - l0) // 3}", flush=True)
