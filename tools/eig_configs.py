"""Wall-clock of the decomposition='eig' route at the BASELINE.json sizes (C3: 2500 training points in 2-D, 10 000 test points;
C5: eigen_errors at N = 4096), next to the Cholesky route on the same device.  python tools/eig_configs.py"""
import sys
import time

import numpy as np
from sklearn.gaussian_process.kernels import RBF, WhiteKernel

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import gsum_b200 as gb  # noqa: E402
from gsum_b200 import ops  # noqa: E402


def timed(fn, reps=1):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


g1 = np.linspace(0, 1, 50)
X = np.stack(np.meshgrid(g1, g1, indexing="ij"), -1).reshape(-1, 2)
n = len(X)
rs = np.random.RandomState(2)
kern = RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-4, 'fixed')
y = np.linalg.cholesky(RBF([0.02, 0.03])(X) + 1e-8 * np.eye(n)) @ rs.randn(n, 6)
Xt = rs.rand(10000, 2)
pri = dict(center=0, disp=0, df=3, scale=1, nugget=1e-10)
for dec in ("cholesky", "eig"):
    t_fit, gp = timed(lambda: gb.ConjugateGaussianProcess(kern, decomposition=dec, **pri).fit(X, y))
    t_std, _ = timed(lambda: gp.predict(Xt, return_std=True))
    t_lml, _ = timed(lambda: gp.log_marginal_likelihood(theta=gp.kernel_.theta)) if dec == "eig" else (float("nan"), None)
    sweeps = gp._eig.sweeps if dec == "eig" else 0
    print(f"C3 {dec:8s}: fit {1e3 * t_fit:8.1f} ms, predict(std) at 10 000 points {1e3 * t_std:8.1f} ms"
          + (f", eigh sweeps {sweeps}" if dec == "eig" else ""), flush=True)

n = 4096
Xd = np.linspace(0, 1, n)[:, None]
cov = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n))
d = gb.Diagnostic(np.zeros(n), cov, random_state=1)
Y = d.samples(64)
t0 = time.perf_counter(); E = d.eigen_errors(Y); t_first = time.perf_counter() - t0
t_next, _ = timed(lambda: d.eigen_errors(Y), reps=3)
t_pc, _ = timed(lambda: d.pivoted_cholesky_errors(Y), reps=3)
print(f"C5 eigen_errors of 64 curves: first call (Jacobi eigendecomposition of the 4096 x 4096 covariance, {d._eigen.sweeps} sweeps) "
      f"{1e3 * t_first:.0f} ms, later calls {1e3 * t_next:.2f} ms (pivoted_cholesky_errors: {1e3 * t_pc:.2f} ms)", flush=True)
