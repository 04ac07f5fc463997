"""Dev tool: ONE pass over the secondary kernels of the path (predict, diagnostics, gradient) at the BASELINE.json sizes,
for `ncu -k regex:<kernel> -c 1` captures (tools/ncu_side.sh).  No timing here.

    python tools/ncu_targets.py [c3] [c5] [grad] [eig]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
import gsum_b200 as gb
from oracle import gsum_oracle as o          # input generators only (partials, cartesian)

which = [a for a in sys.argv[1:] if a in ("c3", "c5", "grad", "eig")] or ["c3", "c5", "grad"]

if "c3" in which:
    rs = np.random.RandomState(2)
    g1 = np.linspace(0, 1, 50); X = o.cartesian(g1, g1); n = len(X)
    Xt = rs.rand(10000, 2)
    kern = RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-6, 'fixed')
    coeffs = np.linalg.cholesky(RBF([0.02, 0.03])(X) + 1e-8 * np.eye(n)) @ rs.randn(n, 6)
    orders = np.arange(6); y = o.partials(coeffs, 0.4, 1.0, orders)
    gp = gb.TruncationGP(kern, ratio=0.4, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
    gp.coeffs_process.predict(Xt, return_std=True)
    gp.coeffs_process.predict(Xt[:4096], return_cov=True)

if "c5" in which:
    n = 4096
    Xd = np.linspace(0, 1, n)[:, None]
    cov = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n)); mean = np.zeros(n)
    d = gb.Diagnostic(mean, cov, random_state=1)
    Y = d.samples(64)
    d.md_squared(Y); d.pivoted_cholesky_errors(Y)
    d.sample_coverage(100000, np.linspace(0, 1, 101), counts=True, per_draw=False)

if "grad" in which:
    n = 1024
    X = np.linspace(0, 1, n)[:, None]
    y = np.linalg.cholesky(RBF(0.05)(X) + 1e-6 * np.eye(n)) @ np.random.RandomState(0).randn(n, 6)
    gp = gb.ConjugateGaussianProcess(C(1.0) * RBF(0.05) + WhiteKernel(1e-6), center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y)
    gp.log_marginal_likelihood(gp.kernel_.theta, eval_gradient=True)

if "eig" in which:
    # decomposition='eig' route at the headline N: Jacobi eigendecomposition of R, R^-1 applied to 2048 right-hand sides,
    # the conditioning products of a predict at 4096 test points
    from gsum_b200 import ops
    n = 1024
    X = np.linspace(0, 1, n)[:, None]
    R = RBF(0.05)(X) + 1e-4 * np.eye(n)
    res = ops.ResidentEigen(R)
    rs = np.random.RandomState(0)
    res.solve(rs.randn(n, 2048))
    Xt = rs.rand(4096)[:, None]
    res.conditional(RBF(0.05)(X, Xt), rs.randn(n, 6), want_var=True, want_cov=True)
    print("eig target: sweeps", res.sweeps)
