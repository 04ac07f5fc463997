"""Golden vectors for SURVEY.md 8(f).4 from the REAL reference: `TruncationPointwise` (gsum/models.py:1573-1836) and
`VariogramFourthRoot` (gsum/helpers.py:525-730), imported by path with the import-only stubs of _reference_loader.py.

    python tests/golden/make_golden_pointwise.py        (build container only; writes pointwise_variogram.npz)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")
out = {}

# ---- TruncationPointwise ---------------------------------------------------------------------------------------------
rs = np.random.RandomState(7)
n, n_o = 40, 5
orders = np.arange(n_o)
x = np.linspace(0.1, 1.0, n)
coeffs = rs.randn(n, n_o) * 1.3
cases = {
    "scalar": dict(ratio=0.45, ref=1.0, df=1.0, scale=1.0, excluded=None),
    "xdep": dict(ratio=0.2 + 0.4 * x, ref=2.0 + x, df=3.0, scale=0.7, excluded=[1]),
    "df0": dict(ratio=0.3 + 0.2 * x, ref=-1.5 * np.ones(n), df=0.0, scale=1.0, excluded=[0, 3]),
}
out["pw_orders"] = orders
out["pw_alpha"] = alpha = np.array([0.5, 0.68, 0.95])
out["pw_ygrid"] = ygrid = np.linspace(-2.0, 3.0, 7)
out["pw_dobs"] = dobs = np.linspace(0.05, 0.95, 10)
out["pw_ratio_grid"] = ratio_grid = np.linspace(0.2, 0.8, 13)
for name, c in cases.items():
    y = helpers.partials(coeffs, c["ratio"], c["ref"], orders)
    tp = models.TruncationPointwise(df=c["df"], scale=c["scale"], excluded=c["excluded"]).fit(y, c["ratio"], c["ref"], orders)
    p = "pw_" + name + "_"
    out[p + "y"] = y
    out[p + "ratio"], out[p + "ref"] = np.atleast_1d(c["ratio"]).astype(float), np.atleast_1d(c["ref"]).astype(float)
    out[p + "prior"] = np.array([c["df"], c["scale"]])
    out[p + "excluded"] = np.array([] if c["excluded"] is None else c["excluded"], dtype=int)
    out[p + "coeffs"], out[p + "df"], out[p + "scale"] = tp.coeffs_, np.array(tp.df_), tp.scale_
    out[p + "dist_scale"] = np.asarray(tp.dist_.kwds["scale"])
    out[p + "interval"] = tp.interval(alpha)
    out[p + "interval_sel"] = tp.interval(alpha, orders=tp._orders_masked[-2:])
    out[p + "pdf"] = tp.pdf(ygrid)
    out[p + "logpdf"] = tp.logpdf(ygrid, orders=tp._orders_masked[:1])
    out[p + "std"] = tp.std()
    out[p + "ll"] = np.array(tp.log_likelihood())
    # the Lambda_b-style scan: log_likelihood at other ratios (scalar ratio: the reference's broadcast of the Jacobian term)
    out[p + "ll_grid"] = np.array([tp.log_likelihood(ratio=q) for q in ratio_grid])
    out[p + "ll_grid_x"] = np.array([tp.log_likelihood(ratio=q * (0.5 + x), ref=np.atleast_1d(c["ref"]) * np.ones(n)) for q in ratio_grid])
    data = y[:, -1] + 0.3 * rs.randn(n)
    out[p + "data"] = data
    out[p + "dci"] = tp.credible_diagnostic(data, dobs)
    print(name, "df", tp.df_, "ll", out[p + "ll"], "dci shape", out[p + "dci"].shape)

# ---- VariogramFourthRoot ---------------------------------------------------------------------------------------------
for name, (nn, d, ncurves) in {"1d": (24, 1, 1), "2d": (20, 2, 3)}.items():
    rs = np.random.RandomState(11 + d)
    X = rs.rand(nn, d) if d > 1 else np.linspace(0, 1, nn)[:, None]
    from sklearn.gaussian_process.kernels import RBF
    K = RBF(0.3)(X) + 1e-8 * np.eye(nn)
    z = (np.linalg.cholesky(K) @ rs.randn(nn, ncurves)).T            # (ncurves, n) as the class expects
    bounds = np.linspace(0.1, 0.9, 6)
    vg = helpers.VariogramFourthRoot(X, z if ncurves > 1 else z[0], bounds)
    p = "vg_" + name + "_"
    out[p + "X"], out[p + "z"], out[p + "bounds"] = X, z, bounds
    out[p + "bin_counts"], out[p + "bin_locations"] = vg.bin_counts, vg.bin_locations
    out[p + "gamma_star_hat"], out[p + "gamma_tilde"] = vg.gamma_star_hat, vg.gamma_tilde
    out[p + "bin_idx"] = vg.bin_idx
    out[p + "cov_diag"] = np.array([np.atleast_1d(vg.cov(b)) * np.ones(ncurves) for b in range(vg.Nb)])
    out[p + "cov_01"] = np.atleast_1d(vg.cov(1, 2))
    for rt in (False, True):
        g, lo, up = vg.compute(rt_scale=rt)
        out[p + f"compute_{int(rt)}"] = np.stack([g, lo, up])
    idx = np.array([[3, 1, 5, 2], [7, 0, 7, 0], [9, 4, 6, 5]])
    out[p + "ijkl"] = idx
    out[p + "rho"] = vg.rho_ijkl(*idx.T)
    out[p + "corr"] = vg.corr_ijkl(*idx.T)
    out[p + "cov_ijkl"] = vg.cov_ijkl(*idx.T)
    print(name, "counts", vg.bin_counts, "cov diag", out[p + "cov_diag"][:, 0])

path = os.path.join(HERE, "pointwise_variogram.npz")
np.savez_compressed(path, **out)
print(f"pointwise_variogram: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")
