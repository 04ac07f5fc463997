"""Golden vectors for the decomposition='eig' route and `Diagnostic.eigen_errors`, from the REAL
reference: `ConjugateGaussianProcess` / `ConjugateStudentProcess` / `TruncationGP` constructed with decomposition='eig'
(gsum/models.py:480-484, 713-717, 810-811, 973-974, 1016-1019, 1165-1166, 1215-1216, 1251-1253) and
`Diagnostic.eigen_errors` (gsum/diagnostics.py:63-68, 106-107).  Run by hand in the build container; writes
eig_route.npz next to this file.  numpy 2.3.5 / scipy 1.18.1 / scikit-learn 1.9.0."""
import os
import sys
import warnings

import numpy as np
from scipy import stats
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")

PRIORS = [dict(center=0, disp=0, df=1, scale=1), dict(center=0.3, disp=1, df=3, scale=0.7),
          dict(center=0.1, disp=0, df=np.inf, scale=1.3)]


def prior_array(p):
    return np.array([p["center"], p["disp"], p["df"], p["scale"]], dtype=float)


def main():
    out = {}
    # ---- conjugate processes, N = 51 (odd on purpose: the Jacobi tournament pads to even) ----
    N = 51
    X = np.linspace(0, 1, N)[:, None]
    Xn = np.linspace(0, 1, 120)[:, None]
    K = RBF(0.2)(X) + 1e-6 * np.eye(N)
    y = stats.multivariate_normal(np.zeros(N), K, allow_singular=True).rvs(5, random_state=0).T
    thetas = np.log(np.array([0.05, 0.1, 0.2, 0.3]))
    Xc, yc = X[::3], y[::3]
    out.update(X=X, y=y, Xn=Xn, thetas=thetas, Xc=Xc, yc=yc, priors=np.stack([prior_array(p) for p in PRIORS]),
               noise=np.array(1e-4), nugget=np.array(1e-10), constant=np.array(1.5), ls=np.array(0.2))
    for ip, p in enumerate(PRIORS):
        for tag, cls in (("g", models.ConjugateGaussianProcess), ("t", models.ConjugateStudentProcess)):
            kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
            gp = cls(kern, nugget=1e-10, decomposition='eig', **p).fit(X, y)
            pre = f"{tag}{ip}_"
            out[pre + "post"] = np.array([gp.center_[0], gp.disp_[0, 0], gp.df_, gp.scale_, gp.cov_factor_])
            kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
            gpf = cls(kfree, nugget=1e-10, optimizer=None, decomposition='eig', **p).fit(X, y)
            out[pre + "lml"] = np.array([gpf.log_marginal_likelihood(theta=[t]) for t in thetas])
            if np.isfinite(gp.df_) and gp.df_ > 2:
                m2, s = gp.predict(Xn, return_std=True)
                m3, cv = gp.predict(Xn[::4], return_cov=True, pred_noise=True)
                m4, s4 = gp.predict(Xn, return_std=True, Xc=Xc, y=yc)
                out.update({pre + "mean": m2, pre + "std": s, pre + "cov": cv, pre + "mean_c": m4, pre + "std_c": s4})
            if ip == 0 and tag == "g":
                out["eigvals"] = gp._eigh_tuple_[0]
                out["corr_sqrt"] = gp.corr_sqrt_
    # ---- analytic gradient on the 'eig' route (models.py:1041-1056 through solve_sqrt(..., 'eig')), all hyperparameters free ----
    gthetas = np.log(np.array([[1.5, 0.2, 1e-4], [0.7, 0.1, 1e-3], [2.5, 0.35, 1e-5]]))
    out["grad_thetas"] = gthetas
    for ip, p in enumerate(PRIORS):
        gp = models.ConjugateGaussianProcess(C(1.5) * RBF(0.2) + WhiteKernel(1e-4), nugget=1e-10, optimizer=None,
                                             decomposition='eig', **p).fit(X, y)
        res = [gp.log_marginal_likelihood(theta=t, eval_gradient=True) for t in gthetas]
        out[f"g{ip}_glml"] = np.array([r[0] for r in res])
        out[f"g{ip}_grad"] = np.array([r[1] for r in res])
    # ---- TruncationGP likelihood cells with decomposition='eig', N = 120, orders 0..4 ----
    Nt = 120
    Xt = np.linspace(0, 1, Nt)[:, None]
    Kt = RBF(0.2)(Xt) + 1e-6 * np.eye(Nt)
    coeffs = stats.multivariate_normal(np.zeros(Nt), Kt, allow_singular=True).rvs(5, random_state=1).T
    orders = np.arange(5)
    yt = helpers.partials(coeffs, 0.5, ref=1.0, orders=orders)
    tgp = models.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1,
                              optimizer=None, decomposition='eig').fit(Xt, yt, orders=orders)
    ls_vals, q_vals = np.array([0.08, 0.2, 0.35]), np.array([0.35, 0.5, 0.65])
    out.update(Xt=Xt, yt=yt, orders=orders, ls_vals=ls_vals, q_vals=q_vals,
               t_ll=np.array([[tgp.log_marginal_likelihood(theta=[np.log(l)], ratio=q) for l in ls_vals] for q in q_vals]),
               t_cov_factor=np.array(tgp.coeffs_process.cov_factor_))
    # ---- Diagnostic.eigen_errors, N = 200 ----
    rs = np.random.RandomState(5)
    Nd = 200
    Xd = np.sort(rs.rand(Nd))[:, None]
    amp = 1.0 + 0.5 * rs.rand(Nd)
    cov = 1.3 * np.outer(amp, amp) * (RBF(0.2)(Xd) + 1e-5 * np.eye(Nd))
    mean = 0.2 + 0.1 * Xd[:, 0]
    d = diagnostics.Diagnostic(mean, cov, random_state=3)
    Y = d.samples(12)
    assert Y.shape == (Nd, 12)
    out.update(Xd=Xd, amp=amp, d_mean=mean, Yd=Y, eigen_errors=d.eigen_errors(Y), d_eigvals=np.linalg.eigh(cov)[0])
    path = os.path.join(HERE, "eig_route.npz")
    np.savez_compressed(path, **out)
    print(f"eig_route: {os.path.getsize(path) / 1024:.1f} KiB; keys={len(out)}; t_ll=\n{out['t_ll']}")
    print("g0 post", out["g0_post"], "lml", out["g0_lml"])
    print("g1 grad", out["g1_grad"])


if __name__ == "__main__":
    main()
