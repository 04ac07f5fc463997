"""Generate the golden input/output vectors from the REAL reference (buqeye/gsum at /root/reference).

Run by hand in the build container (`python tests/golden/make_golden.py`); the .npz files it writes are
committed and are what the test-suite reads (the GPU box has no /root/reference).  Every output below is
produced by the reference's own classes — `ConjugateGaussianProcess`, `ConjugateStudentProcess`,
`TruncationGP`, `TruncationTP`, `Diagnostic`, `pivoted_cholesky` — imported by path with import-only
stubs (see _reference_loader.py).  numpy 2.3.5 / scipy 1.18.1 / scikit-learn 1.9.0.
"""
import os
import sys
import warnings

import numpy as np
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")

PRIORS = [dict(center=0, disp=0, df=1, scale=1), dict(center=0.3, disp=1, df=3, scale=0.7),
          dict(center=-0.2, disp=0.5, df=5, scale=2.0), dict(center=0.1, disp=0, df=np.inf, scale=1.3)]


def prior_array(p):
    return np.array([p["center"], p["disp"], p["df"], p["scale"]], dtype=float)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={list(arrays)}")


# ---------------------------------------------------------------------------------------------------
# 1. Known-answer test of the publication notebook (cells 5, 31-32, 45, 52-53, 58): argmax (36, 39)
# ---------------------------------------------------------------------------------------------------
def kat_notebook():
    x = np.linspace(0, 1, 100)
    X = x[:, None]
    kernel = RBF(0.2, 'fixed') + WhiteKernel(1e-10, 'fixed')
    gp = models.ConjugateGaussianProcess(kernel, center=0, df=np.inf, scale=1, nugget=0)
    coeffs_all = -gp.sample_y(X, n_samples=21, random_state=3)
    data = helpers.partials(coeffs_all, 0.5, ref=10, orders=np.arange(21))[:, :4]
    mask = np.array([(i - 1) % 24 == 0 for i in range(100)])
    orders = np.arange(4)
    tgp = models.TruncationGP(RBF(0.2) + WhiteKernel(1e-10, 'fixed'), ref=10, ratio=0.5, center=0, disp=0, df=1, scale=1,
                              optimizer=None).fit(X[mask], data[mask], orders=orders)
    ls_vals = np.linspace(1e-3, 0.5, 100)
    ratio_vals = np.linspace(0.3, 0.7, 80)
    ll = np.array([[tgp.log_marginal_likelihood(theta=[np.log(ls_)], ratio=q) for ls_ in ls_vals] for q in ratio_vals])
    am = np.unravel_index(np.argmax(ll), ll.shape)
    print("KAT argmax", am, "best Q", ratio_vals[am[0]], "best ls", ls_vals[am[1]], "max", ll.max())
    assert am == (36, 39)
    save("kat_notebook_grid", X=X[mask], y=data[mask], orders=orders, ls_vals=ls_vals, ratio_vals=ratio_vals, ll=ll,
         ref=np.array(10.0), noise=np.array(1e-10), nugget=np.array(1e-10), argmax=np.array(am),
         sqrt_cov_factor=np.sqrt(tgp.coeffs_process.cov_factor_))


# ---------------------------------------------------------------------------------------------------
# 2. C1 — ConjugateGaussianProcess / ConjugateStudentProcess, N = 50, 5 curves: fit, LML, predict
# ---------------------------------------------------------------------------------------------------
def c1_conjugate():
    N = 50
    X = np.linspace(0, 1, N)[:, None]
    Xn = np.linspace(0, 1, 201)[:, None]
    from scipy import stats
    K = RBF(0.2)(X) + 1e-6 * np.eye(N)
    y = stats.multivariate_normal(np.zeros(N), K, allow_singular=True).rvs(5, random_state=0).T
    thetas = np.log(np.array([0.05, 0.1, 0.2, 0.3, 0.45]))
    out = dict(X=X, y=y, Xn=Xn, thetas=thetas, priors=np.stack([prior_array(p) for p in PRIORS]),
               noise=np.array(1e-4), nugget=np.array(1e-10), constant=np.array(1.5), ls=np.array(0.2))
    Xc, yc = X[::3], y[::3]
    out.update(Xc=Xc, yc=yc)
    for ip, p in enumerate(PRIORS):
        for tag, cls in (("g", models.ConjugateGaussianProcess), ("t", models.ConjugateStudentProcess)):
            kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
            gp = cls(kern, nugget=1e-10, **p).fit(X, y)
            pre = f"{tag}{ip}_"
            out[pre + "post"] = np.array([gp.center_[0], gp.disp_[0, 0], gp.df_, gp.scale_, gp.cov_factor_,
                                          gp.log_marginal_likelihood_value_])
            kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
            gpf = cls(kfree, nugget=1e-10, optimizer=None, **p).fit(X, y)
            out[pre + "lml"] = np.array([gpf.log_marginal_likelihood(theta=[t]) for t in thetas])
            if np.isfinite(gp.df_) and ip < 3:
                m = gp.predict(Xn)
                m2, s = gp.predict(Xn, return_std=True)
                m3, cv = gp.predict(Xn[::4], return_cov=True, pred_noise=True)
                m4, s4 = gp.predict(Xn, return_std=True, Xc=Xc, y=yc)
                out.update({pre + "mean": m, pre + "std": s, pre + "cov": cv, pre + "mean_c": m4, pre + "std_c": s4})
            if ip == 0 and tag == "g":
                out["corr_L"] = gp.corr_L_
                out["corr"] = gp.corr_
    save("c1_conjugate", **out)


# ---------------------------------------------------------------------------------------------------
# 3. C2 — TruncationGP / TruncationTP likelihood sub-grids, N = 200, orders 0..5 (+ x-dependent ref / Q variant)
# ---------------------------------------------------------------------------------------------------
def c2_truncation_grid():
    N = 200
    X = np.linspace(0, 1, N)[:, None]
    orders = np.arange(6)
    y = datasets.make_gaussian_partial_sums(X, orders=6, kernel=RBF(0.2) + WhiteKernel(1e-6), ratio=0.5, ref=1., random_state=1)
    ls_vals = np.linspace(0.02, 0.5, 64)[::9]          # 8 of the 64
    q_vals = np.linspace(0.3, 0.7, 64)[::9]
    out = dict(X=X, y=y, orders=orders, ls_vals=ls_vals, q_vals=q_vals, noise=np.array(1e-6), nugget=np.array(1e-10),
               priors=np.stack([prior_array(p) for p in PRIORS]))
    for ip, p in enumerate(PRIORS):
        for tag, cls in (("g", models.TruncationGP), ("t", models.TruncationTP)):
            gp = cls(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, optimizer=None, **p).fit(X, y, orders=orders)
            out[f"{tag}{ip}_ll"] = np.array([[gp.log_marginal_likelihood(theta=[np.log(l)], ratio=q) for l in ls_vals] for q in q_vals])
    # x-dependent ref and ratio, excluded order 0, orders with a gap
    orders2 = np.array([0, 2, 3, 4, 5])
    ref_fn = lambda X: 1.0 + X[:, 0]
    ratio_fn = lambda X, lam=1.0: (0.2 + 0.4 * X[:, 0]) / lam
    y2 = helpers.partials(helpers.coefficients(y, 0.5, 1., orders)[:, orders2], ratio_fn(X), ref_fn(X), orders2)
    lams = np.array([0.8, 1.0, 1.3])
    out.update(orders2=orders2, y2=y2, lams=lams)
    for tag, cls in (("g", models.TruncationGP), ("t", models.TruncationTP)):
        gp = cls(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=ratio_fn, ref=ref_fn, excluded=[0], optimizer=None,
                 **PRIORS[1]).fit(X, y2, orders=orders2)
        out[f"{tag}_xdep_ll"] = np.array([[gp.log_marginal_likelihood(theta=[np.log(l)], lam=lam) for l in ls_vals] for lam in lams])
    save("c2_truncation_grid", **out)


# ---------------------------------------------------------------------------------------------------
# 4. TruncationGP / TruncationTP predict, well-conditioned K_oo (no nugget in the reference's K_oo): 2-D inputs
# ---------------------------------------------------------------------------------------------------
def c3_truncation_predict():
    g = np.linspace(0, 1, 12)
    X = helpers.cartesian(g, g)                      # 144 x 2, spacing 0.09 vs ls 0.05/0.07 -> cond(K_oo) modest
    Xn = np.random.RandomState(2).rand(90, 2)
    orders = np.arange(6)
    kern_true = RBF([0.05, 0.07]) + WhiteKernel(1e-8)
    y = datasets.make_gaussian_partial_sums(X, orders=6, kernel=kern_true, ratio=0.4, ref=1., random_state=2)
    ref_fn = lambda X: 1.0 + 0.5 * X[:, 0]
    ratio_fn = lambda X: 0.3 + 0.2 * X[:, 1]
    out = dict(X=X, Xn=Xn, orders=orders, y=y, ls=np.array([0.05, 0.07]), noise=np.array(1e-6), nugget=np.array(1e-10))
    p = PRIORS[1]
    out["prior"] = prior_array(p)
    for tag, cls in (("g", models.TruncationGP), ("t", models.TruncationTP)):
        for vt, kw in (("const", dict(ratio=0.4, ref=1.0, excluded=None)), ("xdep", dict(ratio=ratio_fn, ref=ref_fn, excluded=[0]))):
            kern = RBF([0.05, 0.07], 'fixed') + WhiteKernel(1e-6, 'fixed')
            gp = cls(kern, optimizer=None, **kw, **p).fit(X, y, orders=orders)
            Koo = gp.cov(X, Xp=X, start=0, end=5)
            pre = f"{tag}_{vt}_"
            out[pre + "cond_Koo"] = np.array(np.linalg.cond(Koo))
            for kind in ("both", "interp", "trunc"):
                if tag == "t" and kind != "both":
                    continue
                m, s = gp.predict(Xn, order=5, return_std=True, kind=kind)
                m2, cv = gp.predict(Xn[:40], order=3, return_cov=True, kind=kind)
                out.update({pre + kind + "_mean": m, pre + kind + "_std": s, pre + kind + "_mean3": m2, pre + kind + "_cov3": cv})
            out[pre + "cp_mean"], out[pre + "cp_std"] = gp.coeffs_process.predict(Xn, return_std=True)
            out[pre + "cov_sym"] = gp.cov(Xn[:30], start=2, end=np.inf)
            out[pre + "cov_cross"] = gp.cov(Xn[:30], Xp=X[:25], start=0, end=4)
            out[pre + "mean_fn"] = gp.mean(Xn, start=1, end=4)
        # constrained truncation error (dX, dy)
        dX = np.array([[0.0, 0.0], [1.0, 1.0]])
        dy = np.array([0.0, 0.0])
        kern = RBF([0.05, 0.07], 'fixed') + WhiteKernel(1e-6, 'fixed')
        gp = cls(kern, optimizer=None, ratio=0.4, ref=1.0, **p).fit(X, y, orders=orders, dX=dX, dy=dy)
        m, s = gp.predict(Xn, order=4, return_std=True, kind='both')
        out.update({f"{tag}_constr_mean": m, f"{tag}_constr_std": s})
    out["dX"], out["dy"] = dX, dy
    save("c3_truncation_predict", **out)


# ---------------------------------------------------------------------------------------------------
# 5. Diagnostics, N = 300 (non-uniform points, heteroscedastic amplitude -> well separated pivots)
# ---------------------------------------------------------------------------------------------------
def c5_diagnostics():
    rs = np.random.RandomState(4)
    N = 300
    Xd = np.sort(rs.rand(N))[:, None]
    amp = 1.0 + 0.5 * rs.rand(N)
    cov = 1.3 * np.outer(amp, amp) * (RBF(0.2)(Xd) + 1e-5 * np.eye(N))
    mean = 0.2 + 0.1 * Xd[:, 0]
    d = diagnostics.Diagnostic(mean, cov, random_state=1)
    Y = d.samples(24)
    iv = np.linspace(0, 1, 101)
    from scipy.linalg.lapack import dpstrf
    c, p, r, info = dpstrf(cov, lower=True)
    save("c5_diagnostics", Xd=Xd, cov=cov, mean=mean, Y=Y, intervals=iv, chol=d._chol, pchol=d._pchol, piv=(p - 1).astype(np.int32),
         rank=np.array(r), md2=d.md_squared(Y), pc_errors=d.pivoted_cholesky_errors(Y), chol_errors=d.cholesky_errors(Y),
         ind_errors=d.individual_errors(Y), coverage=d.credible_interval(Y, iv), md2_1d=np.array(d.md_squared(Y[:, 0])),
         coverage_1d=d.credible_interval(Y[:, 0], iv))


# ---------------------------------------------------------------------------------------------------
# 6. pivoted-Cholesky known answers: gsum/tests/test.py:75-122 and examples/model_checking_tests.ipynb cell 6
# ---------------------------------------------------------------------------------------------------
def kat_pivoted_cholesky():
    Ls = [np.array([[7., 0, 0, 0, 0, 0], [9, 13, 0, 0, 0, 0], [4, 10, 6, 0, 0, 0], [18, 1, 2, 14, 0, 0], [5, 11, 20, 3, 17, 0],
                    [19, 12, 16, 15, 8, 21]]),
          np.array([[1, 0, 0], [2, 3, 0], [4, 5, 6.]]),
          np.array([[6, 0, 0], [3, 2, 0], [4, 1, 5.]])]
    # tabulated answers of the reference's own test (values from TensorFlow-Probability / GPyTorch), atol 1e-4
    pchols = [np.array([[3.4444, -1.3545, 4.084, 1.7674, -1.1789, 3.7562], [8.4685, 1.2821, 3.1179, 12.9197, 0.0000, 0.0000],
                        [7.5621, 4.8603, 0.0634, 7.3942, 4.0637, 0.0000], [15.435, -4.8864, 16.2137, 0.0000, 0.0000, 0.0000],
                        [18.8535, 22.103, 0.0000, 0.0000, 0.0000, 0.0000], [38.6135, 0.0000, 0.0000, 0.0000, 0.0000, 0.0000]]),
              np.array([[0.4558, 0.3252, 0.8285], [2.6211, 2.4759, 0.0000], [8.7750, 0.0000, 0.0000]]),
              np.array([[3.7033, 4.7208, 0.0000], [2.1602, 2.1183, 1.9612], [6.4807, 0.0000, 0.0000]])]
    out = {}
    for i, (L, pc) in enumerate(zip(Ls, pchols)):
        M = L @ L.T
        G = helpers.pivoted_cholesky(M)
        np.testing.assert_allclose(pc, G, atol=1e-4)
        out[f"M{i}"], out[f"table{i}"], out[f"G{i}"] = M, pc, G
    np.random.seed(1)
    r = np.random.rand(4, 4)
    m = r + r.T + 2 * np.eye(4)
    G = helpers.pivoted_cholesky(m)
    print("notebook 4x4 G row0", G[0])
    from scipy.linalg.lapack import dpstrf
    _, p, _, _ = dpstrf(m, lower=True)
    assert list(p) == [4, 1, 3, 2]
    out.update(M_nb=m, G_nb=G, piv_nb=np.array(p) - 1)
    save("kat_pivoted_cholesky", **out)


if __name__ == "__main__":
    kat_notebook()
    c1_conjugate()
    c2_truncation_grid()
    c3_truncation_predict()
    c5_diagnostics()
    kat_pivoted_cholesky()
