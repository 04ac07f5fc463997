"""CPU: the oracle (oracle/gsum_oracle.py) against every golden vector generated from the real reference and against
the reference's own known answers.  This is what pins the checker the GPU tests rely on."""
import numpy as np
import pytest
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

from oracle import gsum_oracle as o
from util import prior_kwargs, relerr

TOL = 1e-12     # same algorithm, same BLAS: differences are memory-alignment noise of cho_solve (~1e-14)


def test_kat_notebook_grid_argmax(golden):
    """docs/notebooks/correlated_EFT_publication.ipynb cell 58: Best Q 0.4822784810126582, best ls 0.19757575757575757."""
    g = golden("kat_notebook_grid")
    assert tuple(g["argmax"]) == (36, 39)
    assert g["ratio_vals"][36] == 0.4822784810126582 and g["ls_vals"][39] == 0.19757575757575757
    assert g["ll"].max() == pytest.approx(-49.682239225445784, rel=1e-13)
    kern = RBF(0.2) + WhiteKernel(1e-10, 'fixed')
    sub_q, sub_l = [0, 36, 79], [5, 39, 99]
    ll = o.lml_grid(kern, g["X"], g["y"], g["orders"], g["ls_vals"][sub_l], g["ratio_vals"][sub_q], 10.0, o.Priors(0, 0, 1, 1))
    assert relerr(ll, g["ll"][np.ix_(sub_q, sub_l)]) < TOL


@pytest.mark.parametrize("ip", range(4))
@pytest.mark.parametrize("tag", ["g", "t"])
def test_c1_fit_lml_predict(golden, ip, tag):
    g = golden("c1_conjugate")
    p = o.Priors(**prior_kwargs(g["priors"][ip]))
    student = tag == "t"
    kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
    f = o.fit_conjugate(kern, g["X"], g["y"], p, nugget=1e-10, student=student)
    post = np.array([f["center"][0], f["disp"][0, 0], f["df"], f["scale"], f["cov_factor"], f["lml"]])
    want = g[f"{tag}{ip}_post"]
    ok = np.isfinite(want)
    assert np.array_equal(np.isnan(post), np.isnan(want))
    assert relerr(post[ok & np.isfinite(post)], want[ok & np.isfinite(post)]) < TOL
    kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
    lml_fn = o.student_lml if student else o.gaussian_lml
    lml = np.array([lml_fn(kfree, [t], g["X"], g["y"], p, 1e-10) for t in g["thetas"]])
    wl = g[f"{tag}{ip}_lml"]
    assert np.array_equal(np.isnan(lml), np.isnan(wl))
    if np.isfinite(wl).all():
        assert relerr(lml, wl) < TOL
    if f"{tag}{ip}_mean" in g:
        pf = o.predict_student if student else o.predict_conjugate
        m, s = pf(f, g["Xn"], return_std=True)
        assert relerr(m, g[f"{tag}{ip}_mean"]) < TOL and relerr(s, g[f"{tag}{ip}_std"]) < 1e-9
        _, cv = pf(f, g["Xn"][::4], return_cov=True, pred_noise=True)
        assert relerr(cv, g[f"{tag}{ip}_cov"]) < 1e-9
        m, s = pf(f, g["Xn"], return_std=True, Xc=g["Xc"], y=g["yc"])
        assert relerr(m, g[f"{tag}{ip}_mean_c"]) < TOL and relerr(s, g[f"{tag}{ip}_std_c"]) < 1e-9


@pytest.mark.parametrize("tag", ["g", "t"])
def test_c2_truncation_grid(golden, tag):
    g = golden("c2_truncation_grid")
    kern = RBF(0.2) + WhiteKernel(1e-6, 'fixed')
    for ip in range(4):
        p = o.Priors(**prior_kwargs(g["priors"][ip]))
        ll = o.lml_grid(kern, g["X"], g["y"], g["orders"], g["ls_vals"][::3], g["q_vals"][::3], 1.0, p, student=tag == "t")
        want = g[f"{tag}{ip}_ll"][::3, ::3]
        assert np.array_equal(np.isnan(ll), np.isnan(want))
        if np.isfinite(want).all():
            assert relerr(ll, want) < TOL
    p = o.Priors(**prior_kwargs(g["priors"][1]))
    X = g["X"]
    ref = 1.0 + X[:, 0]
    for a, lam in enumerate(g["lams"]):
        q = (0.2 + 0.4 * X[:, 0]) / lam
        for b in (0, 4, 7):
            ll = o.truncation_lml(kern, [np.log(g["ls_vals"][b])], X, g["y2"], g["orders2"], q, ref, p, excluded=[0], student=tag == "t")
            assert ll == pytest.approx(g[f"{tag}_xdep_ll"][a, b], rel=TOL)


@pytest.mark.parametrize("variant", ["const", "xdep"])
def test_c3_truncation_predict(golden, variant):
    g = golden("c3_truncation_predict")
    p = o.Priors(**prior_kwargs(g["prior"]))
    X, Xn, y, orders = g["X"], g["Xn"], g["y"], g["orders"]
    if variant == "const":
        ratio_fn, ref_fn, excl = (lambda X: 0.4 * np.ones(len(X))), (lambda X: np.ones(len(X))), None
    else:
        ratio_fn, ref_fn, excl = (lambda X: 0.3 + 0.2 * X[:, 1]), (lambda X: 1.0 + 0.5 * X[:, 0]), [0]
    mask = ~np.isin(orders, excl)
    coeffs = o.coefficients(y, ratio_fn(X), ref_fn(X), orders)[:, mask]
    kern = RBF([0.05, 0.07], 'fixed') + WhiteKernel(1e-6, 'fixed')
    f = o.fit_conjugate(kern, X, coeffs, p)
    for kind in ("both", "interp", "trunc"):
        m, s = o.predict_truncation(f, Xn, 5, y[:, 5], ratio_fn, ref_fn, return_std=True, kind=kind, excluded=excl)
        assert relerr(m, g[f"g_{variant}_{kind}_mean"]) < 1e-11 and relerr(s, g[f"g_{variant}_{kind}_std"]) < 1e-9
    K = o.truncation_cov(f, Xn[:30], X[:25], ratio_fn, ref_fn, 0, 4, excl)
    assert relerr(K, g[f"g_{variant}_cov_cross"]) < TOL


def test_c5_diagnostics(golden):
    g = golden("c5_diagnostics")
    cov, mean, Y = g["cov"], g["mean"], g["Y"]
    G, piv = o.pivoted_cholesky(cov, return_pivots=True)
    assert np.array_equal(piv, g["piv"]) and np.array_equal(G, g["pchol"])
    ch = np.linalg.cholesky(cov)
    assert np.array_equal(ch, g["chol"])
    assert relerr(o.md_squared(Y, mean, ch), g["md2"]) < TOL
    assert relerr(o.pivoted_cholesky_errors(Y, mean, G), g["pc_errors"]) < TOL
    assert relerr(o.cholesky_errors(Y.T, mean, ch).T, g["chol_errors"]) < TOL
    assert np.array_equal(o.credible_interval(Y, mean, cov, g["intervals"]), g["coverage"])
    assert relerr(o.individual_errors(Y, mean, cov), g["ind_errors"]) < TOL
    # the restated dpstrf picks the same pivots as LAPACK and reproduces its factor
    L, piv2, rank, info = o.dpstrf_restated(cov)
    assert info == 0 and rank == cov.shape[0] and np.array_equal(piv2, g["piv"])
    inv = np.argsort(piv2)
    assert relerr(L[inv], g["pchol"]) < 1e-10
    # property: sum of squared pivoted-Cholesky errors == squared Mahalanobis distance
    assert relerr((g["pc_errors"] ** 2).sum(0), g["md2"]) < 1e-8


def student_diag_covs(g):
    """cov / cov0 of tests/golden/make_golden_student_diag.py rebuilt from the saved (Xd, amp)."""
    from sklearn.gaussian_process.kernels import RBF
    n = g["Xd"].shape[0]
    aa = np.outer(g["amp"], g["amp"])
    return 1.3 * aa * (RBF(0.2)(g["Xd"]) + 1e-5 * np.eye(n)), 0.9 * aa * (RBF(0.25)(g["Xd"]) + 2e-5 * np.eye(n))


def test_c5_student_diagnostics_and_kl(golden):
    """Student-t Diagnostic (gsum/diagnostics.py:51-55) and Diagnostic.kl (116-146) against the reference's outputs."""
    g = golden("c5_student_diag")
    cov, cov0 = student_diag_covs(g)
    mean, df, Y = g["mean"], float(g["df"]), g["Y"]
    assert relerr(o.mvt_draws_from_z(mean, cov, df, g["z"], g["x"]), Y) < 1e-14
    assert np.array_equal(o.credible_interval(Y, mean, cov, g["intervals"], df=df), g["coverage"])
    ch = np.linalg.cholesky(cov)
    assert relerr(o.md_squared(Y, mean, ch), g["md2"]) < TOL
    assert relerr(o.cholesky_errors(Y.T, mean, ch).T, g["chol_errors"]) < TOL
    assert float(o.kl_divergence(mean, cov, ch, g["mean0"], cov0)) == pytest.approx(float(g["kl"]), rel=1e-12)
    assert float(o.kl_divergence(mean, cov, ch, mean, cov)) == pytest.approx(float(g["kl_self"]), rel=1e-12)


def test_kat_pivoted_cholesky(golden):
    """gsum/tests/test.py:75-122 (tabulated, atol 1e-4) and examples/model_checking_tests.ipynb cell 6 (pivots [4 1 3 2])."""
    g = golden("kat_pivoted_cholesky")
    for i in range(3):
        G = o.pivoted_cholesky(g[f"M{i}"])
        np.testing.assert_allclose(g[f"table{i}"], G, atol=1e-4)
        L, piv, rank, info = o.dpstrf_restated(g[f"M{i}"])
        np.testing.assert_allclose(L[np.argsort(piv)], G, atol=1e-12)
    G, piv = o.pivoted_cholesky(g["M_nb"], return_pivots=True)
    assert list(piv + 1) == [4, 1, 3, 2]
    np.testing.assert_allclose(G[0], [0.27726151, 1.66047284, 0, 0], atol=1e-8)
    L, piv2, _, _ = o.dpstrf_restated(g["M_nb"])
    assert np.array_equal(piv2, piv)


def test_dpstrf_restated_rank_deficient():
    """Stop criterion N * DLAMCH('Epsilon') * max diag: same rank and leading pivots as LAPACK on a rank-30 matrix."""
    from scipy.linalg.lapack import dpstrf
    A = np.random.RandomState(0).randn(100, 30)
    M = A @ A.T
    _, p, r, info = dpstrf(M, lower=True)
    L, piv, rank, info2 = o.dpstrf_restated(M)
    assert info == info2 == 1 and rank == r == 30 and np.array_equal(piv[:rank], (p - 1)[:r])


def test_series_helpers_roundtrip():
    rs = np.random.RandomState(0)
    c = rs.randn(20, 5)
    q, ref, orders = 0.2 + 0.5 * rs.rand(20), 1 + rs.rand(20), np.array([0, 2, 3, 4, 6])
    y = o.partials(c, q, ref, orders)
    assert relerr(o.coefficients(y, q, ref, orders), c) < 1e-12
    x = 0.37
    assert o.geometric_sum(x, 2, 5) == pytest.approx(sum(x ** i for i in range(2, 6)))
    assert o.geometric_sum(x, 0, np.inf, excluded=[1]) == pytest.approx(1 / (1 - x) - x)
    with pytest.raises(ValueError):
        o.geometric_sum(x, 3, 2)
    assert o.cartesian(np.arange(2), np.arange(3)).tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2]]


# ---- analytic likelihood gradient (gsum/models.py:957-1056) ------------------------------------------
@pytest.mark.parametrize("ip", range(4))
def test_c1_gradient_oracle(golden, ip):
    """The oracle's restatement of log_marginal_likelihood(theta, eval_gradient=True) against the reference's own output,
    all three hyperparameters of C * RBF + White free, every prior branch."""
    g = golden("c1_gradient")
    p = o.Priors(**prior_kwargs(g["priors"][ip]))
    kern = C(1.5) * RBF(0.2) + WhiteKernel(1e-4)
    for t, lml, grad in zip(g["thetas"], g[f"g{ip}_lml"], g[f"g{ip}_grad"]):
        ll, gr = o.gaussian_lml_gradient(kern, t, g["X"], g["y"], p, 1e-10)
        assert ll == pytest.approx(lml, rel=1e-12)
        assert relerr(gr, grad) < 1e-9
        # and the value agrees with the gradient-free path
        assert ll == pytest.approx(o.gaussian_lml(kern, t, g["X"], g["y"], p, 1e-10), rel=1e-12)


def test_gradient_oracle_aniso_and_finite_difference(golden):
    g = golden("c1_gradient")
    p = o.Priors(**prior_kwargs(g["priors"][2]))
    kern = C(1.2) * RBF([0.3, 0.15]) + WhiteKernel(1e-4)
    for t, lml, grad in zip(g["aniso_thetas"], g["aniso_lml"], g["aniso_grad"]):
        ll, gr = o.gaussian_lml_gradient(kern, t, g["X2"], g["y2"], p, 1e-10)
        assert ll == pytest.approx(lml, rel=1e-12) and relerr(gr, grad) < 1e-9
    # the analytic gradient is the derivative of the likelihood (central differences in log-theta)
    t0 = g["aniso_thetas"][1]
    _, gr = o.gaussian_lml_gradient(kern, t0, g["X2"], g["y2"], p, 1e-10)
    for i in range(len(t0)):
        tp, tm = t0.copy(), t0.copy()
        tp[i] += 1e-4
        tm[i] -= 1e-4
        fd = (o.gaussian_lml(kern, tp, g["X2"], g["y2"], p, 1e-10) - o.gaussian_lml(kern, tm, g["X2"], g["y2"], p, 1e-10)) / 2e-4
        assert gr[i] == pytest.approx(fd, rel=1e-3)      # difference quotient of an ill-conditioned likelihood: ~1e-4 noise


def test_theta_layout_matches_sklearn_order():
    """The host-side map from kernel.theta entries to the device's derivative slots follows sklearn's ordering."""
    from gsum_b200.kernels import theta_layout
    assert theta_layout(C(1.5) * RBF(0.2) + WhiteKernel(1e-4), 1) == [(0, 1.0), (1, 1.0), (2, 1.0)]
    assert theta_layout(C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed'), 1) == [(1, 1.0)]
    assert theta_layout(RBF([0.3, 0.1]) * C(2.0) + WhiteKernel(1e-4), 2) == [(1, 1.0), (2, 1.0), (0, 1.0), (3, 1.0)]
    k = C(1.5) * RBF(0.2) + WhiteKernel(1e-4)
    assert len(theta_layout(k, 1)) == len(k.theta)


@pytest.mark.parametrize("ip", range(3))
def test_student_gradient_oracle_is_the_derivative(golden, ip):
    """The reference's Student-t gradient branch crashes (models.py:1200 passes eval_gradient as Y), so there is no golden
    vector: the oracle's restatement of models.py:1260-1271 is pinned as the derivative of the (golden-pinned) evidence."""
    g = golden("c1_gradient")
    p = o.Priors(**prior_kwargs(g["priors"][ip]))
    kern = C(1.5) * RBF(0.2) + WhiteKernel(1e-4)
    t0 = g["thetas"][1]
    ll, gr = o.student_lml_gradient(kern, t0, g["X"], g["y"], p, 1e-10)
    assert ll == pytest.approx(o.student_lml(kern, t0, g["X"], g["y"], p, 1e-10), rel=1e-12)
    for i in range(len(t0)):
        tp, tm = t0.copy(), t0.copy()
        tp[i] += 1e-4
        tm[i] -= 1e-4
        fd = (o.student_lml(kern, tp, g["X"], g["y"], p, 1e-10) - o.student_lml(kern, tm, g["X"], g["y"], p, 1e-10)) / 2e-4
        assert gr[i] == pytest.approx(fd, rel=1e-3, abs=1e-3)


# ---- decomposition='eig' route: oracle restatement vs the real reference ----
@pytest.mark.parametrize("ip", range(3))
@pytest.mark.parametrize("tag", ["g", "t"])
def test_eig_route_oracle(golden, ip, tag):
    g = golden("eig_route")
    p = o.Priors(**prior_kwargs(g["priors"][ip]))
    student = tag == "t"
    kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
    f = o.fit_conjugate(kern, g["X"], g["y"], p, nugget=1e-10, student=student, decomposition="eig")
    post = np.array([f["center"][0], f["disp"][0, 0], f["df"], f["scale"], f["cov_factor"]])
    want = g[f"{tag}{ip}_post"]
    ok = np.isfinite(want)
    assert np.array_equal(np.isnan(post), np.isnan(want)) and relerr(post[ok], want[ok]) < TOL
    kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
    lml_fn = o.student_lml if student else o.gaussian_lml
    lml = np.array([lml_fn(kfree, [t], g["X"], g["y"], p, 1e-10, decomposition="eig") for t in g["thetas"]])
    wl = g[f"{tag}{ip}_lml"]
    assert np.array_equal(np.isnan(lml), np.isnan(wl))
    if np.isfinite(wl).all():
        assert relerr(lml, wl) < TOL
    if f"{tag}{ip}_mean" in g:
        pf = o.predict_student if student else o.predict_conjugate
        m, s = pf(f, g["Xn"], return_std=True)
        assert relerr(m, g[f"{tag}{ip}_mean"]) < TOL and relerr(s, g[f"{tag}{ip}_std"]) < 1e-9
        _, cv = pf(f, g["Xn"][::4], return_cov=True, pred_noise=True)
        assert relerr(cv, g[f"{tag}{ip}_cov"]) < 1e-9
        m, s = pf(f, g["Xn"], return_std=True, Xc=g["Xc"], y=g["yc"])
        assert relerr(m, g[f"{tag}{ip}_mean_c"]) < TOL and relerr(s, g[f"{tag}{ip}_std_c"]) < 1e-9


def test_eig_route_truncation_and_eigen_errors_oracle(golden):
    g = golden("eig_route")
    kern = RBF(0.2) + WhiteKernel(1e-6, 'fixed')
    n = len(g["Xt"])
    ll = np.array([[o.truncation_lml(kern, [np.log(l)], g["Xt"], g["yt"], g["orders"], q * np.ones(n), np.ones(n),
                                     o.Priors(0, 0, 1, 1), decomposition="eig") for l in g["ls_vals"]] for q in g["q_vals"]])
    assert relerr(ll, g["t_ll"]) < TOL
    cov = 1.3 * np.outer(g["amp"], g["amp"]) * (RBF(0.2)(g["Xd"]) + 1e-5 * np.eye(len(g["Xd"])))
    E = o.eigen_errors(g["Yd"], g["d_mean"], o.eigen_factor(cov))
    assert relerr(E, g["eigen_errors"]) < 1e-9


@pytest.mark.parametrize("ip", range(3))
def test_eig_route_gradient_oracle(golden, ip):
    g = golden("eig_route")
    p = o.Priors(**prior_kwargs(g["priors"][ip]))
    kern = C(1.5) * RBF(0.2) + WhiteKernel(1e-4)
    res = [o.gaussian_lml_gradient(kern, t, g["X"], g["y"], p, 1e-10, decomposition="eig") for t in g["grad_thetas"]]
    assert relerr(np.array([r[0] for r in res]), g[f"g{ip}_glml"]) < TOL
    assert relerr(np.array([r[1] for r in res]), g[f"g{ip}_grad"]) < 1e-9


@pytest.mark.parametrize("case", ["scalar", "xdep", "df0"])
def test_pointwise_oracle(golden, case):
    """the oracle's restatement of TruncationPointwise against the real class' outputs."""
    g = golden("pointwise_variogram")
    p = "pw_" + case + "_"
    ratio, ref = g[p + "ratio"], g[p + "ref"]
    ratio = ratio[0] if ratio.size == 1 else ratio
    ref = ref[0] if ref.size == 1 else ref
    df0, scale0 = g[p + "prior"]
    excluded = g[p + "excluded"].tolist() or None
    f = o.pointwise_fit(g[p + "y"], ratio, ref, g["pw_orders"], df0, scale0, excluded)
    tight = dict(rtol=1e-13, atol=0)
    np.testing.assert_allclose(f["coeffs"], g[p + "coeffs"], **tight)
    assert f["df"] == float(g[p + "df"])
    np.testing.assert_allclose(f["scale"], g[p + "scale"], **tight)
    np.testing.assert_allclose(f["trunc_scale"], g[p + "dist_scale"], **tight)
    np.testing.assert_allclose(o.pointwise_interval(f, g["pw_alpha"]), g[p + "interval"], rtol=1e-12)
    np.testing.assert_allclose(o.pointwise_interval(f, g["pw_alpha"], orders=f["orders_masked"][-2:]), g[p + "interval_sel"], rtol=1e-12)
    np.testing.assert_allclose(o.pointwise_pdf(f, g["pw_ygrid"]), g[p + "pdf"], rtol=1e-12)
    np.testing.assert_allclose(o.pointwise_pdf(f, g["pw_ygrid"], orders=f["orders_masked"][:1], log=True), g[p + "logpdf"], rtol=1e-12)
    np.testing.assert_allclose(f["dist"].std(), g[p + "std"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(o.pointwise_log_likelihood(f), g[p + "ll"], rtol=1e-13)
    np.testing.assert_allclose([o.pointwise_log_likelihood(f, ratio=q) for q in g["pw_ratio_grid"]], g[p + "ll_grid"], rtol=1e-13)
    n = g[p + "y"].shape[0]
    x = np.linspace(0.1, 1.0, n)
    ref_full = np.atleast_1d(ref) * np.ones(n)
    np.testing.assert_allclose([o.pointwise_log_likelihood(f, ratio=q * (0.5 + x), ref=ref_full) for q in g["pw_ratio_grid"]],
                               g[p + "ll_grid_x"], rtol=1e-13)
    assert np.array_equal(o.pointwise_credible_diagnostic(f, g[p + "data"], g["pw_dobs"]), g[p + "dci"])


@pytest.mark.parametrize("case", ["1d", "2d"])
def test_variogram_oracle(golden, case):
    """the oracle's restatement of VariogramFourthRoot against the real class' outputs."""
    g = golden("pointwise_variogram")
    p = "vg_" + case + "_"
    z = g[p + "z"]
    vg = o.VariogramOracle(g[p + "X"], z if z.shape[0] > 1 else z[0], g[p + "bounds"])
    assert np.array_equal(vg.bin_counts, g[p + "bin_counts"]) and np.array_equal(vg.bin_idx, g[p + "bin_idx"])
    tight = dict(rtol=1e-13, atol=0, equal_nan=True)
    np.testing.assert_allclose(vg.bin_locations, g[p + "bin_locations"], **tight)
    np.testing.assert_allclose(vg.gamma_star_hat, g[p + "gamma_star_hat"], **tight)
    np.testing.assert_allclose(vg.gamma_tilde, g[p + "gamma_tilde"], **tight)
    nc = z.shape[0]
    np.testing.assert_allclose(np.array([np.atleast_1d(vg.cov(b)) * np.ones(nc) for b in range(vg.Nb)]), g[p + "cov_diag"], rtol=1e-11, equal_nan=True)
    np.testing.assert_allclose(np.atleast_1d(vg.cov(1, 2)), g[p + "cov_01"], rtol=1e-11)
    for rt in (False, True):
        np.testing.assert_allclose(np.stack(vg.compute(rt_scale=rt)), g[p + f"compute_{int(rt)}"], rtol=1e-11, equal_nan=True)
    i, j, k, l = g[p + "ijkl"].T
    np.testing.assert_allclose(vg.rho_ijkl(i, j, k, l), g[p + "rho"], **tight)
    np.testing.assert_allclose(vg.corr_ijkl(i, j, k, l), g[p + "corr"], rtol=1e-12)
    np.testing.assert_allclose(vg.cov_ijkl(i, j, k, l), g[p + "cov_ijkl"], rtol=1e-12)


# ---- free helpers either side of the path (gsum/helpers.py:202-368, gsum/datasets.py:64-66) --------------------------------
def test_correlation_helpers_oracle(golden):
    g = golden("helpers_datasets")
    for i, ls in enumerate(g["corr_ls"]):
        for tag, X, Xp in (("1d", g["corr_X1"], None), ("3d", g["corr_X2"], None), ("3d_cross", g["corr_X2"], g["corr_Xp2"])):
            assert np.array_equal(o.rbf_corr(X, Xp, ls=ls), g[f"rbf_{tag}_{i}"])
            assert np.array_equal(o.gaussian_corr(X, Xp, ls=ls), g[f"gauss_{tag}_{i}"])
    assert np.array_equal(o.rbf_corr(g["rbf_ls0_X"], ls=0), g["rbf_ls0"])


def test_kl_gauss_oracle(golden):
    g = golden("helpers_datasets")
    assert o.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], cov1=g["kl_cov1"]) == pytest.approx(float(g["kl_from_cov"]), rel=1e-13)
    assert o.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_from_chol"]), rel=1e-13)
    assert o.kl_gauss(0.2, 1.3, -0.4, cov1=0.9) == pytest.approx(float(g["kl_scalar"]), rel=1e-13)
    assert o.kl_gauss(np.zeros(60), g["kl_cov0"], 0.25, chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_scalar_mean1"]), rel=1e-13)
    with pytest.raises(ValueError):
        o.kl_gauss(0.0, 1.0, 0.0)


def test_pdf_summaries_oracle(golden):
    import scipy.stats as st
    g = golden("helpers_datasets")
    x, pdf, alphas = g["pdf_x"], g["pdf_vals"], g["pdf_alphas"]
    assert np.array_equal(np.array([o.hpd_pdf(pdf, a, x) for a in alphas]), g["hpd_pdf"])
    assert o.median_pdf(pdf, x) == float(g["median_pdf"])
    assert np.allclose(np.array([o.hpd(st.norm(0.3, 1.1), a) for a in alphas]), g["hpd_norm"], rtol=0, atol=1e-12)
    assert np.allclose(np.array([o.hpd(st.t(4.5, loc=-1.0, scale=0.6), a) for a in alphas]), g["hpd_t"], rtol=0, atol=1e-12)
    dist = st.norm(np.linspace(-1, 1, 7), np.linspace(0.5, 2.0, 7))
    m, iv = o.predictions(dist, dob=[0.68, 0.95])
    assert np.array_equal(m, g["pred_mean"]) and np.array_equal(iv, g["pred_intervals"])
    assert np.array_equal(o.predictions(dist, dob=0.5)[1], g["pred_interval_single"])


def test_partial_sums_covariance_oracle(golden):
    """The covariance `make_gaussian_partial_sums` draws from (gsum/datasets.py:64-66), and the reference's own 4000 draws
    against it: the sample covariance sits where a Wishart sample of that size must (the same bound the device draws meet)."""
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
    g = golden("helpers_datasets")
    K = o.gaussian_partial_sums_cov(ConstantKernel(1.5) * RBF(0.25) + WhiteKernel(1e-3), g["ds_X"], nugget=1e-4)
    assert np.array_equal(K, g["ds_K"])
    n_draw = 4000
    sd = np.sqrt((K ** 2 + np.outer(np.diag(K), np.diag(K))) / (n_draw - 1))       # std of a sample-covariance entry
    assert np.max(np.abs(g["ds_ref_sample_cov"] - K) / sd) < 5.0
    assert np.max(np.abs(g["ds_ref_sample_mean"] - 0.5) / np.sqrt(np.diag(K) / n_draw)) < 5.0
