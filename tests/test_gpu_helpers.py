"""The free helpers and data generators either side of the path on the device: `gaussian`, `rbf`, `kl_gauss` (gsum/helpers.py:
233-261, 310-368) against outputs of the real reference (tests/golden/make_golden_helpers.py) and against the oracle on fresh
inputs; `make_gaussian_partial_sums*` (gsum/datasets.py) exactly against the oracle's factor applied to the same normals, and
statistically against the covariance they must have (the bound the reference's own draws meet, tests/test_oracle.py)."""
import numpy as np
import pytest
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

import gsum_b200 as gb
from oracle import gsum_oracle as o
from util import relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def test_correlation_functions_against_reference_golden(ctx, golden):
    g = golden("helpers_datasets")
    for i, ls in enumerate(g["corr_ls"]):
        for tag, X, Xp in (("1d", g["corr_X1"], None), ("3d", g["corr_X2"], None), ("3d_cross", g["corr_X2"], g["corr_Xp2"])):
            for name, fn in (("rbf", gb.rbf), ("gauss", gb.gaussian)):
                got, want = fn(X, Xp, ls=ls), g[f"{name}_{tag}_{i}"]
                assert got.shape == want.shape
                # entries down to exp(-400): relative agreement, entry by entry
                assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + 1e-300), (name, tag, i)
            if Xp is None:
                assert np.array_equal(np.diag(gb.rbf(X, ls=ls)), np.ones(len(X)))
    assert np.array_equal(gb.rbf(g["rbf_ls0_X"], ls=0), g["rbf_ls0"])


def test_correlation_functions_against_oracle_fresh_inputs(ctx):
    rs = np.random.RandomState(21)
    for n, m, d, ls in ((1, 1, 1, 0.5), (65, 130, 2, 0.3), (257, 40, 4, 1.7)):
        X, Xp = rs.rand(n, d), rs.rand(m, d)
        assert relerr(gb.rbf(X, Xp, ls=ls), o.rbf_corr(X, Xp, ls=ls)) < RTOL
        assert relerr(gb.rbf(X, ls=ls), o.rbf_corr(X, ls=ls)) < RTOL
        assert relerr(gb.gaussian(X, ls=ls), o.gaussian_corr(X, ls=ls)) < RTOL
        assert relerr(gb.gaussian(X, Xp, ls=ls), o.gaussian_corr(X, Xp, ls=ls)) < RTOL     # Xp is not rescaled (reference quirk)


def test_kl_gauss(ctx, golden):
    g = golden("helpers_datasets")
    assert gb.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], cov1=g["kl_cov1"]) == pytest.approx(float(g["kl_from_cov"]), rel=RTOL)
    assert gb.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_from_chol"]), rel=RTOL)
    assert gb.kl_gauss(0.2, 1.3, -0.4, cov1=0.9) == pytest.approx(float(g["kl_scalar"]), rel=RTOL)
    assert gb.kl_gauss(np.zeros(60), g["kl_cov0"], 0.25, chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_scalar_mean1"]), rel=RTOL)
    # fresh inputs against the oracle; KL of a distribution with itself is 0 (up to the 1e-5 I of `stabilize` on the cov1 route)
    rs = np.random.RandomState(3)
    n = 200
    x = np.sort(rs.rand(n))[:, None]
    c1 = RBF(0.1)(x) + 1e-2 * np.eye(n)
    B = rs.randn(n, n) / np.sqrt(n)
    c0 = 0.5 * c1 + 0.1 * B @ B.T
    m0, m1 = rs.randn(n), rs.randn(n)
    assert gb.kl_gauss(m0, c0, m1, cov1=c1) == pytest.approx(o.kl_gauss(m0, c0, m1, cov1=c1), rel=RTOL)
    L1 = np.linalg.cholesky(c1)
    assert abs(gb.kl_gauss(m0, c1, m0, chol1=L1)) < 1e-9
    with pytest.raises(np.linalg.LinAlgError):
        gb.kl_gauss(m0, c0 - 2.0 * np.eye(n), m1, chol1=L1)                               # cov0 not positive definite


def test_partial_sums_equal_the_factor_applied_to_the_same_normals(ctx, golden):
    """Exact part: with the seed's normals z, the generated coefficients are mean + G z for the dpstrf factor G of kernel(X) +
    nugget I, through x-dependent ratio / ref and orders with gaps.  At 30 points the trailing pivots all sit at the noise level
    (1.1e-3) and their order is decided by rounding — any order gives a valid factor, but not the same numbers — so G is the
    device's own `pivoted_cholesky` there (itself checked against dpstrf in tests/test_gpu_diagnostics.py); at 9 scattered
    points every pivot is well separated and G is LAPACK's, through the oracle."""
    g = golden("helpers_datasets")
    kern = ConstantKernel(1.5) * RBF(0.25) + WhiteKernel(1e-3)
    ratio_fn, ref_fn = (lambda X: 0.3 + 0.2 * X[:, 0]), (lambda X: 2.0 - X[:, 0])
    mean_fn = lambda X: 0.5 * np.ones(X.shape[0])
    orders = g["ds_orders"]
    for X, factor in ((g["ds_X"], gb.pivoted_cholesky), (np.sort(np.random.RandomState(8).rand(9))[:, None], o.pivoted_cholesky)):
        y = gb.make_gaussian_partial_sums(X, orders=orders, kernel=kern, mean=mean_fn, ratio=ratio_fn, ref=ref_fn, nugget=1e-4, random_state=5)
        assert y.shape == (len(X), len(orders))
        K = o.gaussian_partial_sums_cov(kern, X, nugget=1e-4)
        G = factor(K)
        assert relerr(G @ G.T, K) < 1e-12
        z = np.random.RandomState(5).standard_normal((len(X), len(orders)))
        want = o.partials(0.5 + G @ z, ratio_fn(X), ref_fn(X), orders)
        assert relerr(y, want) < 1e-10
        assert np.array_equal(y, gb.make_gaussian_partial_sums(X, orders=orders, kernel=kern, mean=mean_fn, ratio=ratio_fn, ref=ref_fn,
                                                               nugget=1e-4, random_state=5))


def test_partial_sums_have_the_covariance_of_the_kernel(ctx, golden):
    g = golden("helpers_datasets")
    kern = ConstantKernel(1.5) * RBF(0.25) + WhiteKernel(1e-3)
    X, K = g["ds_X"], g["ds_K"]
    n_draw = 4000
    big = gb.make_gaussian_partial_sums(X, orders=n_draw, kernel=kern, mean=lambda X: 0.5 * np.ones(len(X)), ratio=1.0, ref=1.0,
                                        nugget=1e-4, random_state=7)
    coeffs = gb.coefficients(big, 1.0, 1.0)
    sd = np.sqrt((K ** 2 + np.outer(np.diag(K), np.diag(K))) / (n_draw - 1))
    assert np.max(np.abs(np.cov(coeffs) - K) / sd) < 5.0
    assert np.max(np.abs(coeffs.mean(axis=1) - 0.5) / np.sqrt(np.diag(K) / n_draw)) < 5.0
    # whitened by the oracle's Cholesky factor the draws are iid N(0, 1): squared Mahalanobis distances ~ chi^2_n
    md2 = o.md_squared(coeffs - 0.5, np.zeros(len(X)), np.linalg.cholesky(K))
    assert abs(np.mean(md2) - len(X)) < 5 * np.sqrt(2 * len(X) / n_draw)


def test_partial_sum_generator_variants_and_errors(ctx, golden):
    g = golden("helpers_datasets")
    Xu, yu = gb.make_gaussian_partial_sums_uniform(n_samples=12, n_features=2, orders=3, random_state=9)
    assert np.array_equal(Xu, g["ds_uniform_X"]) and yu.shape == tuple(g["ds_uniform_y_shape"]) and np.isfinite(yu).all()
    Xg, yg = gb.make_gaussian_partial_sums_on_grid(n_samples=9, n_features=1, orders=4, random_state=9)     # singular RBF(0.5): rank cut
    assert np.array_equal(Xg, g["ds_grid_X"]) and yg.shape == tuple(g["ds_grid_y_shape"]) and np.isfinite(yg).all()
    Xg2, yg2 = gb.make_gaussian_partial_sums_on_grid(n_samples=5, n_features=2, orders=2, nugget=1e-6)
    assert Xg2.shape == (25, 2) and yg2.shape == (25, 2)
    Xs = np.array([[0.0], [0.5], [0.5], [1.0]])
    ys = gb.make_gaussian_partial_sums(Xs, orders=3, kernel=RBF(0.3))
    assert ys.shape == (4, 3) and np.allclose(ys[1], ys[2], rtol=0, atol=1e-7)
    with pytest.raises(np.linalg.LinAlgError):
        gb.make_gaussian_partial_sums(Xs, orders=3, kernel=RBF(0.3), allow_singular=False)
    with pytest.raises(NotImplementedError):
        gb.make_gaussian_partial_sums(Xs, kernel=Matern(0.3))
    # the generated data feed the path: the grid's argmax sits near the truth used to generate them
    X = np.linspace(0, 1, 120)[:, None]
    y = gb.make_gaussian_partial_sums(X, orders=6, kernel=RBF(0.15) + WhiteKernel(1e-6), ratio=0.5, ref=1.0, random_state=2)
    gp = gb.TruncationGP(RBF(0.15) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
    gp.fit(X, y, orders=np.arange(6))
    ls_vals, q_vals = np.linspace(0.05, 0.4, 15), np.linspace(0.3, 0.7, 9)
    ll = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)
    iq, il = np.unravel_index(np.argmax(ll), ll.shape)
    assert abs(q_vals[iq] - 0.5) <= 0.1 and abs(ls_vals[il] - 0.15) <= 0.05


def test_legacy_generators(ctx):
    """`generate_coefficients` / `toy_data` (gsum/helpers.py:36-68): device correlation matrix, factor and draws; normals from
    numpy's global generator as in the reference."""
    X = np.linspace(0, 1, 12)[:, None]
    np.random.seed(4)
    c = gb.generate_coefficients(X, size=3000, beta=0.7, sd=1.5, noise=0.05, ls=0.3)
    assert c.shape == (3000, 12)
    K = 1.5 ** 2 * o.rbf_corr(X, ls=0.3) + 0.05 ** 2 * np.eye(12)
    sd = np.sqrt((K ** 2 + np.outer(np.diag(K), np.diag(K))) / 2999)
    assert np.max(np.abs(np.cov(c.T) - K) / sd) < 5.0 and np.max(np.abs(c.mean(axis=0) - 0.7) / np.sqrt(np.diag(K) / 3000)) < 5.0
    np.random.seed(4)
    assert gb.toy_data(X, orders=np.arange(12), ls=0.3).shape == (12, 12)
