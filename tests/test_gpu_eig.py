"""GPU: the decomposition='eig' route — device Jacobi eigensolver, the solves built on it, the
facade's fit / log_marginal_likelihood / predict with decomposition='eig' and `Diagnostic.eigen_errors` — against
golden vectors from the real reference (tests/golden/make_golden_eig.py), LAPACK and the oracle.

Tolerances.  The reference's 'eig' route forms Q diag(1/eig) Q^T explicitly, so its own results carry an error of
about cond(R) * eps; the goldens were made with noise 1e-4 (cond ~ 1e6) where that floor is ~1e-10.  Eigenvectors are
compared up to sign."""
import numpy as np
import pytest
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, WhiteKernel

import gsum_b200 as gb
from gsum_b200 import ops
from oracle import gsum_oracle as o
from util import prior_kwargs, relerr

pytestmark = pytest.mark.gpu


def _spd(n, seed, ls=0.2, noise=1e-6):
    rs = np.random.RandomState(seed)
    X = np.sort(rs.rand(n))[:, None]
    return RBF(ls)(X) + noise * np.eye(n)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 129, 300])
def test_eigh_against_lapack(ctx, n):
    A = _spd(n, n)
    w, V, sweeps = ops.eigh(A, return_sweeps=True)
    wl, Vl = np.linalg.eigh(A)
    assert np.all(np.diff(w) >= 0) and sweeps <= 25
    # absolute accuracy of LAPACK is eps * |A|; Jacobi is at least that good
    assert np.max(np.abs(w - wl)) < 1e-13 * max(1.0, abs(wl[-1]))
    assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-12
    assert np.max(np.abs(A @ V - V * w[None, :])) < 1e-12 * max(1.0, abs(wl[-1]))
    # sign convention: the largest-magnitude component of every eigenvector is positive
    big = V[np.argmax(np.abs(V), axis=0), np.arange(n)]
    assert np.all(big > 0)


def test_eigh_graded_matrix(ctx):
    """Graded positive-definite matrix D H D, D = diag(1 .. 1e-8): eigenvalues to eps |A| (LAPACK's accuracy class),
    residuals and orthogonality at rounding level."""
    rs = np.random.RandomState(0)
    n = 40
    H = rs.rand(n, n)
    H = H @ H.T / n + np.eye(n)
    d = np.logspace(0, -8, n)
    A = d[:, None] * H * d[None, :]
    w, V = ops.eigh(A)
    wl = np.linalg.eigvalsh(A)
    assert np.max(np.abs(w - wl)) < 1e-14 * wl[-1]
    assert np.max(np.abs(A @ V - V * w[None, :])) < 1e-14 * wl[-1]
    assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-13


def test_eigh_factor_mode_opt_in(ctx, monkeypatch):
    """GSUM_B200_EIGH_FACTOR=1 iterates on the pivoted-Cholesky factor with an eigenvalue-scale rotation criterion and a
    Newton-Schulz polish: fewer sweeps, same spectrum to eps |A|, orthogonal vectors, looser R^-1 (documented in eig.cuh)."""
    A = _spd(300, 11, noise=1e-4)
    w1, V1, s1 = ops.eigh(A, return_sweeps=True)
    monkeypatch.setenv("GSUM_B200_EIGH_FACTOR", "1")
    w2, V2, s2 = ops.eigh(A, return_sweeps=True)
    assert s2 < s1
    assert np.max(np.abs(w1 - w2)) < 1e-13 * w1[-1]
    for V, w in ((V1, w1), (V2, w2)):
        assert np.max(np.abs(V.T @ V - np.eye(300))) < 1e-12
        assert np.max(np.abs(A @ V - V * w[None, :])) < 1e-12 * w[-1]
    Y = np.random.RandomState(0).randn(300, 5)
    Xs = np.linalg.solve(A, Y)
    assert relerr(ops.eig_solve((w1, V1), Y), Xs) < 1e-8 and relerr(ops.eig_solve((w2, V2), Y), Xs) < 1e-5


def test_eigh_without_graph_is_bit_identical(ctx, monkeypatch):
    """The CUDA-graph replay of a sweep and plain stream launches (the fallback for a non-capturable caller stream) run the
    same kernels in the same order: bit-identical output."""
    A = _spd(130, 5)
    w1, V1 = ops.eigh(A)
    monkeypatch.setenv("GSUM_B200_EIGH_NOGRAPH", "1")
    w2, V2 = ops.eigh(A)
    assert np.array_equal(w1, w2) and np.array_equal(V1, V2)
    monkeypatch.setenv("GSUM_B200_EIGH_TWOPASS", "1")
    w3, V3 = ops.eigh(A)
    assert np.max(np.abs(w3 - w1)) < 1e-14 * w1[-1]


def test_eigh_indefinite_and_repeated(ctx):
    rs = np.random.RandomState(3)
    Q, _ = np.linalg.qr(rs.randn(50, 50))
    lam = np.concatenate([[-2.0, -2.0, -0.5], np.zeros(3), np.ones(10), np.linspace(2.5, 5, 34)])
    A = (Q * lam[None, :]) @ Q.T
    A = 0.5 * (A + A.T)
    w, V = ops.eigh(A)
    assert np.max(np.abs(w - np.sort(lam))) < 1e-13 * 5
    assert np.max(np.abs(A @ V - V * w[None, :])) < 1e-13 * 5 * 10
    # non-finite input is reported (numpy raises LinAlgError as well)
    C_ = A.copy(); C_[3, 4] = C_[4, 3] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        ops.eigh(C_)
    # a +lambda / -lambda pair is outside the solver's scope and is reported, not mis-solved
    lam[-1] = 2.0
    B = (Q * lam[None, :]) @ Q.T
    with pytest.raises(np.linalg.LinAlgError):
        ops.eigh(0.5 * (B + B.T))


def test_eig_solve_modes_and_resident(ctx):
    n, k = 150, 37
    A = _spd(n, 7, noise=1e-4)
    rs = np.random.RandomState(1)
    Y, mean = rs.randn(n, k), rs.randn(n)
    wl, Vl = np.linalg.eigh(A)
    X0 = (Vl @ np.diag(1.0 / wl) @ Vl.T) @ Y
    w, V = ops.eigh(A)
    X = ops.eig_solve((w, V), Y)
    # the explicit inverse through an eigendecomposition is accurate to ~cond(A) eps (4e5 * 2e-16) times a modest factor,
    # for LAPACK's vectors and for ours alike: compare both with the backward-stable direct solve
    Xs = np.linalg.solve(A, Y)
    assert relerr(X, Xs) < 10 * relerr(X0, Xs) + 1e-10 and relerr(X, X0) < 1e-8 and relerr(A @ X, Y) < 1e-8
    assert relerr(ops.eig_solve((w, V), Y[:, 0]), X[:, 0]) < 1e-13
    res = ops.ResidentEigen(A)
    assert np.array_equal(res.w, w) and np.array_equal(res.V, V)
    assert np.array_equal(res.solve(Y), X)
    E = res.solve(Y, mean=mean, mode=1)
    # (row-wise comparison with LAPACK's vectors is ill-posed inside the degenerate cluster at the noise level: check the
    # kernel against its own eigenvectors, and the basis-independent sum of squares = (Y - m)^T A^-1 (Y - m))
    assert relerr(E, (V.T @ (Y - mean[:, None])) / np.sqrt(w)[:, None]) < 1e-12
    assert relerr(np.sum(E ** 2, axis=0), np.sum((Y - mean[:, None]) * np.linalg.solve(A, Y - mean[:, None]), axis=0)) < 1e-8
    # conditioning products
    m = 45
    R_on, D = rs.randn(n, m), rs.randn(n, 3)
    lin, var, cov = res.conditional(R_on, D, want_var=True, want_cov=True)
    Ainv = Vl @ np.diag(1.0 / wl) @ Vl.T
    assert relerr(lin, R_on.T @ Ainv @ D) < 1e-8 and relerr(cov, R_on.T @ Ainv @ R_on) < 1e-8
    assert relerr(var, np.diag(R_on.T @ Ainv @ R_on)) < 1e-8 and relerr(var, np.diag(cov)) < 1e-13


@pytest.mark.parametrize("ip", range(3))
@pytest.mark.parametrize("tag", ["g", "t"])
def test_eig_route_fit_lml_predict_golden(ctx, golden, ip, tag):
    g = golden("eig_route")
    cls = gb.ConjugateGaussianProcess if tag == "g" else gb.ConjugateStudentProcess
    pri = prior_kwargs(g["priors"][ip])
    kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
    gp = cls(kern, nugget=1e-10, decomposition='eig', **pri).fit(g["X"], g["y"])
    post = np.array([gp.center_[0], gp.disp_[0, 0], gp.df_, gp.scale_, gp.cov_factor_])
    want = g[f"{tag}{ip}_post"]
    assert np.array_equal(np.isnan(post), np.isnan(want))
    ok = np.isfinite(want) & (want != 0)
    assert np.max(np.abs(post[ok] - want[ok]) / np.abs(want[ok])) < 1e-9
    assert np.all(post[want == 0] == 0)
    kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
    gpf = cls(kfree, nugget=1e-10, optimizer=None, decomposition='eig', **pri).fit(g["X"], g["y"])
    lml = np.array([gpf.log_marginal_likelihood(theta=[t]) for t in g["thetas"]])
    wl = g[f"{tag}{ip}_lml"]
    assert np.array_equal(np.isnan(lml), np.isnan(wl))
    if np.isfinite(wl).all():
        assert relerr(lml, wl) < 1e-9
    if tag == "g":
        # analytic gradient on the 'eig' route against the reference's (models.py:1041-1056 through solve_sqrt(..., 'eig'))
        gpg = cls(C(1.5) * RBF(0.2) + WhiteKernel(1e-4), nugget=1e-10, optimizer=None, decomposition='eig', **pri).fit(g["X"], g["y"])
        res = [gpg.log_marginal_likelihood(theta=t, eval_gradient=True) for t in g["grad_thetas"]]
        assert relerr(np.array([r[0] for r in res]), g[f"g{ip}_glml"]) < 1e-9
        assert relerr(np.array([r[1] for r in res]), g[f"g{ip}_grad"]) < 1e-7
    if ip == 0 and tag == "g":
        w, V = gp._eigh_tuple_
        assert np.max(np.abs(w - g["eigvals"])) < 1e-13 * g["eigvals"][-1]
        assert relerr(gp.corr_sqrt_ @ gp.corr_sqrt_.T, g["corr_sqrt"] @ g["corr_sqrt"].T) < 1e-12
    if f"{tag}{ip}_mean" not in g:
        return
    Xn = g["Xn"]
    m = gp.predict(Xn)
    assert relerr(m, g[f"{tag}{ip}_mean"]) < 1e-9
    m, s = gp.predict(Xn, return_std=True)
    # the posterior covariance is a difference of O(1) terms that cancels to ~1e-5: |error| ~ 1e-12 absolute
    assert relerr(m, g[f"{tag}{ip}_mean"]) < 1e-9 and relerr(s, g[f"{tag}{ip}_std"]) < 1e-6
    m, cv = gp.predict(Xn[::4], return_cov=True, pred_noise=True)
    assert relerr(cv, g[f"{tag}{ip}_cov"]) < 2e-6
    m, s = gp.predict(Xn, return_std=True, Xc=g["Xc"], y=g["yc"])
    assert relerr(m, g[f"{tag}{ip}_mean_c"]) < 1e-9 and relerr(s, g[f"{tag}{ip}_std_c"]) < 1e-6


def test_eig_route_agrees_with_cholesky_route(ctx, golden):
    """Both decompositions of the same well-conditioned R give the same posterior and likelihood."""
    g = golden("eig_route")
    kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
    a = gb.ConjugateGaussianProcess(kern, nugget=1e-10, center=0.3, disp=1, df=3, scale=0.7, decomposition='eig').fit(g["X"], g["y"])
    b = gb.ConjugateGaussianProcess(kern, nugget=1e-10, center=0.3, disp=1, df=3, scale=0.7).fit(g["X"], g["y"])
    for name in ("center_", "disp_", "scale_", "cov_factor_", "log_marginal_likelihood_value_"):
        assert relerr(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))) < 1e-9
    ma, sa = a.predict(g["Xn"], return_std=True)
    mb, sb = b.predict(g["Xn"], return_std=True)
    assert relerr(ma, mb) < 1e-9 and relerr(sa, sb) < 1e-6


def test_eig_route_truncation_lml_golden(ctx, golden):
    g = golden("eig_route")
    tgp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1,
                          optimizer=None, decomposition='eig').fit(g["Xt"], g["yt"], orders=g["orders"])
    assert relerr(np.array(tgp.coeffs_process.cov_factor_), g["t_cov_factor"]) < 1e-7
    ll = np.array([[tgp.log_marginal_likelihood(theta=[np.log(l)], ratio=q) for l in g["ls_vals"]] for q in g["q_vals"]])
    # noise 1e-6: the explicit inverse of the reference's eig route sits at cond * eps ~ 1e-8 of the quadratic forms
    assert relerr(ll, g["t_ll"]) < 1e-7
    # the whole surface in one call (one eigendecomposition per length scale, reused for every Q), scalar and x-dependent Q
    grid = tgp.log_marginal_likelihood_grid(g["ls_vals"], g["q_vals"])
    assert grid.shape == ll.shape and relerr(grid, g["t_ll"]) < 1e-7 and relerr(grid, ll) < 1e-9
    tq = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=lambda X, q=0.5: q * np.ones(len(X)), ref=1, center=0, disp=0,
                         df=1, scale=1, optimizer=None, decomposition='eig').fit(g["Xt"], g["yt"], orders=g["orders"])
    gx = tq.log_marginal_likelihood_grid(g["ls_vals"], ratio_kws_list=[dict(q=q) for q in g["q_vals"]])
    assert relerr(gx, g["t_ll"]) < 1e-7


def test_diagnostic_eigen_errors_golden(ctx, golden):
    g = golden("eig_route")
    n = len(g["Xd"])
    cov = 1.3 * np.outer(g["amp"], g["amp"]) * (RBF(0.2)(g["Xd"]) + 1e-5 * np.eye(n))
    d = gb.Diagnostic(g["d_mean"], cov, random_state=3)
    E = d.eigen_errors(g["Yd"])
    want = g["eigen_errors"]
    assert E.shape == want.shape
    # every row carries the arbitrary sign of its eigenvector.  Row-wise agreement is limited by the eigenvectors'
    # own conditioning: the ~180 eigenvalues near the 1e-5 noise floor are ~5e-8 apart, so eps |cov| / gap ~ 1e-6
    sgn = np.sign(np.sum(E * want, axis=1))
    assert relerr(E * sgn[:, None], want) < 2e-5
    # sign-free invariants: the squared errors sum to the squared Mahalanobis distance
    assert relerr(np.sum(E ** 2, axis=0), d.md_squared(g["Yd"])) < 1e-8
    assert relerr(d.eigen_errors(g["Yd"][:, 0]), E[:, 0]) == 0
    assert relerr(d._eig @ d._eig.T, cov) < 1e-12
    assert np.max(np.abs(np.sort(d._eigen.w) - g["d_eigvals"])) < 1e-13 * g["d_eigvals"][-1]


def test_eigh_larger_matrix_properties(ctx):
    """N = 1024 (the headline size): residual, orthogonality and R^-1 through the eigendecomposition."""
    n = 1024
    X = np.linspace(0, 1, n)[:, None]
    A = RBF(0.05)(X) + 1e-4 * np.eye(n)
    res = ops.ResidentEigen(A)
    w, V = res.w, res.V
    assert res.sweeps <= 25 and np.all(np.diff(w) >= 0) and w[0] > 0.9e-4
    assert np.max(np.abs(V.T @ V - np.eye(n))) < 1e-11
    assert np.max(np.abs(A @ V - V * w[None, :])) < 1e-11 * w[-1]
    rs = np.random.RandomState(0)
    Y = rs.randn(n, 7)
    assert relerr(A @ res.solve(Y), Y) < 1e-8
    assert abs(np.sum(np.log(w)) - np.linalg.slogdet(A)[1]) < 1e-8 * abs(np.linalg.slogdet(A)[1])


def test_c3_size_eig_route_agrees_with_cholesky_route(ctx):
    """BASELINE config 3 size (2-D inputs, 50 x 50 = 2500 training points, anisotropic RBF): the 'eig' route's posterior,
    likelihood and predictive mean / std against the Cholesky route on the same device (size-independent property:
    both decompositions represent the same R)."""
    g1 = np.linspace(0, 1, 50)
    X = o.cartesian(g1, g1)
    n = len(X)
    rs = np.random.RandomState(2)
    kern = RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-4, 'fixed')
    y = np.linalg.cholesky(RBF([0.02, 0.03])(X) + 1e-8 * np.eye(n)) @ rs.randn(n, 3)
    Xt = rs.rand(1500, 2)
    pri = dict(center=0.1, disp=2.0, df=3, scale=0.8, nugget=1e-10)
    a = gb.ConjugateGaussianProcess(kern, decomposition='eig', **pri).fit(X, y)
    b = gb.ConjugateGaussianProcess(kern, **pri).fit(X, y)
    for name in ("center_", "disp_", "scale_", "cov_factor_", "log_marginal_likelihood_value_"):
        assert relerr(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))) < 1e-9, name
    ma, sa = a.predict(Xt, return_std=True)
    mb, sb = b.predict(Xt, return_std=True)
    assert relerr(ma, mb) < 1e-8 and relerr(sa, sb) < 1e-6
    w, V = a._eigh_tuple_
    assert w[0] > 0.9e-4 and abs(np.sum(np.log(w)) - 2 * np.sum(np.log(np.diag(b.corr_L_)))) < 1e-8 * n


def test_c5_size_eigen_errors_properties(ctx):
    """BASELINE config 5 size (N = 4096): eigen_errors of 16 held-out curves; the sign-free invariant
    sum_k e_k^2 = squared Mahalanobis distance ties the Jacobi eigendecomposition to the Cholesky factor."""
    n = 4096
    Xd = np.linspace(0, 1, n)[:, None]
    cov = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n))
    d = gb.Diagnostic(np.zeros(n), cov, random_state=1)
    Y = d.samples(16)
    E = d.eigen_errors(Y)
    assert E.shape == (n, 16) and np.all(np.isfinite(E))
    assert relerr(np.sum(E ** 2, axis=0), d.md_squared(Y)) < 1e-7
    w = d._eigen.w
    assert np.all(np.diff(w) >= 0) and abs(w.sum() - np.trace(cov)) < 1e-10 * np.trace(cov)
    # draws from N(0, cov): the errors are standard normal whatever the basis
    assert abs(np.mean(E)) < 0.02 and abs(np.std(E) - 1.0) < 0.02


def test_eig_route_fit_with_free_length_scale(ctx):
    """`fit` with the default optimizer on the 'eig' route (L-BFGS on the analytic gradient, gsum/models.py:630-669):
    same optimum as the Cholesky route."""
    rs = np.random.RandomState(4)
    X = np.linspace(0, 1, 80)[:, None]
    y = np.linalg.cholesky(RBF(0.17)(X) + 1e-6 * np.eye(80)) @ rs.randn(80, 4)
    kw = dict(center=0, disp=0, df=3, scale=1, nugget=1e-10)
    a = gb.ConjugateGaussianProcess(RBF(0.4) + WhiteKernel(1e-4, 'fixed'), decomposition='eig', **kw).fit(X, y)
    b = gb.ConjugateGaussianProcess(RBF(0.4) + WhiteKernel(1e-4, 'fixed'), **kw).fit(X, y)
    la, lb = np.exp(a.kernel_.theta[0]), np.exp(b.kernel_.theta[0])
    assert 0.1 < la < 0.3 and abs(la - lb) < 1e-4 * lb
    assert relerr(np.array(a.log_marginal_likelihood_value_), np.array(b.log_marginal_likelihood_value_)) < 1e-7


def test_eig_route_grid_sharded_nccl_single_rank(ctx, golden):
    """The 'eig' grid through the sharded path (length scales dealt over the ranks, one all-gather) on a one-rank NCCL
    group: identical to the plain call."""
    import os
    import torch
    import torch.distributed as dist
    g = golden("eig_route")
    tgp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1,
                          optimizer=None, decomposition='eig').fit(g["Xt"], g["yt"], orders=g["orders"])
    want = tgp.log_marginal_likelihood_grid(g["ls_vals"], g["q_vals"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29900 + os.getpid() % 90))
    torch.cuda.set_device(0)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        got = tgp.log_marginal_likelihood_grid(g["ls_vals"], g["q_vals"], group=dist.group.WORLD)
    finally:
        if created:
            dist.destroy_process_group()
    assert np.array_equal(got, want)
