"""GPU: Diagnostic (Mahalanobis, Cholesky / pivoted-Cholesky errors, draws, credible-interval coverage) vs the reference."""
import numpy as np
import pytest
import scipy.stats as st
from sklearn.gaussian_process.kernels import RBF

import gsum_b200 as gb
from gsum_b200 import ops
from oracle import gsum_oracle as o
from util import as_close_as_reference, ld_cholesky, ld_forward_solve, relerr

pytestmark = pytest.mark.gpu


def test_c5_golden(ctx, golden):
    g = golden("c5_diagnostics")
    d = gb.Diagnostic(g["mean"], g["cov"], random_state=1)
    Y = g["Y"]
    assert relerr(d._chol, g["chol"]) < 1e-10
    assert np.array_equal(d._piv, g["piv"])                                    # pivot order bit-exact vs LAPACK dpstrf
    assert relerr(d._pchol, g["pchol"]) < 1e-10
    assert relerr(d._pchol @ d._pchol.T, g["cov"]) < 1e-14
    assert relerr(d.md_squared(Y), g["md2"]) < 1e-10
    # cond(cov) ~ 2e7: the reference's own results (LAPACK trsm; LU on the row-permuted factor) sit 2.2e-10 / 2.9e-10 from the
    # extended-precision values, so rtol 1e-10 is arbitrated, not widened: the device must be as close to the exact errors as the
    # reference is (util.as_close_as_reference: <= 4 x the reference's distance + 1e-10)
    V = (Y.T - g["mean"]).T
    exact_ce = ld_forward_solve(ld_cholesky(g["cov"]), V).astype(float)
    piv = g["piv"]
    exact_pc = ld_forward_solve(ld_cholesky(g["cov"][np.ix_(piv, piv)]), V[piv]).astype(float)
    ce, pce = d.cholesky_errors(Y), d.pivoted_cholesky_errors(Y)
    assert as_close_as_reference(ce, g["chol_errors"], exact_ce, 1e-10), (relerr(ce, exact_ce), relerr(g["chol_errors"], exact_ce))
    assert as_close_as_reference(pce, g["pc_errors"], exact_pc, 1e-10), (relerr(pce, exact_pc), relerr(g["pc_errors"], exact_pc))
    assert relerr(d.individual_errors(Y), g["ind_errors"]) < 1e-15
    assert np.array_equal(d.credible_interval(Y, g["intervals"]), g["coverage"])   # integer counts / N: exact
    assert d.md_squared(Y[:, 0]) == pytest.approx(float(g["md2_1d"]), rel=1e-10)
    # 1-d y: the documented return shape is (n_intervals,).  (The reference transposes a 1-d y into N one-point "curves",
    # gsum/diagnostics.py:168 `np.atleast_2d(y).T`, and returns an (N, n_intervals) array — a reference bug, not mirrored.)
    assert g["coverage_1d"].shape == (len(g["mean"]), len(g["intervals"]))
    assert np.array_equal(d.credible_interval(Y[:, 0], g["intervals"]), g["coverage"][0])
    # property: sum of squared PC errors == MD^2 — an oracle-free cross-check of K6 against K7
    assert relerr((d.pivoted_cholesky_errors(Y) ** 2).sum(0), d.md_squared(Y)) < 1e-9
    assert relerr((d.cholesky_errors(Y) ** 2).sum(0), d.md_squared(Y)) < 1e-14


def test_pivoted_cholesky_known_answers(ctx, golden):
    """gsum/tests/test.py:75-122 (test_oracle_examples, atol 1e-4) and model_checking_tests.ipynb cell 6."""
    g = golden("kat_pivoted_cholesky")
    for i in range(3):
        np.testing.assert_allclose(g[f"table{i}"], gb.pivoted_cholesky(g[f"M{i}"]), atol=1e-4)
        np.testing.assert_allclose(g[f"G{i}"], gb.pivoted_cholesky(g[f"M{i}"]), atol=1e-12)
    G, Lp, piv, rank, status = ops.pivoted_cholesky(g["M_nb"])
    assert list(piv + 1) == [4, 1, 3, 2] and rank == 4 and status == 0
    np.testing.assert_allclose(G, g["G_nb"], atol=1e-14)
    assert np.all(np.triu(Lp, 1) == 0)


@pytest.mark.parametrize("n", [64, 65, 200, 700])
def test_pivoted_cholesky_vs_lapack(ctx, n):
    """Pivots identical to LAPACK's on covariances with separated Schur diagonals (random points, varying amplitude);
    blocked path (n > 64) included."""
    rs = np.random.RandomState(n)
    X = np.sort(rs.rand(n))[:, None]
    amp = 1.0 + 0.5 * rs.rand(n)
    cov = 1.3 * np.outer(amp, amp) * (RBF(0.2)(X) + 1e-5 * np.eye(n))
    G, Lp, piv, rank, status = ops.pivoted_cholesky(cov)
    Gr, pr = o.pivoted_cholesky(cov, return_pivots=True)
    assert status == 0 and rank == n and np.array_equal(piv, pr)
    assert relerr(G, Gr) < 1e-10 and relerr(G @ G.T, cov) < 1e-14
    inv = np.argsort(piv)
    assert np.array_equal(Lp[inv], G)


def test_pivoted_cholesky_rank_deficient(ctx):
    """info > 0 -> LinAlgError('M is not positive-semidefinite') like gsum/helpers.py:189-190; same rank and leading pivots."""
    from scipy.linalg.lapack import dpstrf
    A = np.random.RandomState(0).randn(100, 30)
    M = A @ A.T
    _, p, r, info = dpstrf(M, lower=True)
    G, Lp, piv, rank, status = ops.pivoted_cholesky(M)
    assert status == 1 and rank == r == 30 and np.array_equal(piv[:rank], (p - 1)[:r])
    with pytest.raises(np.linalg.LinAlgError):
        gb.pivoted_cholesky(M)
    with pytest.raises(np.linalg.LinAlgError):
        gb.Diagnostic(np.zeros(100), M)


def test_symmetric_grid_ties_documented(ctx):
    """On a regular grid the Schur-complement diagonals tie exactly by symmetry; rounding then decides and the order may
    differ from LAPACK's.  The factorisation itself must still be a valid pivoted Cholesky."""
    n = 256
    X = np.linspace(0, 1, n)[:, None]
    cov = 1.3 * (RBF(0.2)(X) + 1e-5 * np.eye(n))
    G, Lp, piv, rank, status = ops.pivoted_cholesky(cov)
    assert status == 0 and sorted(piv.tolist()) == list(range(n)) and piv[0] == 0 and piv[1] == n - 1
    assert relerr(G @ G.T, cov) < 1e-14
    d = np.diag(Lp)
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-9))          # pivoted Cholesky invariant: non-increasing diagonal
    # Where exactly the order can leave LAPACK's: the device mirrors dpstrf's algorithm (block size 64, the
    # running diagonal work(i) += a(j,i)^2 inside a block, first maximum wins), but the SUMMATION ORDER inside LAPACK's DGEMV /
    # DSYRK calls belongs to the BLAS build (SIMD partial sums) and cannot be mirrored.  So: either the pivots are dpstrf's, or at
    # the first position where they differ the two candidates' Schur-complement diagonals are equal to rounding (a tie of the
    # symmetric grid, x <-> 1 - x), and the two orders are mirror images of each other from there on.
    Gr, pr = o.pivoted_cholesky(cov, return_pivots=True)
    if not np.array_equal(piv, pr):
        j = int(np.nonzero(piv != pr)[0][0])
        pos = np.argsort(piv)                                               # original index -> position in the device order
        schur = lambda c: cov[c, c] - np.sum(Lp[pos[c], :j].astype(np.longdouble) ** 2)
        da, db = schur(piv[j]), schur(pr[j])
        assert abs(da - db) <= 1e-9 * abs(da), (j, piv[j], pr[j], da, db)
        assert piv[j] + pr[j] == n - 1                                      # the mirror point of the same tie


def test_draws_with_supplied_z_and_helpers(ctx, golden):
    g = golden("c5_diagnostics")
    n = len(g["mean"])
    Z = np.random.RandomState(0).standard_normal((n, 37))
    D, _ = ops.draws(g["chol"], g["mean"], Z=Z)
    assert relerr(D, o.draws_from_z(g["mean"], g["chol"], Z)) < 1e-14
    assert relerr(gb.mahalanobis(D.T, g["mean"], g["chol"]), o.mahalanobis(D.T, g["mean"], g["chol"])) < 1e-10
    assert relerr(gb.cholesky_errors(D.T, g["mean"], g["chol"]), o.cholesky_errors(D.T, g["mean"], g["chol"])) < 1e-9
    assert gb.mahalanobis(D[:, 0], g["mean"], g["chol"]) == pytest.approx(o.mahalanobis(D[:, 0], g["mean"], g["chol"]), rel=1e-10)
    # examples/model_checking_tests.ipynb cell 2: chol-based and pinv-based Mahalanobis distances agree
    y = D[:, 3]
    pin = np.linalg.pinv(g["cov"])
    assert gb.mahalanobis(y, g["mean"], g["chol"]) == pytest.approx(np.sqrt((y - g["mean"]) @ pin @ (y - g["mean"])), rel=1e-5)


def test_device_rng_draws_and_fused_coverage(ctx, golden):
    """Philox draws: moments, MD^2 ~ chi2(N), and the fused coverage equals the separate coverage kernel bit for bit."""
    g = golden("c5_diagnostics")
    d = gb.Diagnostic(g["mean"], g["cov"], random_state=3)
    n = len(g["mean"])
    S = d.samples(4000, device_rng=True)
    assert S.shape == (n, 4000)
    sd = np.sqrt(np.diag(g["cov"]))
    assert np.max(np.abs(S.mean(1) - g["mean"]) / sd) < 5 / np.sqrt(4000)
    md2 = d.md_squared(S)
    assert abs(md2.mean() - n) < 5 * np.sqrt(2 * n / 4000)
    assert st.kstest(md2, st.chi2(n).cdf).pvalue > 1e-3
    iv = g["intervals"]
    cov_fused = d.sample_coverage(4000, iv)
    assert np.array_equal(cov_fused, d.credible_interval(S, iv))
    assert np.max(np.abs(cov_fused.mean(0) - iv)) < 0.01                      # calibrated: coverage ~ nominal level
    assert np.array_equal(S, d.samples(4000, device_rng=True))                # counter-based: reproducible
    host = d.samples(16)                                                       # numpy RandomState stream through m + L z
    z = np.random.RandomState(3).standard_normal((n, 16))
    assert relerr(host, o.draws_from_z(g["mean"], d._chol, z)) < 1e-13


def test_large_draw_count_grid_limits(ctx):
    """> 65535 draws: exercises the row-on-x launch geometry of the staging kernels."""
    n = 96
    X = np.linspace(0, 1, n)[:, None]
    cov = RBF(0.3)(X) + 1e-3 * np.eye(n)
    L = np.linalg.cholesky(cov)
    iv = np.linspace(0.05, 0.95, 7)
    sd = np.sqrt(np.diag(cov))
    lower, upper = st.norm(loc=np.zeros(n), scale=sd).interval(np.atleast_2d(iv).T)
    _, cv = ops.draws(L, np.zeros(n), n_draws=70000, seed=11, lower=lower, upper=upper, want_draws=False)
    assert cv.shape == (70000, 7) and np.max(np.abs(cv.mean(0) - iv)) < 0.01


def test_c5_full_size_properties(ctx):
    """Config C5 at full size (N = 4096; 64 held-out curves; 1e5 draws): pivot vector bit-exact against LAPACK dpstrf,
    G G^T = cov, sum of squared PC errors == MD^2 == sum of squared Cholesky errors, MD^2 ~ chi2(N), and the fused
    draw + coverage pass is calibrated (coverage of 1e5 draws within 0.2 % of the nominal level at 101 levels)."""
    from scipy.linalg.lapack import dpstrf
    n = 4096
    X = np.linspace(0, 1, n)[:, None]
    cov = 1.3 * (RBF(0.2)(X) + 1e-5 * np.eye(n))
    mean = np.zeros(n)
    d = gb.Diagnostic(mean, cov, random_state=4)
    c_, p_, r_, info = dpstrf(cov, lower=True)
    assert info == 0 and r_ == n and np.array_equal(d._piv, p_ - 1)
    assert relerr(d._pchol @ d._pchol.T, cov) < 1e-13
    Y = d.samples(64)
    md2 = d.md_squared(Y)
    E, Ep = d.cholesky_errors(Y), d.pivoted_cholesky_errors(Y)
    assert relerr((E ** 2).sum(0), md2) < 1e-12 and relerr((Ep ** 2).sum(0), md2) < 1e-8
    assert abs(md2.mean() - n) < 6 * np.sqrt(2 * n / 64)
    iv = np.linspace(0, 1, 101)
    cv = d.sample_coverage(100000, iv)
    assert cv.shape == (100000, 101) and np.max(np.abs(cv.mean(0) - iv)) < 2e-3


def test_c5_student_diag_golden_and_kl(ctx, golden):
    """Student-t Diagnostic (gsum/diagnostics.py:51-55: multivariate-t draws, t interval end points) and Diagnostic.kl
    (116-146) against outputs of the reference itself (tests/golden/make_golden_student_diag.py)."""
    from test_oracle import student_diag_covs
    g = golden("c5_student_diag")
    cov, cov0 = student_diag_covs(g)
    mean, df, Y = g["mean"], float(g["df"]), g["Y"]
    d = gb.Diagnostic(mean, cov, df=df, random_state=3)
    # deterministic half of MVT.rvs: m + sqrt((df-2)/df) L z / sqrt(x) with the reference run's own (z, x)
    scale = np.sqrt((df - 2.0) / df) / np.sqrt(g["x"])
    Lh = np.linalg.cholesky(cov)
    D, _ = ops.draws(Lh, mean, Z=g["z"], draw_scale=scale)                             # same factor: isolates the draw pass
    assert relerr(D, mean[:, None] + (Lh @ g["z"]) * scale[None, :]) < 1e-14
    assert relerr(D, Y) < 1e-9              # the reference factors sigma = cov (df-2)/df itself: 6e-11 apart at cond 1.4e7
    D, _ = ops.draws(d._chol, mean, Z=g["z"], draw_scale=scale)                        # device factor: cond(cov) ~ 1e6 * eps
    assert relerr(D, Y) < 1e-9
    assert np.array_equal(d.credible_interval(Y, g["intervals"]), g["coverage"])     # t(df, mean, sd) end points, integer counts
    assert relerr(d.md_squared(Y), g["md2"]) < 1e-10
    assert relerr(d.cholesky_errors(Y), g["chol_errors"]) < 1e-9
    assert relerr(d.pivoted_cholesky_errors(Y), g["pc_errors"]) < 1e-8
    assert relerr(d.individual_errors(Y), g["ind_errors"]) < 1e-15
    gd = gb.Diagnostic(mean, cov, random_state=1)
    assert gd.kl(g["mean0"], cov0) == pytest.approx(float(g["kl"]), rel=1e-10)
    assert gd.kl(mean, cov) == pytest.approx(float(g["kl_self"]), rel=1e-10)
    # facade draws: numpy stream for (z, chi2) through the same device pass; device generator: heavy tails, covariance = cov
    rs = np.random.RandomState(3)
    z = rs.standard_normal((len(mean), 8))
    x = rs.chisquare(df, 8) / df
    assert relerr(d.samples(8), o.mvt_draws_from_z(mean, cov, df, z, x)) < 1e-9
    S = d.samples(20000, device_rng=True)
    sd = np.sqrt(np.diag(cov))
    assert np.max(np.abs(S.var(1) / sd ** 2 - 1.0)) < 0.15                           # cov of MVT(sigma = cov (df-2)/df) is cov
    md2 = d.md_squared(S) / len(mean) * df / (df - 2.0)                                 # ~ F(N, df)
    assert st.kstest(md2, st.f(len(mean), df).cdf).pvalue > 1e-3


def test_draw_axis_shards_reproduce_the_unsharded_run(ctx, golden):
    """slices of the draw axis generated with `first_draw` equal the unsharded call column for column,
    and the int64 coverage counts add up exactly — Gaussian and Student-t."""
    from gsum_b200.distributed import shard_range
    g = golden("c5_diagnostics")
    iv = g["intervals"][::10]
    n, nd = len(g["mean"]), 1000
    for df in (None, 5.0):
        d = gb.Diagnostic(g["mean"], g["cov"], df=df, random_state=9)
        cov_full, cnt_full = d.sample_coverage(nd, iv, counts=True)
        assert np.array_equal(cnt_full, np.rint(cov_full * n).astype(np.int64).sum(0))
        parts, cnts = [], np.zeros(len(iv), dtype=np.int64)
        for r in range(3):
            lo, hi = shard_range(nd, 3, r)
            c, k = d.sample_coverage(hi - lo, iv, first_draw=lo, n_total=nd, counts=True)
            parts.append(c)
            cnts += k
        assert np.array_equal(np.concatenate(parts), cov_full) and np.array_equal(cnts, cnt_full)
    L, mean = d._chol, g["mean"]
    full, _ = ops.draws(L, mean, n_draws=200, seed=4)
    part, _ = ops.draws(L, mean, n_draws=70, seed=4, first_draw=130)
    assert np.array_equal(part, full[:, 130:])


def test_sharded_draws_and_predict_nccl_single_rank(ctx, golden):
    """The NCCL branches of `sample_coverage_sharded` / `predict_sharded` on a one-rank group equal the plain calls."""
    import os
    import torch
    import torch.distributed as dist
    from gsum_b200.distributed import predict_sharded, sample_coverage_sharded
    g = golden("c5_diagnostics")
    iv = g["intervals"][::5]
    d = gb.Diagnostic(g["mean"], g["cov"], random_state=2)
    rs = np.random.RandomState(0)
    X = np.linspace(0, 1, 90)[:, None]
    y = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(90)) @ rs.randn(90, 3)
    from sklearn.gaussian_process.kernels import WhiteKernel
    gp = gb.ConjugateGaussianProcess(RBF(0.2, 'fixed') + WhiteKernel(1e-6, 'fixed'), center=0, disp=0, df=1, scale=1).fit(X, y)
    Xn = rs.rand(257, 1)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29900 + os.getpid() % 300))
    torch.cuda.set_device(0)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        cv = sample_coverage_sharded(d, 3000, iv, group=dist.group.WORLD)
        m, s = predict_sharded(gp, Xn, return_std=True, group=dist.group.WORLD)
    finally:
        if created:
            dist.destroy_process_group()
    cov1, cnt1 = d.sample_coverage(3000, iv, counts=True)
    assert np.array_equal(cv, cnt1.astype(np.float64) / (3000.0 * len(g["mean"]))) and relerr(cv, cov1.mean(0)) < 1e-14
    mw, sw = gp.predict(Xn, return_std=True)
    assert np.array_equal(m, mw) and np.array_equal(s, sw)


def test_coverage_nested_and_unordered_intervals_agree(ctx, golden):
    """Increasing credible levels give nested intervals (bisection + histogram path); the same levels in a shuffled order
    take the comparison-per-interval path.  Both must reproduce the reference's integer counts."""
    g = golden("c5_diagnostics")
    d = gb.Diagnostic(g["mean"], g["cov"], random_state=1)
    Y, iv = g["Y"], g["intervals"]
    perm = np.random.RandomState(0).permutation(len(iv))
    assert np.array_equal(d.credible_interval(Y, iv), g["coverage"])
    assert np.array_equal(d.credible_interval(Y, iv[perm]), g["coverage"][:, perm])
    assert np.array_equal(d.credible_interval(Y, iv[perm]), o.credible_interval(Y, g["mean"], g["cov"], iv[perm]))
    one = d.credible_interval(Y, iv[40:41])                                  # a single interval
    assert np.array_equal(one[:, 0], g["coverage"][:, 40])
    # a point exactly on an interval's end is outside (strict inequalities), also on the bisection path
    lower, upper = d._bounds(iv)
    Yb = Y.copy()
    Yb[:, 0] = upper[50]
    Yb[:, 1] = lower[70]
    assert np.array_equal(d.credible_interval(Yb, iv), o.credible_interval(Yb, g["mean"], g["cov"], iv))


@pytest.mark.gpu
def test_mahalanobis_inv_and_sqrt_mat():
    """mahalanobis(inv=...) (gsum/helpers.py:521-522; the notebook's Mahalanobis identity uses it) against the numpy expression
    of the reference and against the `chol` route; sqrt_mat raises the reference's own TypeError (helpers.py:508-509)."""
    import gsum_b200 as gb
    rs = np.random.RandomState(5)
    for n, k in [(7, 1), (64, 3), (301, 11)]:
        X = np.linspace(0, 1, n)[:, None]
        cov = 1.3 * (RBF(0.2)(X) + 1e-5 * np.eye(n))
        mean = rs.randn(n)
        L = np.linalg.cholesky(cov)
        y = mean + (L @ rs.randn(n, k)).T
        inv = np.linalg.inv(cov)
        got = gb.mahalanobis(y, mean, inv=inv)
        want = np.squeeze(np.sqrt(np.diag((y - mean) @ inv @ (y - mean).T)))
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-10       # entries of inv ~ 1e5: the sum cancels five digits
        via_chol = gb.mahalanobis(y, mean, chol=L)
        assert np.max(np.abs(np.atleast_1d(got) - via_chol) / via_chol) < 1e-6        # cond(cov) ~ 1e5 in inv
    with pytest.raises(TypeError):
        gb.mahalanobis(y, mean, sqrt_mat=L)
