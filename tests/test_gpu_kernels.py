"""GPU: the primitive entry points (K1 kernel matrix, K2 Cholesky, K3 triangular solves) against numpy/scipy/sklearn."""
import numpy as np
import pytest
from scipy.linalg import cho_solve, solve_triangular
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, WhiteKernel

from gsum_b200 import ops
from util import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d", [(1, 1), (50, 1), (200, 1), (130, 2), (65, 3)])
def test_kernel_matrix(ctx, n, d):
    rs = np.random.RandomState(n)
    X, X2 = rs.rand(n, d), rs.rand(77, d)
    ls = [0.2] if d == 1 else list(0.1 + 0.3 * rs.rand(d))
    kern = C(1.5) * RBF(ls if d > 1 else ls[0]) + WhiteKernel(1e-6)
    K = ops.kernel_matrix(X, None, ls, 1.5, 1e-6)
    Kr = kern(X)
    assert np.array_equal(np.diag(K), np.diag(Kr))                # diagonal is exact: (c * 1 + noise)
    assert np.max(np.abs(K - Kr)) < 4e-16                          # exp() within an ulp of glibc's
    assert np.array_equal(K, K.T)
    assert np.max(np.abs(ops.kernel_matrix(X, X2, ls, 1.5, 1e-6) - kern(X, X2))) < 4e-16    # WhiteKernel is zero for Y given


@pytest.mark.parametrize("n,batch", [(1, 1), (50, 3), (64, 2), (65, 2), (200, 4), (512, 2)])
def test_cholesky(ctx, n, batch):
    X = np.sort(np.random.RandomState(n).rand(n))[:, None]
    A = np.stack([RBF(0.05 * (b + 1))(X) + 1e-4 * np.eye(n) for b in range(batch)])
    L, info, logdet = ops.cholesky(A, return_info=True)
    Lr = np.linalg.cholesky(A)
    assert not info.any()
    assert np.all(np.triu(L, 1) == 0.0)
    assert relerr(L @ np.swapaxes(L, 1, 2), A) < 5e-15                 # backward error at the rounding level
    assert relerr(L, Lr) < 1e-10
    assert np.allclose(logdet, 2 * np.log(np.diagonal(Lr, axis1=1, axis2=2)).sum(1), rtol=1e-11, atol=1e-9)
    assert relerr(ops.cholesky(A[0]), Lr[0]) < 1e-10


def test_cholesky_not_positive_definite(ctx):
    A = np.eye(100)
    A[70, 70] = -1.0
    _, info, _ = ops.cholesky(A, return_info=True)
    assert info == 71                                                   # LAPACK potrf convention: 1-based leading minor
    with pytest.raises(np.linalg.LinAlgError):
        ops.cholesky(A)
    ok = np.eye(100)
    L, info, _ = ops.cholesky(np.stack([ok, A, 4 * ok]), return_info=True)
    assert info.tolist() == [0, 71, 0] and np.array_equal(L[2], 2 * ok)  # a failing matrix does not disturb its batch mates


@pytest.mark.parametrize("n,k", [(50, 3), (200, 70), (333, 5), (128, 129)])
def test_cho_solve(ctx, n, k):
    rs = np.random.RandomState(n + k)
    X = np.sort(rs.rand(n))[:, None]
    L = np.linalg.cholesky(RBF(0.1)(X) + 1e-4 * np.eye(n))
    B = rs.randn(n, k)
    assert relerr(ops.cho_solve(L, B), cho_solve((L, True), B)) < 1e-11
    assert relerr(ops.cho_solve(L, B, forward_only=True), solve_triangular(L, B, lower=True)) < 1e-12
    assert relerr(ops.cho_solve(L, B[:, 0]), cho_solve((L, True), B[:, 0])) < 1e-11


def test_cholesky_batch_under_contention_is_deterministic(ctx):
    """64 ill-conditioned matrices (cond ~ 1e8) factored together, three times: no spurious non-positive pivot, identical
    bits every time.  Guards the intra-POTRF ordering (a factor-writing lane overtaking warps that still load the
    unfactored block showed up exactly here: sporadic info = 8 j + 1 for the smooth matrices of the batch)."""
    n = 256
    X = np.linspace(0, 1, n)[:, None]
    A = np.stack([RBF(l)(X) + 1e-6 * np.eye(n) for l in np.geomspace(0.01, 0.3, 64)])
    first = None
    for _ in range(3):
        L, info, logdet = ops.cholesky(A, return_info=True)
        assert not info.any()
        if first is None:
            first = (L, logdet)
        else:
            assert np.array_equal(L, first[0]) and np.array_equal(logdet, first[1])
    assert relerr(first[0] @ np.swapaxes(first[0], 1, 2), A) < 5e-15
    # single-tile matrices exercise the diagonal-tile workers alone
    n = 64
    X = np.linspace(0, 1, n)[:, None]
    A = np.stack([RBF(l)(X) + 1e-6 * np.eye(n) for l in np.geomspace(0.01, 0.3, 100)])
    L, info, _ = ops.cholesky(A, return_info=True)
    assert not info.any() and relerr(L @ np.swapaxes(L, 1, 2), A) < 5e-15


@pytest.mark.parametrize("n", [300, 1024])
def test_chain_mode_is_bit_identical_to_many_matrices_schedule(n, monkeypatch):
    """The two regimes of the factorisation — chain mode (few matrices: one chain CTA per matrix owns the diagonal band,
    csrc/chain.cuh) and the many-matrices schedule (factor CTAs claim diagonal tiles) — perform the same operations on every
    element in the same order: factor, forward solve and likelihood grid are bit-identical, so a cell does not depend on how
    many length scales share its launch (a shard of 16 and the 1-GPU grid of 128 give the same numbers)."""
    from gsum_b200 import _lib
    monkeypatch.setenv("GSUM_B200_CHAIN_MAX", "0")
    many = _lib.Context(0)
    monkeypatch.delenv("GSUM_B200_CHAIN_MAX")
    chain = _lib.Context(0)
    try:
        rs = np.random.RandomState(3)
        X = np.sort(rs.rand(n))[:, None]
        A = np.stack([RBF(0.05 * (b + 1))(X) + 1e-5 * np.eye(n) for b in range(5)])
        Lc, info, logdet_c = ops.cholesky(A, return_info=True, ctx=chain)
        Lm, info_m, logdet_m = ops.cholesky(A, return_info=True, ctx=many)
        assert not info.any() and not info_m.any()
        assert np.array_equal(Lc, Lm) and np.array_equal(logdet_c, logdet_m)
        assert relerr(Lc, np.linalg.cholesky(A)) < 1e-10
        B = rs.randn(n, 70)
        assert relerr(ops.cho_solve(Lc[0], B, ctx=chain), cho_solve((Lc[0], True), B)) < 1e-10
        dy = rs.randn(n, 4)
        args = (X, dy, 1.0, np.arange(4), np.array([[0.05], [0.2], [0.11]]), np.array([0.3, 0.5, 0.7]))
        ll_c = ops.lml_grid(*args, noise=1e-5, nugget=1e-10, ctx=chain)
        ll_m = ops.lml_grid(*args, noise=1e-5, nugget=1e-10, ctx=many)
        assert np.array_equal(ll_c, ll_m)
        # a failing matrix (negative diagonal entry) in the batch: same status, the healthy ones untouched
        A2 = A.copy()
        A2[2, 100, 100] = -1.0
        Lc2, info_c2, _ = ops.cholesky(A2, return_info=True, ctx=chain)
        Lm2, info_m2, _ = ops.cholesky(A2, return_info=True, ctx=many)
        assert np.array_equal(info_c2, info_m2) and info_c2[2] != 0 and np.array_equal(Lc2[[0, 1, 3, 4]], Lc[[0, 1, 3, 4]])
    finally:
        many.close()
        chain.close()


def test_large_single_matrix_8192(ctx):
    """Beyond the BASELINE sizes: one N = 8192 matrix (128 tile columns) through kernel build, Cholesky, solve and the
    cooperative pivoted Cholesky (64 CTAs) — checked through residuals, which need no O(N^3) host work."""
    from sklearn.gaussian_process.kernels import RBF
    n = 8192
    X = np.sort(np.random.RandomState(7).rand(n))[:, None]
    A = ops.kernel_matrix(X, None, 0.01, constant=1.0, noise=1e-3)
    ref = RBF(0.01)(X[4000:4200], X[100:400])
    assert np.max(np.abs(A[4000:4200, 100:400] - ref)) < 1e-15
    L = ops.cholesky(A)
    assert np.all(np.triu(L, 1) == 0)
    v = np.random.RandomState(0).randn(n, 3)
    Av = A @ v
    assert np.max(np.abs(L @ (L.T @ v) - Av)) / np.max(np.abs(Av)) < 1e-13
    assert np.max(np.abs(ops.cho_solve(L, Av) - v)) < 1e-8
    G, Lp, piv, rank, status = ops.pivoted_cholesky(A)
    assert status == 0 and rank == n and sorted(piv.tolist()) == list(range(n))
    assert np.max(np.abs(G @ (G.T @ v) - Av)) / np.max(np.abs(Av)) < 1e-13
    d = np.diag(Lp)
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-9))                 # pivoted Cholesky invariant: non-increasing diagonal
