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
