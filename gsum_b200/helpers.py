"""Series and linear-algebra helpers with the reference's names and signatures (gsum/helpers.py).

`coefficients`, `partials`, `geometric_sum` and `cartesian` are O(N·n) elementwise host utilities — the
grid path does not call them (the coefficient extraction is fused into the device RHS staging,
csrc/lml.cuh:stage_rhs_kernel); they exist so user code written against gsum keeps working.
`pivoted_cholesky`, `cholesky_errors` and `mahalanobis` run on the GPU through the C ABI.
"""
from __future__ import annotations

import numpy as np

from . import ops

__all__ = ["cartesian", "coefficients", "partials", "geometric_sum", "pivoted_cholesky", "cholesky_errors",
           "mahalanobis"]


def cartesian(*arrays):
    """Cartesian product of 1-d arrays; earlier arrays vary slowest (gsum/helpers.py:19-33)."""
    grids = np.meshgrid(*arrays, indexing="ij")
    return np.stack(grids, axis=-1).reshape(-1, len(arrays))


def _order_differences(y):
    """[y_0, y_1 - y_0, ...]: the order-by-order corrections (gsum/helpers.py:98-99)."""
    dy = np.empty_like(y, dtype=np.float64)
    dy[..., 0] = y[..., 0]
    dy[..., 1:] = y[..., 1:] - y[..., :-1]
    return dy


def coefficients(y, ratio, ref=1, orders=None):
    """Coefficients c_n = (y_n - y_{n-1}) / (ref * ratio**n) of a power series (gsum/helpers.py:71-101).

    y : (n_samples, n_curves) partial sums; ratio, ref : scalar or (n_samples,); orders : (n_curves,).
    """
    y = np.asarray(y)
    if y.ndim != 2:
        raise ValueError("y must be 2d")
    if orders is None:
        orders = np.arange(y.shape[-1])
    if len(orders) != y.shape[-1]:
        raise ValueError("partials and orders must have the same length")
    ref, ratio, orders = np.atleast_1d(ref, ratio, orders)
    return _order_differences(y) / (ref[:, None] * ratio[:, None] ** orders)


def partials(coeffs, ratio, ref=1, orders=None):
    """Partial sums y_k = ref * sum_{n<=k} c_n ratio**n (gsum/helpers.py:104-146)."""
    coeffs = np.asarray(coeffs)
    if orders is None:
        orders = np.arange(coeffs.shape[-1])
    ratio, ref = np.atleast_1d(ratio), np.atleast_1d(ref)
    if ratio.ndim == 1:
        ratio = ratio[:, None]
    if ref.ndim == 1:
        ref = ref[:, None]
    return np.cumsum(ref * coeffs * ratio ** orders, axis=-1)


def geometric_sum(x, start, end, excluded=None):
    """sum_{i=start}^{end} x**i without the `excluded` powers; `end` may be inf (gsum/helpers.py:149-182)."""
    if end < start:
        raise ValueError("end must be greater than or equal to start")
    total = (x ** start - x ** (end + 1)) / (1 - x)
    if excluded is not None:
        for n in np.atleast_1d(excluded):
            if start <= n <= end:
                total -= x ** n
    return total


def pivoted_cholesky(M):
    """G with M = G Gᵀ from a pivoted Cholesky, rows in the original order (gsum/helpers.py:185-199).

    Same pivoting rule as LAPACK ``dpstrf`` (see csrc/diag.cuh); raises LinAlgError when M is not
    positive definite to working precision, like the reference.
    """
    G, _, _, _, status = ops.pivoted_cholesky(M)
    if status > 0:
        raise np.linalg.LinAlgError("M is not positive-semidefinite")
    return G


def cholesky_errors(y, mean, chol):
    """L^{-1}(y - mean) for y of shape (n_curves, N) or (N,) (gsum/helpers.py:504-505)."""
    y = np.asarray(y, dtype=np.float64)
    single = y.ndim == 1
    E, _ = ops.cholesky_errors(chol, mean, np.ascontiguousarray(np.atleast_2d(y).T), want_errors=True)
    return E[:, 0] if single else E.T


def mahalanobis(y, mean, chol=None, inv=None, sqrt_mat=None):
    """Mahalanobis distance of each curve in y (n_curves, N) (gsum/helpers.py:512-522).

    `chol`: norms of L^{-1}(y - mean) (forward solve on the device).  `inv`: sqrt of (y - mean)^T inv (y - mean) per curve
    (quadratic forms on the device; the reference's np.diag of the full product).  `sqrt_mat`: the reference calls
    `numpy.linalg.solve(sqrt_mat, ..., lower=True)` (helpers.py:508-509), which raises TypeError on every numpy — the
    same error is raised here, there is no result to reproduce."""
    if (chol is not None) and (inv is not None) and (sqrt_mat is not None):
        raise ValueError("Only one of chol, inv, or sqrt_mat can be given")
    if chol is None and sqrt_mat is not None:
        raise TypeError("solve() got an unexpected keyword argument 'lower'")
    if chol is None:
        if inv is None:
            raise TypeError("mahalanobis needs one of chol, inv, sqrt_mat")
        y2 = np.atleast_2d(np.asarray(y, dtype=np.float64))
        q = ops.quadratic_forms(inv, mean, np.ascontiguousarray(y2.T))
        return np.squeeze(np.sqrt(q))
    y = np.asarray(y, dtype=np.float64)
    single = y.ndim == 1
    _, md2 = ops.cholesky_errors(chol, mean, np.ascontiguousarray(np.atleast_2d(y).T), want_errors=False, want_md2=True)
    md = np.sqrt(md2)
    return md[0] if single else md
