"""Series and linear-algebra helpers with the reference's names and signatures (gsum/helpers.py).

`coefficients`, `partials`, `geometric_sum` and `cartesian` are O(N·n) elementwise host utilities — the
grid path does not call them (the coefficient extraction is fused into the device RHS staging,
csrc/lml.cuh:stage_rhs_kernel); they exist so user code written against gsum keeps working.
`pivoted_cholesky`, `cholesky_errors`, `mahalanobis`, the correlation functions `rbf` / `gaussian` and `kl_gauss` run on
the GPU through the C ABI.  `predictions`, `hpd`, `hpd_pdf` and `median_pdf` are summaries of ONE univariate distribution
or of one tabulated pdf (a few quantile calls, a 1-d trapezoid rule): presentation utilities either side of the path with
no array arithmetic to move to the device, kept so that `from gsum import ...` lines keep working.
"""
from __future__ import annotations

import functools
import inspect

import numpy as np

from . import ops

__all__ = ["cartesian", "coefficients", "partials", "geometric_sum", "pivoted_cholesky", "cholesky_errors",
           "mahalanobis", "stabilize", "rbf", "gaussian", "kl_gauss", "predictions", "hpd", "hpd_pdf", "median_pdf",
           "lazy_property", "default_attributes"]


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


def stabilize(M):
    """M + 1e-5 I (gsum/helpers.py:202-203)."""
    M = np.asarray(M)
    return M + 1e-5 * np.eye(*M.shape)


def _correlation(X, Xp, ls):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be 2d: (n_samples, n_features)")
    if Xp is not None:
        Xp = np.asarray(Xp, dtype=np.float64)
    ls = np.asarray(ls, dtype=np.float64)
    if ls.ndim > 1 or (ls.ndim == 1 and ls.shape[0] not in (1, X.shape[1])):
        raise ValueError("ls must be a scalar or one length scale per feature")
    return ops.kernel_matrix(X, Xp, ls.reshape(-1), 1.0, 0.0)


def gaussian(X, Xp=None, ls=1):
    """exp(-|x - x'|^2 / (2 ls^2)) for X (N, d) -> (N, N), on the device (K1) (gsum/helpers.py:233-251).

    With an explicit Xp (M, d) the reference rescales X by ls but not Xp (helpers.py:242-246), i.e. it returns
    exp(-|x / ls - x'|^2 / 2); that behaviour is kept.  The reference expands the square (|x|^2 + |x'|^2 - 2 x.x', clipped
    at 0), which loses ~eps |x|^2 / ls^2 in the exponent; the device kernel takes the differences first, so entries agree
    to that level (rtol 1e-10 for |x| / ls up to ~1e2) and the diagonal of gaussian(X) is exactly 1."""
    if Xp is None:
        return _correlation(X, None, ls)
    return _correlation(np.asarray(X, dtype=np.float64) * 1.0 / ls, Xp, 1.0)


def rbf(X, Xp=None, ls=1):
    """exp(-|x - x'|^2 / (2 ls^2)), the indicator of x == x' when ls == 0 (gsum/helpers.py:254-261); on the device (K1)."""
    if np.ndim(ls) == 0 and ls == 0:
        X = np.asarray(X, dtype=np.float64)
        Xp = X if Xp is None else np.asarray(Xp, dtype=np.float64)
        return np.where((X[:, None, :] == Xp[None, :, :]).all(axis=-1), 1.0, 0.0)
    return _correlation(X, Xp, ls)


def kl_gauss(mu0, cov0, mu1, cov1=None, chol1=None):
    """D_KL(N(mu0, cov0) || N(mu1, cov1)) = [tr(cov1^-1 cov0) + (mu1 - mu0)^T cov1^-1 (mu1 - mu0) - k + ln det cov1 / det cov0] / 2
    (gsum/helpers.py:310-368).  Exactly one of `cov1` (factored after `stabilize`, as the reference does) and `chol1` (its
    lower Cholesky factor) must be given.

    Device work: the Cholesky factor of cov1 (K2), cov1^-1 cov0 by a two-sided triangular solve (K3), the quadratic form by a
    forward solve with its squared norm (K6), and ln det cov0 from a device Cholesky of cov0 (the reference's LU-based
    `slogdet` gives the same number for a covariance matrix; a cov0 that is not positive definite raises LinAlgError here)."""
    mu0, mu1 = np.atleast_1d(np.asarray(mu0, dtype=np.float64)), np.atleast_1d(np.asarray(mu1, dtype=np.float64))
    cov0 = np.atleast_2d(np.asarray(cov0, dtype=np.float64))
    if chol1 is not None and cov1 is None:
        chol1 = np.ascontiguousarray(np.atleast_2d(np.asarray(chol1, dtype=np.float64)))
    elif cov1 is not None and chol1 is None:
        chol1 = ops.cholesky(stabilize(np.atleast_2d(np.asarray(cov1, dtype=np.float64))))
    else:
        raise ValueError("Exactly one of cov1 or chol1 must be given.")
    k = cov0.shape[0]
    _, info, logdet0 = ops.cholesky(cov0, return_info=True)
    if info:
        raise np.linalg.LinAlgError("cov0 is not positive definite")
    logdet1 = 2.0 * float(np.sum(np.log(np.diag(chol1))))
    _, md2 = ops.cholesky_errors(chol1, mu0, np.ascontiguousarray(np.broadcast_to(mu1, (k,))[:, None]), want_errors=False, want_md2=True)
    tr_mat = float(np.trace(ops.cho_solve(chol1, cov0)))
    return 0.5 * (tr_mat + float(md2[0]) - k + logdet1 - float(logdet0))


def predictions(dist, dob=None):
    """Mean of a frozen scipy distribution and, for degrees of belief `dob`, its central intervals with shape
    (len(dob), 2, len(mean)) squeezed (gsum/helpers.py:206-230)."""
    mean = dist.mean()
    if dob is None:
        return mean
    levels = np.atleast_2d(dob).T
    bounds = np.asarray(dist.interval(levels))                       # (2, len(dob), len(mean))
    return mean, np.squeeze(np.swapaxes(bounds, 0, 1))


def hpd(dist, alpha, *args):
    """Highest-probability-density interval of mass `alpha` of a univariate scipy distribution (frozen, or a family with
    its shape arguments in *args): the CDF window [s, s + alpha] of smallest width, found by a Nelder-Mead search over s
    started at 1 - alpha, as the reference does (gsum/helpers.py:264-278)."""
    from scipy.optimize import fmin
    if args:
        dist = dist(*args)
    width = lambda s: dist.ppf(s + alpha) - dist.ppf(s)
    start = fmin(width, 1 - alpha, ftol=1e-8, disp=False)[0]
    return dist.ppf([start, alpha + start])


def hpd_pdf(pdf, alpha, x):
    """Highest-density interval of mass `alpha` of a pdf tabulated on x: the level set {pdf > p} whose trapezoid-rule mass
    over {pdf >= p} is closest to alpha, p scanned over the tabulated heights (gsum/helpers.py:281-295)."""
    pdf, x = np.asarray(pdf), np.asarray(x)
    heights = np.unique(pdf)
    miss = np.array([(np.trapezoid(pdf[pdf >= h], x=x[pdf >= h]) - alpha) ** 2 for h in heights])
    inside = x[pdf > heights[np.argmin(miss)]]
    return np.array([np.min(inside), np.max(inside)])


def median_pdf(pdf, x):
    """First grid point at which the trapezoid-rule CDF of the tabulated pdf exceeds 1/2 (gsum/helpers.py:298-307)."""
    pdf, x = np.asarray(pdf), np.asarray(x)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(x))])
    above = np.nonzero(cdf > 0.5)[0]
    return x[above[0]] if above.size else x[-1]


def lazy_property(function):
    """Read-only property evaluated on first access and kept on the instance as `_cache_<name>` (gsum/helpers.py:371-386)."""
    slot = "_cache_" + function.__name__

    @functools.wraps(function)
    def getter(self):
        try:
            return getattr(self, slot)
        except AttributeError:
            value = function(self)
            setattr(self, slot, value)
            return value
    return property(getter)


def default_attributes(**kws):
    """Method decorator: an argument left at None (or an empty *args / **kwargs) is replaced by the instance attribute named
    in `kws` for that parameter — `@default_attributes(x='x', y='_y')` makes `def f(self, x=None, y=None)` default to
    `self.x`, `self._y` at call time (gsum/helpers.py:416-501).  numpy arrays are never treated as missing."""
    def decorator(function):
        sig = inspect.signature(function)
        P = inspect.Parameter

        def missing(value, kind):
            if isinstance(value, np.ndarray):
                return False
            if kind in (P.POSITIONAL_OR_KEYWORD, P.KEYWORD_ONLY):
                return value is None
            if kind == P.VAR_POSITIONAL:
                return value == ()
            return kind == P.VAR_KEYWORD and value == {}

        @functools.wraps(function)
        def wrapper(self, *args, **kwargs):
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            for name in list(bound.arguments):
                if name in kws and missing(bound.arguments[name], sig.parameters[name].kind):
                    bound.arguments[name] = getattr(self, kws[name])
            return function(*bound.args, **bound.kwargs)
        return wrapper
    return decorator
