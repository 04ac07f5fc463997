"""Synthetic inputs of the path: Gaussian-process coefficient draws summed into EFT-style partial sums (gsum/datasets.py).

The generators sit immediately upstream of `TruncationGP.fit` — SURVEY.md 8(d) builds every benchmark configuration from this
recipe — and do the path's own linear algebra in reverse: kernel matrix (K1), a rank-revealing factorisation (K7, the device
`dpstrf`) and `mean + G z` (K9, the draws kernel).  All three run on the device through the C ABI; the host supplies the
standard normals from the caller's `random_state` and the O(N n) series arithmetic of `partials`.

The reference samples through `scipy.stats.multivariate_normal(...).rvs`, i.e. numpy's SVD-based sampler.  A Cholesky-type
factor applied to the same normals gives the same distribution but not the same numbers (any square root of K is as good as
another; the signs of singular vectors are a LAPACK accident), so parity here is distributional — the same statement
`BaseConjugateProcess.sample_y` makes — and the tests check the draws against K itself.
"""
from __future__ import annotations

import numpy as np
from sklearn.gaussian_process.kernels import RBF
from sklearn.utils import check_random_state

from . import ops
from .helpers import cartesian, gaussian, partials
from .kernels import flatten_kernel

__all__ = ["make_gaussian_partial_sums", "make_gaussian_partial_sums_uniform", "make_gaussian_partial_sums_on_grid", "toy_data",
           "generate_coefficients"]


def _gaussian_draws(mean, K, n_draws, rng, allow_singular=True):
    """mean[:, None] + G z with K = G G^T from the device pivoted Cholesky and z (n, n_draws) standard normals of `rng`.

    `dpstrf` stops at the numerical rank (remaining diagonal <= n eps max diag): the columns beyond it are dropped, which is
    what `allow_singular=True` means for `scipy.stats.multivariate_normal` (gsum/datasets.py:69, its eigenvalue cut-off is
    of the same order); with `allow_singular=False` a rank-deficient matrix is scipy's LinAlgError."""
    n = K.shape[0]
    _, Lp, piv, rank, _ = ops.pivoted_cholesky(K)
    if rank < n:
        if not allow_singular:
            raise np.linalg.LinAlgError("When `allow_singular is False`, the input matrix must be symmetric positive definite.")
        Lp[:, rank:] = 0.0
    inv = np.empty(n, dtype=np.int64)
    inv[piv] = np.arange(n)
    z = rng.standard_normal((n, n_draws))
    d, _ = ops.draws(Lp, np.zeros(n), Z=z)                       # rows in pivot order
    return np.asarray(mean, dtype=np.float64)[:, None] + d[inv]


def make_gaussian_partial_sums(X, orders=5, kernel=None, mean=None, ratio=0.3, ref=1., nugget=0, random_state=0,
                               allow_singular=True):
    """Partial sums y_k = ref * sum_{n in orders, n <= k} c_n ratio^n of GP coefficient curves c_n ~ N(mean(X), kernel(X) +
    nugget I) at the inputs X (gsum/datasets.py:8-72).

    X : (n_samples, n_features).  orders : int (orders 0 .. orders-1) or array of orders.  kernel : an sklearn kernel of the
    family the device builder evaluates ([Constant *] RBF [+ WhiteKernel]; default RBF(0.5)) — anything else raises
    NotImplementedError, there is no CPU fallback.  mean : callable X -> (n_samples,), default zero.  ratio, ref : scalar or
    callable of X.  Returns y with shape (n_samples, len(orders)).
    """
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be 2d: (n_samples, n_features)")
    if kernel is None:
        kernel = RBF(0.5)
    if isinstance(orders, (int, np.integer)):
        orders = np.arange(orders)
    orders = np.asarray(orders)
    if callable(ratio):
        ratio = ratio(X)
    if callable(ref):
        ref = ref(X)
    m = np.zeros(X.shape[0]) if mean is None else np.asarray(mean(X), dtype=np.float64)
    k = flatten_kernel(kernel)
    K = ops.kernel_matrix(X, None, k.ls_for(X.shape[1]), k.constant, k.noise)
    if nugget:
        K[np.diag_indices_from(K)] += nugget
    coeffs = _gaussian_draws(m, K, len(orders), check_random_state(random_state), allow_singular=allow_singular)
    return partials(coeffs=coeffs, ratio=ratio, ref=ref, orders=orders)


def make_gaussian_partial_sums_uniform(n_samples=100, n_features=1, orders=5, kernel=None, mean=None, ratio=0.3, ref=1.,
                                       nugget=0, random_state=0, allow_singular=True):
    """(X, y) with X uniform on [0, 1]^n_features (gsum/datasets.py:75-129).  As in the reference, X comes from a generator
    seeded with `random_state` and the coefficients from a second one seeded the same way."""
    X = check_random_state(random_state).rand(n_samples, n_features)
    y = make_gaussian_partial_sums(X=X, orders=orders, kernel=kernel, mean=mean, ratio=ratio, ref=ref, nugget=nugget,
                                   random_state=random_state, allow_singular=allow_singular)
    return X, y


def make_gaussian_partial_sums_on_grid(n_samples=100, n_features=1, orders=5, kernel=None, mean=None, ratio=0.3, ref=1.,
                                       nugget=0, random_state=0, allow_singular=True):
    """(X, y) with X the full grid linspace(0, 1, n_samples)^n_features, shape (n_samples ** n_features, n_features)
    (gsum/datasets.py:132-191).

    For n_features > 1 the reference's loop variable shadows the grid vector (datasets.py:182 builds `cartesian(0, 1, ...)`,
    a single point); the documented grid is what is generated here."""
    x = np.linspace(0, 1, n_samples)
    X = cartesian(*([x] * n_features)) if n_features > 1 else x[:, None]
    y = make_gaussian_partial_sums(X=X, orders=orders, kernel=kernel, mean=mean, ratio=ratio, ref=ref, nugget=nugget,
                                   random_state=random_state, allow_singular=allow_singular)
    return X, y


def generate_coefficients(X, size=1, basis=None, corr=None, beta=0, sd=1, noise=1e-5, **corr_kwargs):
    """`size` curves ~ N(basis(X) beta, sd^2 corr(X, **corr_kwargs) + noise^2 I), shape (size, n_samples) — the legacy generator
    of gsum/helpers.py:55-68.  `corr` is a callable returning the correlation matrix (default: `gaussian`, on the device); the
    factorisation and the draws run on the device, the normals come from numpy's global generator as in the reference."""
    X = np.asarray(X, dtype=np.float64)
    K = sd ** 2 * np.asarray((gaussian if corr is None else corr)(X, **corr_kwargs), dtype=np.float64)
    K[np.diag_indices_from(K)] += noise ** 2
    B = np.ones((len(X), 1)) if basis is None else basis(X)
    mean = np.dot(B, np.atleast_1d(beta))
    return _gaussian_draws(mean, K, size, np.random.mtrand._rand).T


def toy_data(X, orders, basis=None, corr=None, beta=0, sd=1, ratio=0.5, ref=1, noise=1e-5, **corr_kwargs):
    """Partial sums of len(orders) curves from `generate_coefficients` (gsum/helpers.py:36-52).  As in the reference the curves are
    the ROWS here — coefficients of shape (len(orders), n_samples) go to `partials` unchanged, which sums along the last axis
    with powers `orders` — so the call is meaningful for len(orders) == n_samples only; kept for import compatibility."""
    coeffs = generate_coefficients(X, size=len(orders), basis=basis, corr=corr, beta=beta, sd=sd, noise=noise, **corr_kwargs)
    return partials(coeffs=coeffs, ratio=ratio, ref=ref, orders=orders)
