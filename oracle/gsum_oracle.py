"""CPU oracle for the conjugate-GP likelihood / prediction / diagnostics path of buqeye/gsum.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``gsum_b200/`` imports it; the product path has no CPU fallback.

It is a numpy/scipy/scikit-learn restatement of the reference's algorithm for the hot path — the
same third-party calls (``numpy.linalg.cholesky``, ``scipy.linalg.cho_solve``, LAPACK ``dpstrf``,
sklearn kernel objects) in the same order as the reference, so that its results agree with the
reference's to the last bit or two — written as plain functions instead of the reference's
estimator classes.  Every function cites the reference ``file:line`` (relative to the reference
checkout, ``gsum/...``) it follows.

Pinning: the reference's own test-suite pins only ``pivoted_cholesky``
(``gsum/tests/test.py:75-122``) and the interpolation property (``test.py:63-72``).  Everything else
is pinned by (i) the stored notebook outputs reproduced in ``tests/golden/`` and (ii) outputs of the
reference itself, imported by path in the build container by ``tests/golden/make_golden.py`` and
frozen as fixtures (``tests/golden/*.npz``).  ``tests/test_oracle.py`` checks this file against all
of them.
"""
from __future__ import annotations

import numpy as np
from numpy.linalg import cholesky, solve
from scipy.linalg import cho_solve, eigh, inv, solve_triangular
from scipy.special import loggamma
import scipy.stats as st

__all__ = [
    "coefficients", "partials", "geometric_sum", "cartesian",
    "Priors", "solve_sqrt", "compute_center", "compute_disp", "compute_df", "compute_scale_sq",
    "compute_cov_factor", "gaussian_lml", "student_lml", "truncation_lml", "lml_grid",
    "fit_conjugate", "predict_conjugate", "predict_student", "truncation_cov", "truncation_mean",
    "truncation_basis", "predict_truncation",
    "pivoted_cholesky", "dpstrf_restated", "cholesky_errors", "mahalanobis", "md_squared",
    "pivoted_cholesky_errors", "credible_interval", "draws_from_z", "individual_errors", "eigen_factor", "eigen_errors",
]


# ----------------------------------------------------------------------------------------------
# Series helpers  (gsum/helpers.py)
# ----------------------------------------------------------------------------------------------

def cartesian(*arrays):
    """gsum/helpers.py:19-33 — cartesian product, earlier arrays vary slowest."""
    return np.stack(np.meshgrid(*arrays, indexing="ij"), -1).reshape(-1, len(arrays))


def coefficients(y, ratio, ref=1, orders=None):
    """gsum/helpers.py:71-101 — partial sums -> coefficients c_n = Δy_n / (ref · Q^n).

    A difference across skipped orders is attributed to the higher order (helpers.py:98-100).
    """
    if y.ndim != 2:
        raise ValueError("y must be 2d")
    if orders is None:
        orders = np.arange(y.shape[-1])
    if len(orders) != y.shape[-1]:
        raise ValueError("partials and orders must have the same length")
    ref, ratio, orders = np.atleast_1d(ref, ratio, orders)
    dy = np.insert(np.diff(y, axis=-1), 0, y[..., 0], axis=-1)
    return dy / (ref[:, None] * ratio[:, None] ** orders)


def partials(coeffs, ratio, ref=1, orders=None):
    """gsum/helpers.py:104-146 — coefficients -> partial sums y_k = ref Σ_{n<=k} c_n Q^n."""
    if orders is None:
        orders = np.arange(coeffs.shape[-1])
    ratio = np.atleast_1d(ratio)
    if ratio.ndim == 1:
        ratio = ratio[:, None]
    ref = np.atleast_1d(ref)
    if ref.ndim == 1:
        ref = ref[:, None]
    return np.cumsum(ref * coeffs * ratio ** orders, axis=-1)


def geometric_sum(x, start, end, excluded=None):
    """gsum/helpers.py:149-182 — Σ_{i=start}^{end} x^i minus excluded terms (x**inf -> 0 for |x|<1)."""
    if end < start:
        raise ValueError("end must be greater than or equal to start")
    s = (x ** start - x ** (end + 1)) / (1 - x)
    if excluded is not None:
        for n in np.atleast_1d(excluded):
            if start <= n <= end:
                s -= x ** n
    return s


# ----------------------------------------------------------------------------------------------
# Conjugate (normal-inverse-chi^2) updates  (gsum/models.py:169-503)
# ----------------------------------------------------------------------------------------------

class Priors:
    """The prior hyperparameters as the reference stores them (gsum/models.py:112-120)."""

    def __init__(self, center=0, disp=0, df=1, scale=1, sd=None):
        self.center0 = np.atleast_1d(center)
        self.disp0 = np.atleast_2d(disp)
        if sd is not None:
            self.df0, self.scale0 = np.inf, sd
        else:
            self.df0, self.scale0 = df, scale


def _ones_basis(X):
    """gsum/models.py:149-150 — default basis: a single constant column."""
    return np.ones((X.shape[0], 1))


def solve_sqrt(L, y):
    """gsum/models.py:459-487 — R^{-1} y from the lower factor (decomposition='cholesky', :479), or from the
    tuple (eig, Q) of `eigh(R)` as Q diag(1/eig) Qᵀ y (decomposition='eig', :480-484)."""
    if isinstance(L, tuple):
        eig, Q = L
        inv_mat = Q @ np.diag(1. / eig) @ Q.T                   # :483
        return inv_mat @ y
    return cho_solve((L, True), y)


def _sqrt_R(R, decomposition):
    """gsum/models.py:966-976 / 710-719: `cholesky(R)` or the tuple `eigh(R)`."""
    if decomposition == "cholesky":
        return cholesky(R)
    if decomposition == "eig":
        return eigh(R)
    raise ValueError('decomposition must be "cholesky" or "eig"')


def _num_y(y):
    return y.shape[1] if y.ndim == 2 else 1          # models.py:601-607


def _avg_y(y):
    return np.copy(y) if y.ndim == 1 else np.average(y, axis=1)   # models.py:609-628


def compute_disp(y, L, basis, disp0):
    """gsum/models.py:233-278 — V = (V0^{-1} + n_c BᵀR⁻¹B)^{-1}; stays 0 when V0 == 0."""
    if np.all(disp0 == 0):
        return np.zeros_like(disp0)
    quad = basis.T @ solve_sqrt(L, basis)
    return inv(inv(disp0) + _num_y(y) * quad)


def compute_center(y, L, basis, center0, disp0):
    """gsum/models.py:169-231 — η = V (V0^{-1} η0 + n_c BᵀR⁻¹ȳ); stays η0 when V0 == 0."""
    if np.all(disp0 == 0):
        return np.copy(center0)
    invR_y_avg = solve_sqrt(L, _avg_y(y))
    disp = compute_disp(y, L, basis, disp0)
    return disp @ (solve(disp0, center0) + _num_y(y) * basis.T @ invR_y_avg)


def compute_df(y, df0):
    """gsum/models.py:280-307 — ν = ν0 + N·n_c."""
    return df0 + y.size


def compute_scale_sq(y, L, basis, center0, disp0, df0, scale0):
    """gsum/models.py:386-457 — τ² = (ν0 τ0² + quad + quad2)/ν, incl. the dense N×N Woodbury `mat`."""
    if df0 == np.inf:
        return scale0 ** 2
    if y.ndim == 1:
        y = y[:, None]
    avg_y = _avg_y(y)
    N, ny = len(avg_y), _num_y(y)
    y_centered = y - avg_y[:, None]
    quad = np.trace(y_centered.T @ solve_sqrt(L, y_centered))
    avg_y_centered = avg_y - basis @ center0
    disp = compute_disp(y, L, basis, disp0)
    invR_basis = solve_sqrt(L, basis)
    invR_avg_yc = solve_sqrt(L, avg_y_centered)
    mat = np.eye(N) - ny * invR_basis @ disp @ basis.T
    quad2 = avg_y_centered @ (ny * mat @ invR_avg_yc)
    return (df0 * scale0 ** 2 + quad + quad2) / compute_df(y, df0)


def compute_cov_factor(scale_sq, df):
    """gsum/models.py:489-503 — σ² = ν τ²/(ν-2), or τ² when ν = inf."""
    return scale_sq if df == np.inf else df * scale_sq / (df - 2)


# ----------------------------------------------------------------------------------------------
# Marginal likelihoods  (gsum/models.py:912-1057, 1184-1273, 1485-1507)
# ----------------------------------------------------------------------------------------------

def gaussian_lml(kernel, theta, X, y, priors, nugget=1e-10, basis_fn=_ones_basis, decomposition="cholesky"):
    """gsum/models.py:912-1057 (eval_gradient=False; decomposition 'cholesky' or 'eig').

    Gaussian log-likelihood of the curves at the plug-in posterior-mean variance.  `kernel` is an
    sklearn kernel object; `theta` its log-hyperparameters (models.py:953).
    """
    R = kernel.clone_with_theta(theta)(X)                       # :953-960
    R[np.diag_indices_from(R)] += nugget                        # :963
    try:
        L_R = _sqrt_R(R, decomposition)                         # :969 / :974
    except np.linalg.LinAlgError:
        return -np.inf                                          # :970-972
    if y.ndim == 1:
        y = y[:, None]
    p = priors
    df = compute_df(y, p.df0)                                   # :986
    basis = basis_fn(X)
    center = compute_center(y, L_R, basis, p.center0, p.disp0)  # :1001
    scale2 = compute_scale_sq(y, L_R, basis, p.center0, p.disp0, p.df0, p.scale0)  # :1002
    mean = basis @ center
    var = compute_cov_factor(scale2, df)                        # :1008
    if decomposition == "cholesky":
        L = np.sqrt(var) * L_R                                  # :1014
        logdet_K = 2 * np.log(np.diag(L)).sum()                 # :1015
    else:
        eig, Q = L_R                                            # :1017
        L = var * eig, Q                                        # :1018
        logdet_K = np.log(var * eig).sum()                      # :1019
    _K = var * R                                                # :1023 (unused temp, kept for timing fidelity)
    y_train = y - mean[:, None]
    N = R.shape[0]
    alpha = solve_sqrt(L, y_train)                              # :1032
    ll = -0.5 * np.einsum("ik,ik->k", y_train, alpha)           # :1035
    ll -= 0.5 * logdet_K
    ll -= N / 2 * np.log(2 * np.pi)
    return ll.sum(-1)                                           # :1039


def gaussian_lml_gradient(kernel, theta, X, y, priors, nugget=1e-10, basis_fn=_ones_basis, decomposition="cholesky"):
    """gsum/models.py:912-1057 with eval_gradient=True (decomposition 'cholesky' or 'eig'): (log-likelihood, d/dtheta).

    The kernel gradient comes from sklearn (`kernel(X, eval_gradient=True)`, models.py:957-958); the conjugate updates
    carry their own derivatives: compute_center (models.py:222-230), compute_scale_sq (450-455), compute_cov_factor
    applied to d(scale^2) (997), and the chain rule of models.py:1024-1025, 1041-1056."""
    k = kernel.clone_with_theta(theta)
    R, dR = k(X, eval_gradient=True)                            # :957-958
    R[np.diag_indices_from(R)] += nugget                        # :963
    try:
        L_R = _sqrt_R(R, decomposition)                         # :969 / :974
    except np.linalg.LinAlgError:
        return -np.inf, np.zeros_like(theta)                    # :970-972
    if y.ndim == 1:
        y = y[:, None]
    p = priors
    df = compute_df(y, p.df0)
    basis = basis_fn(X)
    ny, N = _num_y(y), R.shape[0]
    # compute_center with gradient (:201-230)
    if np.all(p.disp0 == 0):
        center, grad_center = np.copy(p.center0), np.zeros((*p.center0.shape, dR.shape[-1]))
    else:
        center = compute_center(y, L_R, basis, p.center0, p.disp0)
        disp = compute_disp(y, L_R, basis, p.disp0)
        invR_basis = solve_sqrt(L_R, basis)
        invR_diff = solve_sqrt(L_R, basis @ center - _avg_y(y))
        grad_center = ny * disp @ np.einsum('ji,jkp,k->ip', invR_basis, dR, invR_diff)
    # compute_scale_sq with gradient (:419-455)
    scale2 = compute_scale_sq(y, L_R, basis, p.center0, p.disp0, p.df0, p.scale0)
    if p.df0 == np.inf:
        dscale2 = np.zeros(dR.shape[-1])
    else:
        avg_y = _avg_y(y)
        y_centered = y - avg_y[:, None]
        invR_yc = solve_sqrt(L_R, y_centered)
        avg_y_centered = avg_y - basis @ p.center0
        disp = compute_disp(y, L_R, basis, p.disp0)
        mat = np.eye(N) - ny * solve_sqrt(L_R, basis) @ disp @ basis.T
        mat_invR_avg_yc = ny * mat @ solve_sqrt(L_R, avg_y_centered)
        dscale2 = -np.einsum('ji,jkp,ki->p', invR_yc, dR, invR_yc)
        dscale2 -= np.einsum('i,ijp,j->p', mat_invR_avg_yc, dR, mat_invR_avg_yc) / ny
        dscale2 /= df
    grad_var = compute_cov_factor(dscale2, df)                  # :997
    grad_mean = basis @ grad_center                             # :999
    mean = basis @ center
    var = compute_cov_factor(scale2, df)
    if decomposition == "cholesky":
        L = np.sqrt(var) * L_R
        logdet_K = 2 * np.log(np.diag(L)).sum()
    else:
        L = var * L_R[0], L_R[1]                                # :1017-1018
        logdet_K = np.log(var * L_R[0]).sum()                   # :1019
    K_gradient = var * dR + grad_var * R[:, :, None]            # :1024-1025
    y_train = y - mean[:, None]
    alpha = solve_sqrt(L, y_train)
    ll = -0.5 * np.einsum("ik,ik->k", y_train, alpha) - 0.5 * logdet_K - N / 2 * np.log(2 * np.pi)
    tmp = np.einsum("ik,jk->ijk", alpha, alpha)                 # :1042
    tmp -= solve_sqrt(L, np.eye(N))[:, :, np.newaxis]           # :1045
    grad_dims = 0.5 * np.einsum("ijl,ijk->kl", tmp, K_gradient)  # :1049-1050
    grad_dims -= grad_mean.T @ alpha                            # :1053
    return ll.sum(-1), grad_dims.sum(-1)                        # :1056


def student_lml(kernel, theta, X, y, priors, nugget=1e-10, basis_fn=_ones_basis, decomposition="cholesky"):
    """gsum/models.py:1184-1273 (eval_gradient=False) — exact normal-inverse-χ² evidence."""
    ny = _num_y(y)
    R = kernel.clone_with_theta(theta)(X)
    R[np.diag_indices_from(R)] += nugget
    N = R.shape[0]
    try:
        L_R = _sqrt_R(R, decomposition)                          # :1211 / :1216
    except np.linalg.LinAlgError:
        return -np.inf
    p = priors
    df = compute_df(y, p.df0)
    basis = basis_fn(X)
    disp = compute_disp(y, L_R, basis, p.disp0)
    scale = np.sqrt(compute_scale_sq(y, L_R, basis, p.center0, p.disp0, p.df0, p.scale0))

    def log_norm(df_, scale_, disp_):                            # :1241-1247
        norm = loggamma(df_ / 2.0) - df_ / 2.0 * np.log(df_ * scale_ ** 2 / 2.0)
        log_det = np.linalg.slogdet(2 * np.pi * disp_)[1]
        if log_det != -np.inf:
            norm += 0.5 * log_det
        return norm

    if decomposition == "cholesky":
        logdet_R = 2 * np.log(np.diag(L_R)).sum()                # :1250
    else:
        logdet_R = np.log(L_R[0]).sum()                          # :1252-1253
    return log_norm(df, scale, disp) - log_norm(p.df0, p.scale0, p.disp0) \
        - ny / 2.0 * (N * np.log(2 * np.pi) + logdet_R)          # :1257-1258


def student_lml_gradient(kernel, theta, X, y, priors, nugget=1e-10, basis_fn=_ones_basis):
    """gsum/models.py:1184-1273 with eval_gradient=True: (evidence, d/dtheta).  compute_disp carries its derivative
    (models.py:270-277), compute_scale_sq as in the Gaussian case (450-455); chain rule at 1260-1271."""
    ny = _num_y(y)
    k = kernel.clone_with_theta(theta)
    R, dR = k(X, eval_gradient=True)
    R[np.diag_indices_from(R)] += nugget
    N = R.shape[0]
    try:
        L_R = cholesky(R)
    except np.linalg.LinAlgError:
        return -np.inf, np.zeros_like(theta)
    p = priors
    df = compute_df(y, p.df0)
    basis = basis_fn(X)
    disp = compute_disp(y, L_R, basis, p.disp0)
    if np.all(p.disp0 == 0):
        grad_disp = np.zeros((*p.disp0.shape, dR.shape[-1]))
    else:
        invRBV = solve_sqrt(L_R, basis) @ disp
        grad_disp = ny * np.einsum('ji,jkp,kl->ilp', invRBV, dR, invRBV)       # :275
    scale_sq = compute_scale_sq(y, L_R, basis, p.center0, p.disp0, p.df0, p.scale0)
    if p.df0 == np.inf:
        grad_scale_sq = np.zeros(dR.shape[-1])
    else:
        avg_y = _avg_y(y)
        y_centered = y - avg_y[:, None]
        invR_yc = solve_sqrt(L_R, y_centered)
        avg_y_centered = avg_y - basis @ p.center0
        mat = np.eye(N) - ny * solve_sqrt(L_R, basis) @ disp @ basis.T
        mat_invR_avg_yc = ny * mat @ solve_sqrt(L_R, avg_y_centered)
        grad_scale_sq = -np.einsum('ji,jkp,ki->p', invR_yc, dR, invR_yc)
        grad_scale_sq -= np.einsum('i,ijp,j->p', mat_invR_avg_yc, dR, mat_invR_avg_yc) / ny
        grad_scale_sq /= df
    scale = np.sqrt(scale_sq)

    def log_norm(df_, scale_, disp_):
        norm = loggamma(df_ / 2.0) - df_ / 2.0 * np.log(df_ * scale_ ** 2 / 2.0)
        log_det = np.linalg.slogdet(2 * np.pi * disp_)[1]
        if log_det != -np.inf:
            norm += 0.5 * log_det
        return norm

    logdet_R = 2 * np.log(np.diag(L_R)).sum()
    ll = log_norm(df, scale, disp) - log_norm(p.df0, p.scale0, p.disp0) - ny / 2.0 * (N * np.log(2 * np.pi) + logdet_R)
    grad = -(ny / 2.0) * np.trace(solve_sqrt(L_R, dR.reshape(N, -1)).reshape(dR.shape), axis1=0, axis2=1)     # :1262-1264
    grad -= (df / 2.0) * grad_scale_sq / scale_sq                # :1265
    if not np.all(disp == 0):
        grad += 0.5 * np.einsum('ij,ijp->p', inv(disp), grad_disp)   # :1268
    return ll, grad


def truncation_lml(kernel, theta, X, y, orders, ratio, ref, priors, nugget=1e-10, excluded=None,
                   student=False, decomposition="cholesky"):
    """gsum/models.py:1485-1507 — ll_y = ll_c(coefficients(y; Q, ref)) − Σ_x[n log|ref| + (Σ orders) log|Q|].

    `ratio` and `ref` are length-N arrays (the reference's `self.ratio(X, **ratio_kws)` / `self.ref(X)`).
    """
    orders = np.asarray(orders)
    mask = ~np.isin(orders, excluded)
    coeffs = coefficients(y=y, ratio=ratio, ref=ref, orders=orders)[:, mask]
    lml = student_lml if student else gaussian_lml
    ll = lml(kernel, theta, X, coeffs, priors, nugget=nugget, decomposition=decomposition)
    orders_in = orders[mask]
    det_factor = np.sum(len(orders_in) * np.log(np.abs(ref)) + np.sum(orders_in) * np.log(np.abs(ratio)))
    return ll - det_factor


def lml_grid(kernel, X, y, orders, ls_vals, ratio_vals, ref, priors, nugget=1e-10, excluded=None,
             student=False, ratio_fn=None):
    """docs/notebooks/correlated_EFT_publication.ipynb cell 53 — the reference's (Q, ℓ) double loop.

    Returns an (n_Q, n_ℓ) array, `[ratio][ls]` orientation.  `kernel` must have the length scale as
    its only free hyperparameter (theta = [log ℓ]).  `ratio_fn(X, q)` builds the ratio vector
    (defaults to a constant, gsum/models.py:1314-1315).
    """
    n = X.shape[0]
    ref_x = ref * np.ones(n) if np.ndim(ref) == 0 else np.asarray(ref, float)
    out = np.empty((len(ratio_vals), len(ls_vals)))
    for a, q in enumerate(ratio_vals):
        ratio_x = q * np.ones(n) if ratio_fn is None else ratio_fn(X, q)
        for b, ls in enumerate(ls_vals):
            out[a, b] = truncation_lml(kernel, np.log(np.atleast_1d(ls)), X, y, orders, ratio_x, ref_x,
                                       priors, nugget=nugget, excluded=excluded, student=student)
    return out


# ----------------------------------------------------------------------------------------------
# fit / predict  (gsum/models.py:671-738, 753-845, 1128-1182, 1337-1354, 1389-1483)
# ----------------------------------------------------------------------------------------------

def fit_conjugate(kernel, X, y, priors, nugget=1e-10, basis_fn=_ones_basis, student=False, decomposition="cholesky"):
    """gsum/models.py:671-738 with the kernel hyperparameters held fixed (optimizer=None / 'fixed').
    With decomposition='eig', `corr_L` is the tuple (eig, Q) the reference keeps as `_eigh_tuple_` (:714-716)."""
    X, y = X.copy(), y.copy()                                    # :692-701 (copy_X_train=True)
    corr = kernel(X)                                             # :708
    L = _sqrt_R(corr + nugget * np.eye(len(X)), decomposition)   # :711 / :714
    basis = basis_fn(X)
    p = priors
    center = compute_center(y, L, basis, p.center0, p.disp0)     # :721
    disp = compute_disp(y, L, basis, p.disp0)                    # :725
    df = compute_df(y, p.df0)                                    # :729
    scale_sq = compute_scale_sq(y, L, basis, p.center0, p.disp0, p.df0, p.scale0)   # :730
    lml = (student_lml if student else gaussian_lml)(kernel, kernel.theta, X, y, p, nugget, basis_fn, decomposition)  # :668-669
    return dict(kernel=kernel, X=X, y=y, basis=basis, corr=corr, corr_L=L, center=center, disp=disp, df=df,
                scale=np.sqrt(scale_sq), cov_factor=compute_cov_factor(scale_sq, df), lml=lml,
                nugget=nugget, basis_fn=basis_fn, decomposition=decomposition)


def predict_conjugate(f, Xnew, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False):
    """gsum/models.py:753-845 — GP posterior mean / std / cov at Xnew from a `fit_conjugate` dict."""
    if return_std and return_cov:
        raise RuntimeError("Only one of return_std or return_cov may be True")
    kern, nugget = f["kernel"], f["nugget"]
    if Xc is None:
        Xc, L = f["X"], f["corr_L"]
    else:
        L = _sqrt_R(kern(Xc) + nugget * np.eye(len(Xc)), f.get("decomposition", "cholesky"))   # :807-811
    if y is None:
        y = f["y"]
    m_old = f["basis_fn"](Xc) @ f["center"]                      # :818
    m_new = f["basis_fn"](Xnew) @ f["center"]
    R_on = kern(Xc, Xnew)                                        # :822
    R_no = R_on.T
    R_nn = kern(Xnew)                                            # :824 (one argument: WhiteKernel contributes)
    if y.ndim == 1:
        y = y[:, None]
    alpha = solve_sqrt(L, y - m_old[:, None])                    # :831
    m_pred = np.squeeze(m_new[:, None] + R_no @ alpha)           # :832
    if return_std or return_cov:
        R_pred = R_nn - R_no @ solve_sqrt(L, R_on)               # :836
        if pred_noise:
            R_pred += nugget * np.eye(len(Xnew))
        var = compute_cov_factor(f["scale"] ** 2, f["df"])       # :840
        K_pred = np.squeeze(var * R_pred)
        if return_std:
            return m_pred, np.sqrt(np.diag(K_pred))
        return m_pred, K_pred
    return m_pred


def predict_student(f, Xnew, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False):
    """gsum/models.py:1128-1182 — adds the mean-uncertainty term σ² B̃ V B̃ᵀ (std added linearly, :1176)."""
    pred = predict_conjugate(f, Xnew, return_std, return_cov, Xc, y, pred_noise)
    kern, nugget = f["kernel"], f["nugget"]
    basis_new = f["basis_fn"](Xnew)
    if Xc is None:
        basis_old, L, R_no = f["basis"], f["corr_L"], kern(Xnew, f["X"])
    else:
        basis_old, R_no = f["basis_fn"](Xc), kern(Xnew, Xc)
        L = _sqrt_R(kern(Xc) + nugget * np.eye(len(Xc)), f.get("decomposition", "cholesky"))   # :1163-1166
    basis = basis_new - R_no @ solve_sqrt(L, basis_old)          # :1171
    mean_cov = f["cov_factor"] * (basis @ f["disp"] @ basis.T)   # :1173
    if return_std:
        return pred[0], pred[1] + np.sqrt(np.diag(mean_cov))
    if return_cov:
        return pred[0], pred[1] + mean_cov
    return pred


def truncation_mean(f, X, ratio_fn, ref_fn, start=0, end=np.inf, excluded=None):
    """gsum/models.py:1337-1340."""
    coeff_mean = f["basis_fn"](X) @ f["center"]
    return ref_fn(X) * geometric_sum(ratio_fn(X), start, end, excluded) * coeff_mean


def truncation_cov(f, X, Xp, ratio_fn, ref_fn, start=0, end=np.inf, excluded=None):
    """gsum/models.py:1342-1348 + 562-599 — ref⊗ref ∘ gs(Q⊗Q) ∘ σ² kernel(X, Xp).

    `Xp` is passed through to the kernel as-is, so a WhiteKernel contributes only when Xp is None.
    """
    coeff_cov = f["cov_factor"] * f["kernel"](X, Xp)
    Xp = X if Xp is None else Xp
    ratio_mat = ratio_fn(X)[:, None] * ratio_fn(Xp)
    ref_mat = ref_fn(X)[:, None] * ref_fn(Xp)
    return ref_mat * geometric_sum(ratio_mat, start, end, excluded) * coeff_cov


def truncation_basis(f, X, ratio_fn, ref_fn, start=0, end=np.inf, excluded=None):
    """gsum/models.py:1350-1354."""
    return ref_fn(X)[:, None] * geometric_sum(ratio_fn(X)[:, None], start, end, excluded) * f["basis_fn"](X)


def predict_truncation(f, Xnew, order, y_order, ratio_fn, ref_fn, return_std=False, return_cov=False,
                       Xc=None, kind="both", excluded=None, dX=None, dy=None):
    """gsum/models.py:1389-1483 — interpolation ⊕ truncation-error process; LU solves, no nugget in K_oo.

    `f` is the `fit_conjugate` dict of the coefficient process; `y_order` the partial sum at `order`
    on the conditioning points (models.py:1422-1428).
    """
    if Xc is None:
        Xc = f["X"]
    if kind not in ("both", "interp", "trunc"):
        raise ValueError('kind must be one of "both", "interp" or "trunc"')
    kw = dict(ratio_fn=ratio_fn, ref_fn=ref_fn, excluded=excluded)
    m_pred, K_pred = 0, 0
    if kind in ("both", "interp"):
        m_old = truncation_mean(f, Xc, start=0, end=order, **kw)
        m_new = truncation_mean(f, Xnew, start=0, end=order, **kw)
        K_oo = truncation_cov(f, Xc, Xc, start=0, end=order, **kw)
        K_on = truncation_cov(f, Xc, Xnew, start=0, end=order, **kw)
        K_nn = truncation_cov(f, Xnew, Xnew, start=0, end=order, **kw)
        m_pred = m_pred + m_new + K_on.T @ solve(K_oo, y_order - m_old)      # :1449-1450
        if return_std or return_cov:
            K_pred = K_pred + K_nn - K_on.T @ solve(K_oo, K_on)              # :1452
    if kind in ("both", "trunc"):
        m_new_t = truncation_mean(f, Xnew, start=order + 1, end=np.inf, **kw)
        K_nn_t = truncation_cov(f, Xnew, Xnew, start=order + 1, end=np.inf, **kw)
        if dX is not None:                                                    # :1464-1473
            m_old_t = truncation_mean(f, dX, start=order + 1, end=np.inf, **kw)
            K_oo_t = truncation_cov(f, dX, dX, start=order + 1, end=np.inf, **kw)
            K_on_t = truncation_cov(f, dX, Xnew, start=order + 1, end=np.inf, **kw)
            m_pred = m_pred + m_new_t + K_on_t.T @ solve(K_oo_t, dy - m_old_t)
            if return_std or return_cov:
                K_pred = K_pred + K_nn_t - K_on_t.T @ solve(K_oo_t, K_on_t)
        else:
            m_pred = m_pred + m_new_t
            if return_std or return_cov:
                K_pred = K_pred + K_nn_t
    if return_cov:
        return m_pred, K_pred
    if return_std:
        return m_pred, np.sqrt(np.diag(K_pred))
    return m_pred


# ----------------------------------------------------------------------------------------------
# Diagnostics  (gsum/helpers.py:185-199, 504-522; gsum/diagnostics.py:38-171)
# ----------------------------------------------------------------------------------------------

def pivoted_cholesky(M, return_pivots=False):
    """gsum/helpers.py:185-199 — LAPACK dpstrf(lower) -> G = L[p_inv] with M = G Gᵀ."""
    from scipy.linalg.lapack import get_lapack_funcs
    (pstrf,) = get_lapack_funcs(("pstrf",), arrays=(M,))
    c, p, _, info = pstrf(M, lower=True)
    if info > 0:
        raise np.linalg.LinAlgError("M is not positive-semidefinite")
    if info < 0:
        raise ValueError("LAPACK reported an illegal value in {}-th argument on entry to pstrf".format(-info))
    L = np.tril(c)
    p = p - 1
    p_inv = np.arange(len(p))[np.argsort(p)]
    G = L[p_inv]
    return (G, p.astype(np.int32)) if return_pivots else G


def dpstrf_restated(M, nb=64):
    """LAPACK 3.x ``dpstrf``/``dpstf2`` (lower), restated in numpy — the *published algorithm* of the
    third-party routine behind gsum/helpers.py:187-188 (scipy's bundled LAPACK; ILAENV gives NB=64 for
    xPOTRF).  Returns (L lower-triangular, piv 0-based, rank, info): Pᵀ M P = L Lᵀ with P = I[:, piv].

    Mirrors the routine's arithmetic *order* for the running diagonal (the `work`/`dot2` array is reset
    at every block start and accumulates the squares of the current block's columns only; pivot =
    first maximum of ``A_ii − work_i``; stop when that maximum ≤ N·eps·max_i A_ii), which is what the
    device kernel in gsum_b200/csrc follows too, so pivot orders can be compared bit for bit.
    """
    A = np.array(M, dtype=float, order="C", copy=True)
    n = A.shape[0]
    piv = np.arange(n)
    ajj = np.max(np.diag(A))
    if n == 0:
        return A, piv.astype(np.int32), 0, 0
    if ajj <= 0 or np.isnan(ajj):
        return np.tril(A), piv.astype(np.int32), 0, 1
    dstop = n * (0.5 * np.finfo(float).eps) * ajj     # DLAMCH('Epsilon') = 2^-53 with rounding
    work = np.zeros(n)
    rank, info = n, 0
    k = 0
    if nb <= 1 or nb >= n:
        nb = n                                     # unblocked dpstf2 == one block spanning everything
    while k < n:
        jb = min(nb, n - k)
        work[k:] = 0.0
        for j in range(k, k + jb):
            if j > k:
                work[j:] += A[j:, j - 1] ** 2
            cand = np.diag(A)[j:] - work[j:]
            if j > 0:
                pvt = j + int(np.argmax(cand))     # MAXLOC: first maximum
                ajj = cand[pvt - j]
                if ajj <= dstop or np.isnan(ajj):
                    A[j, j] = ajj
                    rank, info = j, 1
                    return np.tril(A), piv.astype(np.int32), rank, info
            else:
                pvt = int(np.argmax(np.diag(A)))
                ajj = A[pvt, pvt]
            if pvt != j:                           # symmetric interchange on the lower triangle
                A[pvt, pvt] = A[j, j]
                A[[j, pvt], :j] = A[[pvt, j], :j]
                if pvt < n - 1:
                    A[pvt + 1:, [j, pvt]] = A[pvt + 1:, [pvt, j]]
                tmp = A[j + 1:pvt, j].copy()
                A[j + 1:pvt, j] = A[pvt, j + 1:pvt]
                A[pvt, j + 1:pvt] = tmp
                work[[j, pvt]] = work[[pvt, j]]
                piv[[j, pvt]] = piv[[pvt, j]]
            ajj = np.sqrt(ajj)
            A[j, j] = ajj
            if j < n - 1:
                A[j + 1:, j] -= A[j + 1:, k:j] @ A[j, k:j]
                A[j + 1:, j] /= ajj
        if k + jb < n:
            P = A[k + jb:, k:k + jb]
            A[k + jb:, k + jb:] -= P @ P.T          # dsyrk on the trailing block (lower half is what matters)
        k += jb
    return np.tril(A), piv.astype(np.int32), rank, info


def cholesky_errors(y, mean, chol):
    """gsum/helpers.py:504-505 — L^{-1}(y − m); y is (n_curves, N)."""
    return solve_triangular(chol, (y - mean).T, lower=True).T


def mahalanobis(y, mean, chol):
    """gsum/helpers.py:512-517 (chol branch)."""
    return np.linalg.norm(cholesky_errors(y, mean, chol), axis=-1)


def md_squared(y, mean, chol):
    """gsum/diagnostics.py:112-114 — y is (N, n_curves) as in the Diagnostic API."""
    return mahalanobis(y.T, mean, chol) ** 2


def pivoted_cholesky_errors(y, mean, G):
    """gsum/diagnostics.py:103-104 — dense solve with the row-permuted factor G."""
    return solve(G, (y.T - mean).T)


def individual_errors(y, mean, cov):
    """gsum/diagnostics.py:84-98."""
    return ((y.T - mean) / np.sqrt(np.diag(cov))).T


def credible_interval(y, mean, cov, intervals, df=None):
    """gsum/diagnostics.py:148-171 — fraction of points inside each central interval; the marginals are
    norm(mean, sd) for df=None and t(df, loc=mean, scale=sd) otherwise (gsum/diagnostics.py:48, 54)."""
    sd = np.sqrt(np.diag(cov))
    udist = st.norm(loc=mean, scale=sd) if df is None else st.t(loc=mean, scale=sd, df=df)
    lower, upper = udist.interval(np.atleast_2d(intervals).T)

    def diagnostic(data_, lower_, upper_):
        return np.average((lower_ < data_) & (data_ < upper_), axis=1)

    dci = np.apply_along_axis(diagnostic, axis=1, arr=np.atleast_2d(y).T, lower_=lower, upper_=upper)
    return np.squeeze(dci) if y.ndim == 1 else dci


def draws_from_z(mean, chol, z):
    """Deterministic half of gsum/diagnostics.py:82 / gsum/models.py:872: m + L z for caller-supplied
    standard-normal z (N, n).  (The reference draws through numpy's SVD-based legacy sampler, whose
    stream cannot be reproduced off-host; the distribution of m + L z is identical.)"""
    return mean[:, None] + chol @ z


def mvt_draws_from_z(mean, cov, df, z, x):
    """Student-t branch of gsum/diagnostics.py:51-55, 82: `MVT(mean, sigma = cov (df - 2) / df, df).rvs`.  MVT is
    statsmodels' (not installed here; no version pinned by the reference): its published sampler
    `multivariate_t_rvs` returns m + z_sigma / sqrt(x) with z_sigma ~ N(0, sigma) and x = chi2_df / df.  Deterministic
    half for caller-supplied standard normals z (N, n) and x (n,): z_sigma = chol(sigma) z."""
    sigma = cov * (df - 2.0) / df
    return mean[:, None] + (np.linalg.cholesky(sigma) @ z) / np.sqrt(x)[None, :]


def eigen_factor(cov):
    """gsum/diagnostics.py:63-68 — `_eig` = Q diag(sqrt(eig)) with the eigenvalues ordered from largest to smallest."""
    e, v = np.linalg.eigh(cov)
    e, v = e[::-1], v[:, ::-1]
    return v @ np.diag(np.sqrt(e))


def eigen_errors(y, mean, eig_factor):
    """gsum/diagnostics.py:106-107 — solve(_eig, (y.T - mean).T); y (N, n_curves)."""
    return solve(eig_factor, (y.T - mean).T)


def kl_divergence(mean1, cov1, chol1, mean0, cov0):
    """gsum/diagnostics.py:116-146 as written: note `logs` uses diag(c1), the covariance's own diagonal (the reference's
    expression), not the diagonal of its Cholesky factor."""
    tr = np.trace(cho_solve((chol1, True), cov0))
    dist = md_squared(mean0, mean1, chol1)
    k = cov1.shape[-1]
    logs = 2 * np.sum(np.log(np.diag(cov1))) - np.linalg.slogdet(cov0)[-1]
    return 0.5 * (tr + dist - k + logs)


# ----------------------------------------------------------------------------------------------
# TruncationPointwise (gsum/models.py:1573-1836), VariogramFourthRoot (gsum/helpers.py:525-730)
# Pinned on tests/golden/pointwise_variogram.npz (outputs of the real classes; make_golden_pointwise.py).
# ----------------------------------------------------------------------------------------------
def pointwise_fit(y, ratio, ref, orders, df0=1, scale0=1, excluded=None):
    """gsum/models.py:1643-1683 — coefficients of the unmasked orders, nu = nu0 + n_orders (:1624-1626), tau (:1628-1632), and the
    Student-t truncation-error distribution per (point, order): loc = y_k, scale = ref sqrt(sum_{n>k} Q^2n) tau (:1676-1680)."""
    y = np.asarray(y, dtype=float)
    if y.ndim == 1:
        y = y[:, None]
    ratio, ref = np.atleast_1d(ratio, ref)
    orders = np.arange(y.shape[-1]) if orders is None else np.asarray(orders)
    if y.shape[-1] != orders.size:
        raise ValueError('The last dimension of `y` must have the same size as `orders`')
    mask = ~np.isin(orders, excluded)
    c = coefficients(y, ratio, ref, orders)[:, mask]
    df = df0 + c.shape[-1]
    scale = np.sqrt((df0 * scale0 ** 2 + (c ** 2).sum(-1)) / df)
    om = orders[mask]
    ratio_sums = np.array([geometric_sum(ratio ** 2, k + 1, np.inf, excluded=excluded) for k in om]).T
    trunc_scale = ref[:, None] * np.sqrt(ratio_sums) * scale[:, None]
    return dict(y=y, ratio=ratio, ref=ref, orders=orders, mask=mask, orders_masked=om, coeffs=c, df=df, scale=scale,
                trunc_scale=trunc_scale, dist=st.t(loc=y[:, mask], scale=trunc_scale, df=df), df0=df0, scale0=scale0)


def _pointwise_idx(f, orders):
    """gsum/models.py:1640-1644 — positions of `orders` among the unmasked orders (squeezed, as the reference does)."""
    if orders is None:
        return slice(None)
    return np.squeeze([np.nonzero(f["orders_masked"] == o) for o in np.atleast_1d(orders)])


def pointwise_interval(f, alpha, orders=None):
    """gsum/models.py:1685-1707."""
    alpha = np.array(alpha)
    if alpha.ndim == 1:
        alpha = alpha[:, None, None]
    return np.array(f["dist"].interval(alpha))[..., _pointwise_idx(f, orders)]


def pointwise_pdf(f, y, orders=None, log=False):
    """gsum/models.py:1709-1741 (pdf / logpdf)."""
    y = np.atleast_1d(y)
    if y.ndim == 1:
        y = y[:, None, None]
    d = f["dist"]
    return (d.logpdf(y) if log else d.pdf(y))[..., _pointwise_idx(f, orders)]


def pointwise_log_likelihood(f, ratio=None, ref=None):
    """gsum/models.py:1748-1793 as written: the Gamma-function and 2 pi terms enter ONCE (not once per point), the tau terms are summed
    over the points, and the change-of-variables term is summed over the broadcast shape of `ref` and `ratio`."""
    ratio = f["ratio"] if ratio is None else ratio
    ref = f["ref"] if ref is None else ref
    c = coefficients(f["y"], ratio, ref, f["orders"])[:, f["mask"]]
    df0, scale0 = f["df0"], f["scale0"]
    df = df0 + c.shape[-1]
    scale = np.sqrt((df0 * scale0 ** 2 + (c ** 2).sum(-1)) / df)
    n = c.shape[-1]
    ll = loggamma(df / 2.) - 0.5 * n * np.log(2 * np.pi)
    if df0 > 0:
        ll += 0.5 * np.sum(df0 * np.log(df0 * scale0 ** 2 / 2.)) - loggamma(df0 / 2.)
    ll -= 0.5 * np.sum(df * np.log(df * scale ** 2 / 2.))
    ll -= np.sum(np.log(np.abs(ref)) + np.sum(f["orders"][f["mask"]]) * np.log(ratio))
    return ll


def pointwise_credible_diagnostic(f, data, dobs):
    """gsum/models.py:1795-1810 — fraction of points inside each central interval, per order."""
    dobs = np.atleast_1d(dobs)
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    lower, upper = f["dist"].interval(dobs[:, None, None])
    return np.average((lower < data) & (data < upper), axis=1)


class VariogramOracle:
    """gsum/helpers.py:525-730 restated with plain arrays instead of record arrays: pairs (i > j) in `np.tril_indices` order,
    distance bins by `np.digitize`, per-bin means of sqrt|z_i - z_j| (gamma*_hat), the fourth-root transform
    gamma~ = (gamma*_hat / mean_factor)^4, and the covariance of two bin means from the correlation of sqrt-differences
    (Cressie & Hawkins: (1 - rho^2) 2F1(3/4, 3/4; 1/2; rho^2) - 1, clipped at |rho| >= 1)."""
    from scipy.special import gamma as _g
    mean_factor = np.sqrt(2 / np.pi) * _g(0.75)
    var_factor = 2. / np.pi * (np.sqrt(np.pi) - _g(0.75) ** 2)
    corr_factor = _g(0.75) ** 2 / (np.sqrt(np.pi) - _g(0.75) ** 2)

    def __init__(self, X, z, bin_bounds):                                            # :546-611
        X = np.asarray(X, dtype=float)
        bin_bounds = np.asarray(bin_bounds, dtype=float)
        N = len(X)
        hij = np.linalg.norm(X[:, None, :] - X, axis=-1)
        self.bin_grid = np.digitize(hij, bin_bounds)
        z = np.atleast_2d(z)
        self.Ncurves = z.shape[0]
        dij = np.sqrt(np.abs(z.T[:, None, :] - z.T[None, :, :]))
        ti, tj = np.tril_indices(N, -1)
        self.i, self.j, self.hij, self.dij = ti, tj, hij[ti, tj], dij[ti, tj]
        self.Nb = Nb = len(bin_bounds) + 1
        self.gamma_star_hat = np.full((Nb, self.Ncurves), np.nan)
        loc = np.zeros(Nb)
        loc[1:-1] = (bin_bounds[1:] + bin_bounds[:-1]) / 2
        loc[0] = 2 * bin_bounds[0] - loc[1]
        loc[-1] = 2 * bin_bounds[-1] - loc[-2]
        self.bin_idx = np.digitize(self.hij, bin_bounds)
        self.bin_mask = np.arange(Nb)[:, None] == self.bin_idx
        self.bin_counts = self.bin_mask.sum(-1)
        for b, m in enumerate(self.bin_mask):
            if np.any(m):
                loc[b] = np.average(self.hij[m], axis=0)
                self.gamma_star_hat[b] = np.average(self.dij[m], axis=0)
        self.bin_locations = loc
        self.gamma_tilde = (self.gamma_star_hat / self.mean_factor) ** 4
        self.gamma_tilde_grid = self.gamma_tilde[self.bin_grid]
        self.gamma_star_mean = self.mean_factor * self.gamma_star_hat

    def rho_ijkl(self, i, j, k, l):                                                  # :613-623
        g = self.gamma_tilde_grid
        return (g[j, k] + g[i, l] - g[i, k] - g[j, l]) / (2 * np.sqrt(g[i, j] * g[k, l]))

    def corr_ijkl(self, i, j, k, l):                                                 # :625-637
        from scipy.special import hyp2f1
        rho = self.rho_ijkl(i, j, k, l)
        corr = ((1 - rho ** 2) * hyp2f1(0.75, 0.75, 0.5, rho ** 2) - 1) * self.corr_factor
        corr[rho >= 1.] = 1.
        corr[rho <= -1.] = -1.
        return corr

    def var_ij(self, i, j):                                                          # :652-654
        return self.var_factor * np.sqrt(self.gamma_tilde_grid[i, j])

    def cov_ijkl(self, i, j, k, l):                                                  # :639-650
        i, j, k, l = np.atleast_1d(i, j, k, l)
        n = i.shape[0], self.Ncurves
        corr = np.where((i == k) & (j == l), np.ones(n).T, self.corr_ijkl(i, j, k, l).T).T
        return corr * np.sqrt(self.var_ij(i, j) * self.var_ij(k, l))

    def cov(self, bin1, bin2=None):                                                  # :656-681
        m1 = self.bin_mask[bin1]
        m2 = m1 if (bin2 is None or bin2 == bin1) else self.bin_mask[bin2]
        nb1, nb2 = m1.sum(), m2.sum()
        if nb1 * nb2 == 0:
            return 0.
        a, b = np.nonzero(m1)[0], np.nonzero(m2)[0]
        A, B = np.repeat(a, len(b)), np.tile(b, len(a))                              # `cartesian` of the two pair lists
        return np.sum(self.cov_ijkl(self.i[A], self.j[A], self.i[B], self.j[B]), axis=0) / (nb1 * nb2)

    def compute(self, rt_scale=False):                                               # :689-715
        gam = self.gamma_star_mean if rt_scale else self.gamma_tilde
        sd = np.zeros((self.Nb, self.Ncurves))
        for b in range(self.Nb):
            sd[b] = np.sqrt(self.cov(b))
        lower, upper = self.gamma_star_mean - sd, self.gamma_star_mean + sd
        if not rt_scale:
            lower, upper = (lower / self.mean_factor) ** 4, (upper / self.mean_factor) ** 4
        return gam, lower, upper


# ----------------------------------------------------------------------------------------------
# Free helpers either side of the path: correlation functions, Gaussian KL, pdf summaries, data generator
# (gsum/helpers.py:202-368, gsum/datasets.py:8-72)
# ----------------------------------------------------------------------------------------------

def gaussian_corr(X, Xp=None, ls=1):
    """gsum/helpers.py:233-251 — exp(-sqd/2) with sqd from the expanded square |x|^2 + |x'|^2 - 2 x.x', clipped at 0."""
    X = X * 1.0 / ls
    Xp = X if Xp is None else Xp                 # NB (helpers.py:244-246): Xp is NOT rescaled by ls when given
    sqd = -2.0 * np.dot(X, Xp.T) + (np.sum(X ** 2, axis=1)[:, None] + np.sum(Xp ** 2, axis=1)[None, :])
    return np.exp(-0.5 * np.clip(sqd, 0.0, np.inf))


def rbf_corr(X, Xp=None, ls=1):
    """gsum/helpers.py:254-261 — exp(-|x - x'|^2 / (2 ls^2)); the indicator of coinciding points for ls == 0."""
    Xp = X if Xp is None else Xp
    dist = np.linalg.norm(X[:, None, ...] - Xp[None, ...], axis=-1)
    if ls == 0:
        return np.where(dist == 0, 1., 0.)
    return np.exp(-0.5 * dist ** 2 / ls ** 2)


def kl_gauss(mu0, cov0, mu1, cov1=None, chol1=None):
    """gsum/helpers.py:310-368 — KL(N0 || N1); cov1 is factored after adding 1e-5 I (helpers.py:202-203, 357)."""
    mu0, mu1 = np.atleast_1d(mu0), np.atleast_1d(mu1)
    cov0 = np.atleast_2d(cov0)
    if (cov1 is None) == (chol1 is None):
        raise ValueError('Exactly one of cov1 or chol1 must be given.')
    if chol1 is None:
        cov1 = np.atleast_2d(cov1)
        chol1 = cholesky(cov1 + 1e-5 * np.eye(*cov1.shape))
    chol1 = np.atleast_2d(chol1)
    k = cov0.shape[0]
    logdet0 = np.linalg.slogdet(cov0)[1]
    logdet1 = 2 * np.sum(np.log(np.diag(chol1)))
    rq = solve(chol1, mu1 - mu0)
    return 0.5 * (np.trace(cho_solve((chol1, True), cov0)) + rq @ rq - k + logdet1 - logdet0)


def hpd(dist, alpha):
    """gsum/helpers.py:264-278 — narrowest CDF window of mass alpha (Nelder-Mead from 1 - alpha, ftol 1e-8)."""
    from scipy.optimize import fmin
    start = fmin(lambda s: dist.ppf(s + alpha) - dist.ppf(s), 1 - alpha, ftol=1e-8, disp=False)[0]
    return dist.ppf([start, alpha + start])


def hpd_pdf(pdf, alpha, x):
    """gsum/helpers.py:281-295."""
    heights = np.unique(pdf)
    errs = np.array([(np.trapezoid(pdf[pdf >= p], x=x[pdf >= p]) - alpha) ** 2 for p in heights])
    interval = np.asarray(x)[pdf > heights[np.argmin(errs)]]
    return np.array([np.min(interval), np.max(interval)])


def median_pdf(pdf, x):
    """gsum/helpers.py:298-307."""
    i = 0
    for i in range(len(x)):
        if np.trapezoid(pdf[:i + 1], x[:i + 1]) > 0.5:
            break
    return x[i]


def predictions(dist, dob=None):
    """gsum/helpers.py:206-230."""
    mean = dist.mean()
    if dob is None:
        return mean
    return mean, np.squeeze(np.asarray(dist.interval(np.atleast_2d(dob).T)).transpose((1, 0, 2)))


def gaussian_partial_sums_cov(kernel, X, nugget=0):
    """gsum/datasets.py:64-66 — the covariance the coefficient curves are drawn from: kernel(X) + nugget I."""
    K = kernel(X)
    return K + nugget * np.eye(K.shape[0])
