import numpy as np


def relerr(a, b):
    """max |a - b| / max |b| (array-level relative error)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def prior_kwargs(arr):
    return dict(center=float(arr[0]), disp=float(arr[1]), df=float(arr[2]), scale=float(arr[3]))


def c4_inputs(n=1024, n_orders=6, seed=3, ls_true=0.05, noise=1e-6):
    """X = linspace(0,1,N), coefficients ~ GP(RBF(0.05) + 1e-6 I), Q = 0.5, ref = 1."""
    from scipy import stats
    from sklearn.gaussian_process.kernels import RBF
    X = np.linspace(0, 1, n)[:, None]
    K = RBF(ls_true)(X) + noise * np.eye(n)
    coeffs = stats.multivariate_normal(np.zeros(n), K, allow_singular=True).rvs(n_orders, random_state=seed).T
    orders = np.arange(n_orders)
    y = np.cumsum(coeffs * 0.5 ** orders, axis=-1)
    return X, y, orders


def lml_extended_precision(X, coeffs, ls, noise, nugget, center0, disp0, df0, scale0, constant=1.0):
    """Gaussian conjugate log-likelihood (gsum/models.py:912-1057, disp/df branches as in SURVEY Appendix B) evaluated in
    x87 extended precision (np.longdouble, eps ~ 1e-19) with an unblocked Cholesky: the arbiter for ill-conditioned cells,
    where two backward-stable FP64 algorithms (LAPACK in the reference, the tiled device factorisation) legitimately differ
    by ~cond(R) * eps."""
    ld = np.longdouble
    Xs = (np.asarray(X, dtype=ld) / np.asarray(ls, dtype=ld))
    d2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    R = ld(constant) * np.exp(ld(-0.5) * d2)
    n = R.shape[0]
    R[np.diag_indices(n)] = ld(constant) + ld(noise) + ld(nugget)
    L = np.zeros_like(R)
    for j in range(n):
        d = R[j, j] - (L[j, :j] ** 2).sum()
        L[j, j] = np.sqrt(d)
        L[j + 1:, j] = (R[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    C = np.asarray(coeffs, dtype=ld)
    W = np.zeros((n, C.shape[1] + 1), dtype=ld)
    B = np.concatenate([np.ones((n, 1), dtype=ld), C], axis=1)
    for i in range(n):
        W[i] = (B[i] - L[i, :i] @ W[:i]) / L[i, i]
    G = W.T @ W
    nc = C.shape[1]
    bb, h, Gc = G[0, 0], G[1:, 0], G[1:, 1:]
    trG, s11, hs = np.trace(Gc), Gc.sum(), h.sum()
    yRy, BRy = s11 / nc ** 2, hs / nc
    eta0, V0, tau0sq = ld(center0), ld(disp0), ld(scale0) ** 2
    df = df0 + n * nc
    V, eta = ld(0), eta0
    if V0 != 0:
        V = 1 / (1 / V0 + nc * bb)
        eta = V * (eta0 / V0 + nc * BRy)
    quad = trG - nc * yRy
    aRa = yRy - 2 * eta0 * BRy + eta0 ** 2 * bb
    BRa = BRy - bb * eta0
    quad2 = nc * (aRa - nc * BRa ** 2 * V)
    tausq = tau0sq if np.isinf(df0) else (ld(df0) * tau0sq + quad + quad2) / ld(df)
    var = tausq if np.isinf(df) else ld(df) * tausq / (ld(df) - 2)
    logdet = 2 * np.log(np.diag(L)).sum()
    Seta = quad + nc * (yRy - 2 * eta * BRy + eta ** 2 * bb)
    ll = -Seta / (2 * var) - ld(nc) / 2 * (n * np.log(var) + logdet) - ld(nc) * n / 2 * np.log(2 * ld(np.pi))
    return float(ll)


# ---- extended-precision arbiter for posteriors and predictions -------------------------------------------------------
# Two backward-stable FP64 routes to the same quantity (LAPACK in the reference, the tiled device factorisation) differ by
# ~cond * eps, and a predictive variance is a difference that cancels.  Where such a comparison exceeds rtol 1e-10 the test
# does not widen the bound: it evaluates the reference's formulas in x87 extended precision (eps ~ 1e-19) and requires the
# device result to be as close to that value as the reference's own FP64 result is (`as_close_as_reference`).
_ld = np.longdouble


def ld_cholesky(A):
    A = np.asarray(A, dtype=_ld)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        L[j, j] = np.sqrt(A[j, j] - (L[j, :j] ** 2).sum())
        L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def ld_cho_solve(L, B):
    """R^-1 B from the lower factor, by forward and back substitution in extended precision."""
    B = np.asarray(B, dtype=_ld)
    one_d = B.ndim == 1
    W = np.array(B[:, None] if one_d else B, dtype=_ld)
    n = L.shape[0]
    for i in range(n):
        W[i] = (W[i] - L[i, :i] @ W[:i]) / L[i, i]
    for i in range(n - 1, -1, -1):
        W[i] = (W[i] - L[i + 1:, i] @ W[i + 1:]) / L[i, i]
    return W[:, 0] if one_d else W


def ld_rbf(X1, X2, ls, constant=1.0):
    a = np.asarray(X1, dtype=_ld) / np.asarray(ls, dtype=_ld)
    b = np.asarray(X2, dtype=_ld) / np.asarray(ls, dtype=_ld)
    d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    return _ld(constant) * np.exp(_ld(-0.5) * d2)


def conjugate_extended_precision(X, y, ls, constant, noise, nugget, center0, disp0, df0, scale0, Xn, student=False,
                                 Xc=None, yc=None, pred_noise=False):
    """fit (gsum/models.py:671-738) and predict (:753-845; Student-t: :1128-1182) of the conjugate process with the default
    constant basis, in extended precision.  Returns dict(center, disp, df, scale, cov_factor, mean, std, cov)."""
    X, Xn = np.asarray(X, dtype=float), np.asarray(Xn, dtype=float)
    y = np.asarray(y, dtype=_ld)
    if y.ndim == 1:
        y = y[:, None]
    n, nc = y.shape
    one = np.ones(n, dtype=_ld)

    def corr(A):
        R = ld_rbf(A, A, ls, constant)
        R[np.diag_indices(len(A))] = _ld(constant) + _ld(noise)          # WhiteKernel on the diagonal of k(X)
        return R
    L = ld_cholesky(corr(X) + _ld(nugget) * np.eye(n, dtype=_ld))
    eta0, V0, tau0sq = _ld(center0), _ld(disp0), _ld(scale0) ** 2
    Ri1 = ld_cho_solve(L, one)
    bb = one @ Ri1
    ybar = y.mean(axis=1)
    V = _ld(0) if V0 == 0 else 1 / (1 / V0 + nc * bb)
    eta = eta0 if V0 == 0 else V * (eta0 / V0 + nc * (one @ ld_cho_solve(L, ybar)))
    df = df0 + n * nc
    if np.isinf(df0):
        tausq = tau0sq
    else:
        yc_ = y - ybar[:, None]
        quad = np.trace(yc_.T @ ld_cho_solve(L, yc_))
        a = ybar - eta0 * one
        Ria = ld_cho_solve(L, a)
        quad2 = nc * (a @ Ria - nc * V * (one @ Ria) ** 2)
        tausq = (_ld(df0) * tau0sq + quad + quad2) / _ld(df)
    var = tausq if np.isinf(df) else _ld(df) * tausq / (_ld(df) - 2)
    out = dict(center=float(eta), disp=float(V), df=float(df), scale=float(np.sqrt(tausq)), cov_factor=float(var))
    if Xc is None:
        Xo, Lo, yo = X, L, y
    else:
        Xo = np.asarray(Xc, dtype=float)
        Lo = ld_cholesky(corr(Xo) + _ld(nugget) * np.eye(len(Xo), dtype=_ld))
        yo = np.asarray(yc, dtype=_ld)
        yo = yo[:, None] if yo.ndim == 1 else yo
    R_on = ld_rbf(Xo, Xn, ls, constant)                               # WhiteKernel contributes nothing to k(X, Y)
    R_nn = corr(Xn)
    alpha = ld_cho_solve(Lo, yo - eta)
    mean = eta + R_on.T @ alpha
    R_pred = R_nn - R_on.T @ ld_cho_solve(Lo, R_on)
    if pred_noise:
        R_pred = R_pred + _ld(nugget) * np.eye(len(Xn), dtype=_ld)
    K = var * R_pred
    std = np.sqrt(np.diag(K))
    if student:
        bt = np.ones(len(Xn), dtype=_ld) - R_on.T @ ld_cho_solve(Lo, np.ones(len(Xo), dtype=_ld))
        mc = var * V * np.outer(bt, bt)
        std = std + np.sqrt(np.diag(mc))
        K = K + mc
    out.update(mean=np.squeeze(mean).astype(float), std=std.astype(float), cov=K.astype(float))
    return out


def as_close_as_reference(dev, ref, exact, rtol=1e-10, k=4.0):
    """True when `dev` matches `ref` to rtol, or else is within k x (the reference's own error) of the extended-precision
    value — array-level errors, scaled by max |exact|."""
    if relerr(dev, ref) < rtol:
        return True
    e_dev, e_ref = relerr(dev, exact), relerr(ref, exact)
    return e_dev <= k * e_ref + rtol


def ld_forward_solve(L, B):
    """L^-1 B by forward substitution in extended precision (L lower triangular)."""
    W = np.array(np.asarray(B, dtype=_ld), dtype=_ld)
    if W.ndim == 1:
        W = W[:, None]
    for i in range(L.shape[0]):
        W[i] = (W[i] - L[i, :i] @ W[:i]) / L[i, i]
    return W
