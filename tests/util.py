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
