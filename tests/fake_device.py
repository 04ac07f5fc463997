"""TEST INFRASTRUCTURE: numpy / LAPACK stand-ins for the wrappers of `gsum_b200.ops`, written from the contract of
include/gsum_b200.h (what each entry point must return), so that the Python host of the facade — argument marshalling, the
truncation scalings, order bookkeeping, caches, error mapping — runs in the CPU suite against the golden vectors of the real
reference.  `install(monkeypatch)` swaps them in for one test; nothing in the product imports this file, and the device kernels
themselves are only ever judged by the `-m gpu` tests, which run the same test bodies on the real library
(tests/test_facade_fake_device.py re-uses the functions of tests/test_gpu_lml.py and tests/test_gpu_predict.py).
"""
import numpy as np
import scipy.linalg as sl
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel

from oracle import gsum_oracle as o

PREDICT_MEAN, PREDICT_VAR, PREDICT_COV = 0, 1, 2


def _kernel(ls, constant, noise):
    ls = np.atleast_1d(np.asarray(ls, dtype=float))
    k = ConstantKernel(constant, 'fixed') * RBF(ls if ls.size > 1 else float(ls[0]), 'fixed')
    return k + WhiteKernel(noise, 'fixed') if noise > 0 else k


def kernel_matrix(X1, X2, length_scale, constant=1.0, noise=0.0, ctx=None):
    """gsum_kernel_matrix: c * RBF(X1, X2), + noise on the diagonal when X2 is None."""
    X1 = np.atleast_2d(np.asarray(X1, dtype=float))
    ls = np.atleast_1d(np.asarray(length_scale, dtype=float))
    k = RBF(ls if ls.size > 1 else float(ls[0]))
    if X2 is None:
        return constant * k(X1) + noise * np.eye(len(X1))
    return constant * k(X1, np.atleast_2d(np.asarray(X2, dtype=float)))


def _gs(x, start, end, excluded):
    return o.geometric_sum(x, start, end, excluded=None if excluded is None else list(np.atleast_1d(excluded)))


def process_cov(X1, X2, length_scale, constant=1.0, noise=0.0, factor=1.0, sc1=None, sc2=None, q1=None, q2=None,
                gs_start=0.0, gs_end=np.inf, excluded=None, kernel_add=0.0, ctx=None):
    """gsum_process_cov: (sc sc') * gs(q q') * (factor * (k + kernel_add)); X2 None = k(X1) with the white noise."""
    X1 = np.atleast_2d(np.asarray(X1, dtype=float))
    n1 = len(X1)
    K = factor * (kernel_matrix(X1, X2, length_scale, constant, noise) + kernel_add)
    if X2 is None:
        sc2, q2 = sc1, q1
    n2 = K.shape[1]
    if q1 is not None:
        K = _gs(np.outer(np.broadcast_to(q1, (n1,)), np.broadcast_to(q2, (n2,))), gs_start, gs_end, excluded) * K
    if sc1 is not None:
        K = np.outer(np.broadcast_to(sc1, (n1,)), np.broadcast_to(sc2, (n2,))) * K
    return K


def _priors(center0, disp0, df0, scale0):
    return o.Priors(center=center0, disp=disp0, df=df0, scale=scale0)


def lml_grid(X, dy, ref, orders, ls, Q, q_x_dependent=False, detf=None, constant=1.0, noise=0.0, nugget=1e-10,
             center0=0.0, disp0=0.0, df0=1.0, scale0=1.0, student=False, return_status=False, ctx=None):
    """gsum_lml_grid: ll[q, l] = ll_c(theta_l; dy / (ref Q_q^orders)) - detf[q]; a length scale whose R has no Cholesky
    factor gives -inf cells, nan logdet and a non-zero status."""
    X = np.atleast_2d(np.asarray(X, dtype=float))
    n = len(X)
    dy = np.asarray(dy, dtype=float)
    if dy.shape[1] > 16:
        raise ValueError("gsum_lml_grid: bad argument")
    ref = np.broadcast_to(np.asarray(ref, dtype=float), (n,))
    orders = np.asarray(orders)
    ls = np.asarray(ls, dtype=float).reshape(len(ls), -1)
    Q = np.asarray(Q, dtype=float)
    n_q, n_ls = Q.shape[0], ls.shape[0]
    detf = np.zeros(n_q) if detf is None else np.broadcast_to(np.asarray(detf, dtype=float), (n_q,))
    pri = _priors(center0, disp0, df0, scale0)
    fn = o.student_lml if student else o.gaussian_lml
    ll, logdet, status = np.empty((n_q, n_ls)), np.empty(n_ls), np.zeros(n_ls, dtype=np.int32)
    for j in range(n_ls):
        kern = _kernel(ls[j], constant, noise)
        R = kern(X) + nugget * np.eye(n)
        try:
            logdet[j] = 2 * np.sum(np.log(np.diag(np.linalg.cholesky(R))))
        except np.linalg.LinAlgError:
            ll[:, j], logdet[j], status[j] = -np.inf, np.nan, 1
            continue
        for i in range(n_q):
            q = Q[i][:, None] if q_x_dependent else Q[i]
            coeffs = dy / (ref[:, None] * q ** orders)
            ll[i, j] = fn(kern, kern.theta, X, coeffs, pri, nugget) - detf[i]
    return (ll, logdet, status) if return_status else ll


class FitHandle:
    """gsum_fit_create / gsum_predict (see the comments of include/gsum_b200.h on both)."""

    def __init__(self, X, y, length_scale, constant=1.0, noise=0.0, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0,
                 scale0=1.0, student=False, want_L=False, ctx=None):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        y = np.asarray(y, dtype=float)
        y = y if y.ndim == 2 else y[:, None]
        self.X, self.y, self.n, self.n_c = X, y, len(X), y.shape[1]
        self.ls, self.c, self.noise, self.nugget = np.atleast_1d(length_scale), float(constant), float(noise), float(nugget)
        kern = _kernel(self.ls, self.c, self.noise)
        try:
            f = o.fit_conjugate(kern, X, y, _priors(center0, disp0, df0, scale0), nugget=nugget, student=student)
        except np.linalg.LinAlgError:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        self.center, self.disp = float(f["center"][0]), float(f["disp"][0, 0])
        self.df, self.scale, self.cov_factor, self.lml = float(f["df"]), float(f["scale"]), float(f["cov_factor"]), float(f["lml"])
        self._L = f["corr_L"]
        self.logdet = float(2 * np.sum(np.log(np.diag(self._L))))
        self.L = self._L.copy() if want_L else None
        self.handle = object()

    def close(self):
        self.handle = None

    def predict(self, Xnew, want=PREDICT_MEAN, Xc=None, yc=None, mean_old=None, mean_new=None, basis_old=None,
                basis_new=None, sc_old=None, sc_new=None, q_old=None, q_new=None, gs_start=0.0, gs_end=np.inf,
                excluded=None, truncation=False, pred_noise=False, want_cond_basis=False, kernel_add=0.0):
        Xn = np.atleast_2d(np.asarray(Xnew, dtype=float))
        m = len(Xn)
        Xo = self.X if Xc is None else np.atleast_2d(np.asarray(Xc, dtype=float))
        n = len(Xo)
        if yc is None:
            if Xc is not None:
                raise ValueError("y must be given together with Xc")
            yo = self.y
        else:
            yo = np.asarray(yc, dtype=float)
            yo = yo if yo.ndim == 2 else yo[:, None]
            if len(yo) != n:
                raise ValueError("conditioning y must have one row per conditioning point")
        vec = lambda v, k: np.zeros(k) if v is None else np.broadcast_to(np.asarray(v, dtype=float), (k,))
        m_old, m_new = vec(mean_old, n), vec(mean_new, m)
        if truncation:
            cov = lambda A, B, sa, sb, qa, qb: process_cov(A, B, self.ls, self.c, 0.0, factor=self.cov_factor, sc1=sa, sc2=sb, q1=qa,
                                                           q2=qb, gs_start=gs_start, gs_end=gs_end, excluded=excluded, kernel_add=kernel_add)
            K_oo, K_on = cov(Xo, Xo, sc_old, sc_old, q_old, q_old), cov(Xo, Xn, sc_old, sc_new, q_old, q_new)
            try:
                L = np.linalg.cholesky(K_oo)
            except np.linalg.LinAlgError:
                raise np.linalg.LinAlgError("Matrix is not positive definite")
            scale, extra = 1.0, 0.0
            K_nn = (lambda: cov(Xn, Xn, sc_new, sc_new, q_new, q_new))

            def K_nn_diag():
                d = np.full(m, self.cov_factor * (self.c + kernel_add))
                if q_new is not None:
                    qn = np.broadcast_to(np.asarray(q_new, dtype=float), (m,))
                    d = _gs(qn * qn, gs_start, gs_end, excluded) * d
                if sc_new is not None:
                    sn = np.broadcast_to(np.asarray(sc_new, dtype=float), (m,))
                    d = (sn * sn) * d
                return d
        else:
            L = self._L if Xc is None else np.linalg.cholesky(kernel_matrix(Xo, None, self.ls, self.c, self.noise) + self.nugget * np.eye(n))
            K_on = kernel_matrix(Xo, Xn, self.ls, self.c, 0.0)
            scale, extra = self.cov_factor, (self.nugget if pred_noise else 0.0)
            K_nn = (lambda: kernel_matrix(Xn, None, self.ls, self.c, self.noise))
            K_nn_diag = (lambda: np.full(m, self.c + self.noise))
        mean = m_new[:, None] + K_on.T @ sl.cho_solve((L, True), yo - m_old[:, None])
        var = None
        if want == PREDICT_VAR:
            var = scale * (K_nn_diag() + extra - np.sum(K_on * sl.cho_solve((L, True), K_on), axis=0))
        elif want == PREDICT_COV:
            var = scale * (K_nn() + extra * np.eye(m) - K_on.T @ sl.cho_solve((L, True), K_on))
            var = 0.5 * (var + var.T)                                   # the device writes one triangle and mirrors it: exactly symmetric
        cb = None
        if want_cond_basis:
            cb = vec(basis_new, m) - K_on.T @ sl.cho_solve((L, True), vec(basis_old, n))
        return mean, var, cb


def pivoted_cholesky(M, ctx=None):
    """gsum_pivoted_cholesky: LAPACK dpstrf(lower) itself -> (G, Lp, piv (0-based), rank, status)."""
    from scipy.linalg.lapack import dpstrf
    c, p, rank, info = dpstrf(np.array(M, dtype=float), lower=1)
    Lp, piv = np.tril(c), (p - 1).astype(np.int32)
    G = np.empty_like(Lp)
    G[piv] = np.where(np.arange(len(piv))[None, :] < rank, Lp, 0.0)
    return G, Lp, piv, int(rank), int(info)


def draws(L, mean, Z=None, n_draws=None, seed=0, lower=None, upper=None, want_draws=True, first_draw=0, draw_scale=None,
          want_coverage=True, want_counts=False, ctx=None):
    """gsum_draws: mean + tril(L) (Z * draw_scale).  Device-generated normals (Z None) are a numpy stream here — the same
    distribution, so only statistical statements carry over; the fused coverage pass is not emulated."""
    if lower is not None or want_counts:
        raise NotImplementedError("fake device: the fused coverage pass is not emulated")
    L = np.asarray(L, dtype=float)
    if Z is None:
        Z = np.random.RandomState(int(seed) + 7919 * int(first_draw)).standard_normal((L.shape[0], int(n_draws)))
    Z = np.asarray(Z, dtype=float)
    if draw_scale is not None:
        Z = Z * np.asarray(draw_scale, dtype=float)[None, :]
    return np.broadcast_to(np.asarray(mean, dtype=float), (L.shape[0],))[:, None] + np.tril(L) @ Z, None


class ResidentFactors:
    """The state behind `Diagnostic` (ops.ResidentFactors): Cholesky and pivoted-Cholesky factors of one covariance."""

    def __init__(self, cov, ctx=None):
        cov = np.asarray(cov, dtype=float)
        self.n = cov.shape[0]
        try:
            self.L, self.chol_info = np.linalg.cholesky(cov), 0
        except np.linalg.LinAlgError:
            self.L, self.chol_info = np.full_like(cov, np.nan), 1
        self.G, self.Lp, self.piv, self.rank, self.status = pivoted_cholesky(cov)
    chol = property(lambda self: self.L)
    pchol = property(lambda self: self.G)
    pchol_L = property(lambda self: self.Lp)
    piv_host = property(lambda self: self.piv)


def pc_errors(Lp, piv, mean, Y, ctx=None):
    """gsum_pc_errors: solve(G, Y - mean) with G = Lp[p_inv], i.e. forward substitution on the pivot-ordered rows."""
    V = np.asarray(Y, dtype=float) - np.broadcast_to(mean, (len(Lp),))[:, None]
    return sl.solve_triangular(np.tril(Lp), V[np.asarray(piv)], lower=True)


def credible_interval(Y, lower, upper, ctx=None):
    """gsum_credible_interval: fraction of the points of each curve strictly inside each interval (gsum/diagnostics.py:163-164)."""
    Y, lower, upper = np.asarray(Y, dtype=float), np.atleast_2d(lower), np.atleast_2d(upper)
    return np.array([[np.mean((lo < y) & (y < up)) for lo, up in zip(lower, upper)] for y in Y.T])


def quadratic_forms(A, mean, Y, ctx=None):
    V = np.asarray(Y, dtype=float) - np.broadcast_to(mean, (len(A),))[:, None]
    return np.einsum("ik,ij,jk->k", V, np.asarray(A, dtype=float), V)


def lml_grad_terms(X, rhs, ls, constant=1.0, noise=0.0, nugget=1e-10, decomposition="cholesky", ctx=None):
    """gsum_lml_grad_terms: G = RHS^T Z, H_p = Z^T dR_p Z, tr_p = trace(R^-1 dR_p) for Z = R^-1 RHS and the derivatives of
    R = c RBF + (noise + nugget) I with respect to log c, log l_q, log noise; log|R|; info != 0 when R has no Cholesky factor."""
    X = np.atleast_2d(np.asarray(X, dtype=float))
    ls = np.atleast_1d(np.asarray(ls, dtype=float))
    n = len(X)
    K = kernel_matrix(X, None, ls, constant, 0.0)
    R = K + (noise + nugget) * np.eye(n)
    P = len(ls) + 2
    r = rhs.shape[1]
    try:
        L = np.linalg.cholesky(R)
    except np.linalg.LinAlgError:
        return np.zeros((r, r)), np.zeros((P, r, r)), np.zeros(P), np.nan, 1
    d2 = ((X[:, None, :] - X[None, :, :]) / ls) ** 2                       # (n, n, d): squared scaled differences per feature
    dR = [K] + ([K * d2.sum(-1)] if len(ls) == 1 else [K * d2[:, :, q] for q in range(len(ls))]) + [noise * np.eye(n)]
    Z = sl.cho_solve((L, True), np.asarray(rhs, dtype=float))
    Rinv = sl.cho_solve((L, True), np.eye(n))
    H = np.array([Z.T @ D @ Z for D in dR])
    tr = np.array([np.sum(Rinv * D) for D in dR])
    return rhs.T @ Z, H, tr, float(2 * np.sum(np.log(np.diag(L)))), 0


def cholesky(A, return_info=False, ctx=None):
    A = np.asarray(A, dtype=float)
    try:
        L = np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        if return_info:
            return A, 1, np.nan
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return (L, 0, 2 * np.sum(np.log(np.diag(L)))) if return_info else L


def cho_solve(L, B, forward_only=False, ctx=None):
    return sl.solve_triangular(L, B, lower=True) if forward_only else sl.cho_solve((L, True), B)


def cholesky_errors(L, mean, Y, want_errors=True, want_md2=False, ctx=None):
    E = sl.solve_triangular(L, np.asarray(Y, dtype=float) - np.broadcast_to(mean, (L.shape[0],))[:, None], lower=True)
    return (E if want_errors else None), (np.sum(E * E, axis=0) if want_md2 else None)


def install(monkeypatch):
    """Swap the stand-ins in for the duration of one test (pytest's monkeypatch undoes it)."""
    from gsum_b200 import ops
    for name in ("kernel_matrix", "process_cov", "lml_grid", "FitHandle", "pivoted_cholesky", "draws", "cholesky", "cho_solve",
                 "cholesky_errors", "ResidentFactors", "pc_errors", "credible_interval", "quadratic_forms", "lml_grad_terms"):
        monkeypatch.setattr(ops, name, globals()[name])
