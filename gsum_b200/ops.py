"""numpy-in / numpy-out wrappers over the C ABI (one function per exported entry point).

These are the calls the facade classes in ``models.py`` / ``diagnostics.py`` make and the calls the
parity tests exercise.  Every array crosses as a host pointer (``GSUM_MEM_HOST``); ``*_device`` variants
take torch CUDA tensors and enqueue on the context's stream.  No numerical work happens in Python here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, PREDICT_COV, PREDICT_MEAN, PREDICT_VAR, DeviceBuffer, as_f64, default_context

MEM_FACTOR_DEVICE = 2        # GSUM_MEM_FACTOR_DEVICE: the factor argument is a DeviceBuffer, everything else numpy

__all__ = [
    "kernel_matrix", "cholesky", "cho_solve", "lml_grid", "grid_normalize", "FitHandle", "process_cov",
    "cholesky_errors", "pivoted_cholesky", "pc_errors", "draws", "credible_interval", "ResidentFactors",
    "eigh", "eig_solve", "eig_conditional", "ResidentEigen",
]


def _p(a):
    return None if a is None else a.ctypes.data


def _factor(L):
    """(pointer, n, memory-kind bits) of a factor given as a numpy array or as a DeviceBuffer resident in HBM."""
    if isinstance(L, DeviceBuffer):
        return L.ptr, L.shape[0], MEM_HOST | MEM_FACTOR_DEVICE, L
    L = as_f64(L)
    return L.ctypes.data, L.shape[0], MEM_HOST, L


def _vec(a, n, name):
    if a is None:
        return None
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
    if a.shape != (n,):
        raise ValueError(f"{name} must have shape ({n},)")
    return a


def kernel_matrix(X1, X2, length_scale, constant=1.0, noise=0.0, ctx=None):
    """c * RBF(X1, X2) (+ noise on the diagonal when X2 is None) — sklearn kernel ``__call__``."""
    ctx = ctx or default_context()
    X1 = as_f64(np.atleast_2d(X1))
    n1, d = X1.shape
    ls = np.ascontiguousarray(np.atleast_1d(length_scale), dtype=np.float64)
    if X2 is None:
        out = np.empty((n1, n1))
        rc = ctx.lib.gsum_kernel_matrix(ctx.handle, _p(X1), n1, None, 0, d, _p(ls), ls.shape[0], constant, noise, _p(out), MEM_HOST)
    else:
        X2 = as_f64(np.atleast_2d(X2))
        out = np.empty((n1, X2.shape[0]))
        rc = ctx.lib.gsum_kernel_matrix(ctx.handle, _p(X1), n1, _p(X2), X2.shape[0], d, _p(ls), ls.shape[0], constant, noise,
                                        _p(out), MEM_HOST)
    ctx.check(rc, "gsum_kernel_matrix")
    return out


def cholesky(A, return_info=False, ctx=None):
    """Lower Cholesky factor(s) of A (n,n) or (batch,n,n) — ``numpy.linalg.cholesky``.

    Raises numpy.linalg.LinAlgError like numpy unless ``return_info`` (then returns L, info, logdet).
    """
    ctx = ctx or default_context()
    A = np.asarray(A)
    single = A.ndim == 2
    L = as_f64(A[None] if single else A, copy=True)
    batch, n, _ = L.shape
    info = np.zeros(batch, dtype=np.int32)
    logdet = np.zeros(batch)
    ctx.check(ctx.lib.gsum_cholesky(ctx.handle, _p(L), n, batch, _p(info), _p(logdet), MEM_HOST), "gsum_cholesky")
    if return_info:
        return (L[0], info[0], logdet[0]) if single else (L, info, logdet)
    if info.any():
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return L[0] if single else L


def cho_solve(L, B, forward_only=False, ctx=None):
    """``scipy.linalg.cho_solve((L, True), B)`` or, with forward_only, ``solve_triangular(L, B, lower=True)``."""
    ctx = ctx or default_context()
    Lp_, _, kind, _keep = _factor(L)
    B = np.asarray(B, dtype=np.float64)
    vec = B.ndim == 1
    X = as_f64(B[:, None] if vec else B, copy=True)
    n, k = X.shape
    ctx.check(ctx.lib.gsum_cho_solve(ctx.handle, Lp_, n, _p(X), k, 1 if forward_only else 0, kind), "gsum_cho_solve")
    return X[:, 0] if vec else X


def lml_grid(X, dy, ref, orders, ls, Q, q_x_dependent=False, detf=None, constant=1.0, noise=0.0, nugget=1e-10,
             center0=0.0, disp0=0.0, df0=1.0, scale0=1.0, student=False, return_status=False, ctx=None):
    """The (Q, l) log-marginal-likelihood grid: returns ll with shape (n_q, n_ls) (``[ratio][ls]``)."""
    ctx = ctx or default_context()
    X = as_f64(np.atleast_2d(X))
    n, d = X.shape
    dy = as_f64(dy)
    n_c = dy.shape[1]
    ref = _vec(ref, n, "ref")
    orders = np.ascontiguousarray(orders, dtype=np.int32)
    ls = as_f64(np.asarray(ls, dtype=np.float64).reshape(len(ls), -1))
    n_ls, ls_dim = ls.shape
    Q = as_f64(Q)
    if q_x_dependent:
        if Q.ndim != 2 or Q.shape[1] != n:
            raise ValueError("x-dependent Q must have shape (n_q, n)")
    elif Q.ndim != 1:
        raise ValueError("scalar Q must have shape (n_q,)")
    n_q = Q.shape[0]
    detf = None if detf is None else _vec(detf, n_q, "detf")
    ll = np.empty((n_q, n_ls))
    logdet = np.empty(n_ls)
    status = np.zeros(n_ls, dtype=np.int32)
    rc = ctx.lib.gsum_lml_grid(ctx.handle, _p(X), n, d, _p(dy), n_c, _p(ref), _p(orders), _p(ls), n_ls, ls_dim, _p(Q), n_q,
                               1 if q_x_dependent else 0, _p(detf), float(constant), float(noise), float(nugget),
                               float(center0), float(disp0), float(df0), float(scale0), 1 if student else 0, _p(ll),
                               _p(logdet), _p(status), MEM_HOST)
    ctx.check(rc, "gsum_lml_grid")
    return (ll, logdet, status) if return_status else ll


def lml_grad_terms(X, rhs, ls, constant=1.0, noise=0.0, nugget=1e-10, decomposition="cholesky", ctx=None):
    """Device half of the analytic likelihood gradient (gsum/models.py:957-1056): with RHS = [basis | curves] (n, r),
    Z = R^-1 RHS and dR_p = dR/dlog(theta_p) for theta = [constant, length scale(s), noise level]:
    returns G = RHS^T Z (r, r), H (P, r, r) = Z^T dR_p Z, tr (P,) = trace(R^-1 dR_p), logdet R, info."""
    ctx = ctx or default_context()
    X = as_f64(np.atleast_2d(X))
    n, d = X.shape
    rhs = as_f64(rhs)
    r = rhs.shape[1]
    ls = as_f64(np.atleast_1d(ls))
    P = ls.shape[0] + 2
    G, H, tr = np.empty((r, r)), np.empty((P, r, r)), np.empty(P)
    logdet, info = np.zeros(1), np.zeros(1, dtype=np.int32)
    fn = ctx.lib.gsum_lml_grad_terms_eig if decomposition == "eig" else ctx.lib.gsum_lml_grad_terms
    rc = ctx.check(fn(ctx.handle, _p(X), n, d, _p(rhs), r, _p(ls), ls.shape[0], float(constant), float(noise),
                      float(nugget), _p(G), _p(H), _p(tr), _p(logdet), _p(info), MEM_HOST), "gsum_lml_grad_terms")
    if rc:
        raise np.linalg.LinAlgError(_eigh_message(ctx))
    return G, H, tr, float(logdet[0]), int(info[0])


def grid_normalize(ll, ctx=None):
    """exp(ll - max) and logsumexp(ll) on the device."""
    ctx = ctx or default_context()
    ll = as_f64(ll)
    post = np.empty_like(ll)
    lse = np.zeros(1)
    ctx.check(ctx.lib.gsum_grid_normalize(ctx.handle, _p(ll), ll.size, _p(post), _p(lse), MEM_HOST), "gsum_grid_normalize")
    return post, float(lse[0])


class _PredictArgs(C.Structure):
    _fields_ = [
        ("Xnew", C.c_void_p), ("m", C.c_int64), ("Xc", C.c_void_p), ("n_cond", C.c_int64), ("yc", C.c_void_p), ("n_y", C.c_int32),
        ("mean_old", C.c_void_p), ("mean_new", C.c_void_p), ("basis_old", C.c_void_p), ("basis_new", C.c_void_p),
        ("sc_old", C.c_void_p), ("sc_new", C.c_void_p), ("q_old", C.c_void_p), ("q_new", C.c_void_p),
        ("gs_start", C.c_double), ("gs_end", C.c_double), ("excluded", C.c_void_p), ("n_excluded", C.c_int32),
        ("kernel_add", C.c_double), ("truncation", C.c_int32), ("want", C.c_int32), ("pred_noise", C.c_int32),
        ("mean_out", C.c_void_p), ("var_out", C.c_void_p), ("cond_basis_out", C.c_void_p),
    ]


class FitHandle:
    """Device-resident fit (``gsum_fit``): X, y and the lower factor stay in HBM between predict calls."""

    def __init__(self, X, y, length_scale, constant=1.0, noise=0.0, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0,
                 scale0=1.0, student=False, want_L=False, ctx=None):
        self.ctx = ctx or default_context()
        X = as_f64(np.atleast_2d(X))
        y = as_f64(y if np.ndim(y) == 2 else np.asarray(y)[:, None])
        self.n, self.d = X.shape
        self.n_c = y.shape[1]
        ls = np.ascontiguousarray(np.atleast_1d(length_scale), dtype=np.float64)
        out7 = np.zeros(7)
        L = np.empty((self.n, self.n)) if want_L else None
        h = C.c_void_p()
        rc = self.ctx.lib.gsum_fit_create(self.ctx.handle, _p(X), self.n, self.d, _p(y), self.n_c, _p(ls), ls.shape[0],
                                          float(constant), float(noise), float(nugget), float(center0), float(disp0),
                                          float(df0), float(scale0), 1 if student else 0, _p(out7), _p(L), MEM_HOST,
                                          C.byref(h))
        rc = self.ctx.check(rc, "gsum_fit_create")
        if rc > 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        self.handle = h
        self.center, self.disp, self.df, self.scale, self.cov_factor, self.lml, self.logdet = (float(v) for v in out7)
        self.L = L
        self.kernel_args = (ls, float(constant), float(noise))

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.gsum_fit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def predict(self, Xnew, want=PREDICT_MEAN, Xc=None, yc=None, mean_old=None, mean_new=None, basis_old=None,
                basis_new=None, sc_old=None, sc_new=None, q_old=None, q_new=None, gs_start=0.0, gs_end=np.inf,
                excluded=None, truncation=False, pred_noise=False, want_cond_basis=False, kernel_add=0.0):
        """One GP conditional (see ``gsum_predict_args`` in include/gsum_b200.h).  Returns (mean, var_or_cov, cond_basis)."""
        Xnew = as_f64(np.atleast_2d(Xnew))
        m = Xnew.shape[0]
        a = _PredictArgs()
        keep = [Xnew]
        a.Xnew, a.m = _p(Xnew), m
        n = self.n
        if Xc is not None:
            Xc = as_f64(np.atleast_2d(Xc))
            n = Xc.shape[0]
            a.Xc, a.n_cond = _p(Xc), n
            keep.append(Xc)
        n_y = self.n_c
        if yc is not None:
            yc = as_f64(yc if np.ndim(yc) == 2 else np.asarray(yc, dtype=np.float64)[:, None])
            if yc.shape[0] != n:
                raise ValueError("conditioning y must have one row per conditioning point")
            n_y = yc.shape[1]
            a.yc, a.n_y = _p(yc), n_y
            keep.append(yc)
        elif Xc is not None:
            raise ValueError("y must be given together with Xc")

        def vec(v, length, name):
            v = _vec(v, length, name)
            if v is not None:
                keep.append(v)
            return _p(v)

        a.mean_old, a.mean_new = vec(mean_old, n, "mean_old"), vec(mean_new, m, "mean_new")
        a.basis_old, a.basis_new = vec(basis_old, n, "basis_old"), vec(basis_new, m, "basis_new")
        a.sc_old, a.sc_new = vec(sc_old, n, "sc_old"), vec(sc_new, m, "sc_new")
        a.q_old, a.q_new = vec(q_old, n, "q_old"), vec(q_new, m, "q_new")
        a.gs_start, a.gs_end = float(gs_start), float(gs_end)
        if excluded is not None:
            ex = np.ascontiguousarray(np.atleast_1d(excluded), dtype=np.int32)
            if ex.size > 8:
                raise NotImplementedError("gsum_b200: at most 8 excluded orders")
            a.excluded, a.n_excluded = _p(ex), ex.size
            keep.append(ex)
        a.kernel_add = float(kernel_add)
        a.truncation, a.want, a.pred_noise = int(bool(truncation)), int(want), int(bool(pred_noise))
        mean = np.empty((m, n_y))
        a.mean_out = _p(mean)
        var = None
        if want == PREDICT_VAR:
            var = np.empty(m)
        elif want == PREDICT_COV:
            var = np.empty((m, m))
        a.var_out = _p(var)
        cb = np.empty(m) if want_cond_basis else None
        a.cond_basis_out = _p(cb)
        rc = self.ctx.check(self.ctx.lib.gsum_predict(self.ctx.handle, self.handle, C.byref(a), MEM_HOST), "gsum_predict")
        if rc > 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        return mean, var, cb


def process_cov(X1, X2, length_scale, constant=1.0, noise=0.0, factor=1.0, sc1=None, sc2=None, q1=None, q2=None,
                gs_start=0.0, gs_end=np.inf, excluded=None, kernel_add=0.0, ctx=None):
    """factor * (kernel(X1, X2) + kernel_add) with the truncation scalings (gsum/models.py:562-599, 1342-1348)."""
    ctx = ctx or default_context()
    X1 = as_f64(np.atleast_2d(X1))
    n1, d = X1.shape
    ls = np.ascontiguousarray(np.atleast_1d(length_scale), dtype=np.float64)
    X2c = None if X2 is None else as_f64(np.atleast_2d(X2))
    n2 = n1 if X2c is None else X2c.shape[0]
    sc1, q1 = _vec(sc1, n1, "sc1"), _vec(q1, n1, "q1")
    sc2, q2 = (None, None) if X2c is None else (_vec(sc2, n2, "sc2"), _vec(q2, n2, "q2"))
    ex = None if excluded is None else np.ascontiguousarray(np.atleast_1d(excluded), dtype=np.int32)
    out = np.empty((n1, n2))
    rc = ctx.lib.gsum_process_cov(ctx.handle, d, _p(ls), ls.shape[0], float(constant), float(noise), _p(X1), n1, _p(X2c), n2,
                                  _p(sc1), _p(sc2), _p(q1), _p(q2), float(gs_start), float(gs_end), _p(ex),
                                  0 if ex is None else ex.size, float(factor), float(kernel_add), _p(out), MEM_HOST)
    ctx.check(rc, "gsum_process_cov")
    return out


def cholesky_errors(L, mean, Y, want_errors=True, want_md2=False, ctx=None):
    """L^{-1}(Y - mean) for Y (n, n_curves) and/or the squared Mahalanobis distances (n_curves,)."""
    ctx = ctx or default_context()
    Lp_, n, kind, _keep = _factor(L)
    Y = as_f64(Y)
    mean = _vec(mean, n, "mean")
    E = np.empty_like(Y) if want_errors else None
    md2 = np.empty(Y.shape[1]) if want_md2 else None
    ctx.check(ctx.lib.gsum_cholesky_errors(ctx.handle, Lp_, n, _p(mean), _p(Y), Y.shape[1], _p(E), _p(md2), kind),
              "gsum_cholesky_errors")
    return E, md2


def quadratic_forms(A, mean, Y, ctx=None):
    """(y_c - mean)^T A (y_c - mean) for every column of Y (n, n_curves) — the diagonal of mahalanobis(inv=...)'s product."""
    ctx = ctx or default_context()
    A, Y = as_f64(A), as_f64(Y)
    n = A.shape[0]
    if A.shape != (n, n) or Y.shape[0] != n:
        raise ValueError("quadratic_forms: A must be (n, n) and Y (n, n_curves)")
    mean = _vec(mean, n, "mean")
    q = np.empty(Y.shape[1])
    ctx.check(ctx.lib.gsum_quadratic_forms(ctx.handle, _p(A), n, _p(mean), _p(Y), Y.shape[1], _p(q), MEM_HOST), "gsum_quadratic_forms")
    return q


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pointwise_fit(y, orders, mask, excluded, ratio, ref, df0, scale0, ctx=None):
    """TruncationPointwise.fit on the device: (coeffs (n, n_m), scale (n,), trunc_scale (n, n_m)); ratio, ref (n,)."""
    ctx = ctx or default_context()
    y = as_f64(y)
    n, n_o = y.shape
    orders, mask, excluded = _i32(orders), _i32(mask), _i32(excluded)
    n_m = int(mask.sum())
    ratio, ref = _vec(ratio, n, "ratio"), _vec(ref, n, "ref")
    coeffs, scale, trunc = np.empty((n, n_m)), np.empty(n), np.empty((n, n_m))
    ctx.check(ctx.lib.gsum_pointwise_fit(ctx.handle, _p(y), n, n_o, _p(orders), _p(mask), _p(excluded) if excluded.size else None,
                                         excluded.size, _p(ratio), _p(ref), float(df0), float(scale0), _p(coeffs), _p(scale), _p(trunc),
                                         MEM_HOST), "gsum_pointwise_fit")
    return coeffs, scale, trunc


def pointwise_loglike_sums(y, orders, mask, ratios, ref, df0, scale0, ctx=None):
    """The two point sums of TruncationPointwise.log_likelihood for every row of `ratios` (n_r, 1 or n); ref (1 or n,)."""
    ctx = ctx or default_context()
    y = as_f64(y)
    n, n_o = y.shape
    ratios, ref = as_f64(np.atleast_2d(ratios)), as_f64(np.atleast_1d(ref))
    orders, mask = _i32(orders), _i32(mask)              # named: the arrays must outlive the call
    n_r = ratios.shape[0]
    S1, S2 = np.empty(n_r), np.empty(n_r)
    ctx.check(ctx.lib.gsum_pointwise_loglike(ctx.handle, _p(y), n, n_o, _p(orders), _p(mask), _p(ratios), n_r, ratios.shape[1],
                                             _p(ref), ref.shape[0], float(df0), float(scale0), _p(S1), _p(S2), MEM_HOST),
              "gsum_pointwise_loglike")
    return S1, S2


def variogram_bins(X, z, bounds, ctx=None):
    """Pairs (np.tril_indices order), their bins and the per-bin sums of VariogramFourthRoot.__init__; z (ncurves, n)."""
    ctx = ctx or default_context()
    X, z, bounds = as_f64(X), as_f64(z), as_f64(bounds)
    n, d = X.shape
    nc, nb, npairs = z.shape[0], bounds.shape[0] + 1, n * (n - 1) // 2
    grid, hij, bidx = np.empty((n, n), np.int32), np.empty(npairs), np.empty(npairs, np.int32)
    dij, counts, hsum, dsum = np.empty((npairs, nc)), np.empty(nb, np.int64), np.empty(nb), np.empty((nb, nc))
    ctx.check(ctx.lib.gsum_variogram_bins(ctx.handle, _p(X), n, d, _p(z), nc, _p(bounds), bounds.shape[0], _p(grid), _p(hij), _p(bidx),
                                          _p(dij), _p(counts), _p(hsum), _p(dsum), MEM_HOST), "gsum_variogram_bins")
    return grid, hij, bidx, dij, counts, hsum, dsum


def variogram_cov(i1, j1, i2, j2, bin_grid, gamma_tilde, tab, var_factor, corr_factor, same_is_one=True, ctx=None):
    """VariogramFourthRoot.cov of two bins given their pair lists; gamma_tilde (nbins, ncurves <= 8) -> (ncurves,)."""
    ctx = ctx or default_context()
    i1, j1, i2, j2, bin_grid = _i32(i1), _i32(j1), _i32(i2), _i32(j2), _i32(bin_grid)
    gamma_tilde, tab = as_f64(gamma_tilde), as_f64(tab)
    out = np.empty(gamma_tilde.shape[1])
    ctx.check(ctx.lib.gsum_variogram_cov(ctx.handle, _p(i1), _p(j1), i1.shape[0], _p(i2), _p(j2), i2.shape[0], _p(bin_grid),
                                         bin_grid.shape[0], _p(gamma_tilde), gamma_tilde.shape[0], gamma_tilde.shape[1], _p(tab),
                                         float(var_factor), float(corr_factor), 1 if same_is_one else 0, _p(out), MEM_HOST), "gsum_variogram_cov")
    return out


def pivoted_cholesky(M, ctx=None):
    """LAPACK dpstrf(lower) on the device: returns (G, Lp, piv, rank, status); M = G G^T, P^T M P = Lp Lp^T."""
    ctx = ctx or default_context()
    M = as_f64(M)
    n = M.shape[0]
    Lp, G = np.empty((n, n)), np.empty((n, n))
    piv = np.zeros(n, dtype=np.int32)
    rank = C.c_int32(0)
    rc = ctx.check(ctx.lib.gsum_pivoted_cholesky(ctx.handle, _p(M), n, _p(Lp), _p(piv), C.addressof(rank), _p(G), MEM_HOST),
                   "gsum_pivoted_cholesky")
    return G, Lp, piv, int(rank.value), rc


def pc_errors(Lp, piv, mean, Y, ctx=None):
    """solve(G, Y - mean) with G = Lp[p_inv] by permutation + forward substitution."""
    ctx = ctx or default_context()
    Lp_, n, kind, _keep = _factor(Lp)
    Y = as_f64(Y)
    if isinstance(piv, DeviceBuffer) != isinstance(Lp, DeviceBuffer):
        raise ValueError("Lp and piv must both be numpy arrays or both DeviceBuffers")
    piv_host = None if isinstance(piv, DeviceBuffer) else np.ascontiguousarray(piv, dtype=np.int32)   # named: outlives the call
    piv_ = piv.ptr if piv_host is None else _p(piv_host)
    mean = _vec(mean, n, "mean")
    E = np.empty_like(Y)
    ctx.check(ctx.lib.gsum_pc_errors(ctx.handle, Lp_, piv_, n, _p(mean), _p(Y), Y.shape[1], _p(E), kind), "gsum_pc_errors")
    return E


def draws(L, mean, Z=None, n_draws=None, seed=0, lower=None, upper=None, want_draws=True, first_draw=0, draw_scale=None,
          want_coverage=True, want_counts=False, ctx=None):
    """mean + L z for caller-supplied Z (n, n_draws) or device Philox normals; optional fused coverage.

    `first_draw` places this call's draws at global positions first_draw .. first_draw + n_draws - 1 of the seed's
    stream (sharding of the draw axis); `draw_scale` (n_draws,) multiplies each draw's normals (multivariate t).
    Returns (draws (n, n_draws) or None, coverage (n_draws, n_alpha) or None) — and, with `want_counts`, a third item:
    the int64 (n_alpha,) number of (draw, point) pairs inside each interval."""
    ctx = ctx or default_context()
    Lp_, n, kind, _keep = _factor(L)
    mean = _vec(mean, n, "mean")
    if Z is not None:
        Z = as_f64(Z)
        n_draws = Z.shape[1]
    if not n_draws:
        raise ValueError("n_draws must be given when Z is None")
    draw_scale = _vec(draw_scale, n_draws, "draw_scale")
    out = np.empty((n, n_draws)) if want_draws else None
    cov, counts, n_alpha = None, None, 0
    if lower is not None:
        lower, upper = as_f64(np.atleast_2d(lower)), as_f64(np.atleast_2d(upper))
        n_alpha = lower.shape[0]
        cov = np.empty((n_draws, n_alpha)) if want_coverage else None
        counts = np.zeros(n_alpha, dtype=np.int64) if want_counts else None
    elif want_counts:
        raise ValueError("coverage counts need lower and upper")
    ctx.check(ctx.lib.gsum_draws(ctx.handle, Lp_, n, _p(mean), _p(Z), n_draws, int(seed), int(first_draw), _p(draw_scale),
                                 _p(out), _p(lower), _p(upper), n_alpha, _p(cov), _p(counts), kind), "gsum_draws")
    if want_counts:
        return out, cov, counts
    return out, cov


class ResidentFactors:
    """Cholesky and pivoted-Cholesky factors of one covariance, computed once and kept in HBM (the state behind
    `Diagnostic`, gsum/diagnostics.py:60-61): the matrix goes up ONCE, both factorisations run on the device copy, and
    every later error / draw call passes the resident factors (GSUM_MEM_FACTOR_DEVICE) instead of re-uploading N x N
    doubles.  `chol`, `pchol` (G = Lp[p_inv], the array gsum's `pivoted_cholesky` returns), `pchol_L` and `piv` are
    copied back to numpy on first access only."""

    def __init__(self, cov, ctx=None):
        self.ctx = ctx = ctx or default_context()
        cov = as_f64(cov)
        n = self.n = cov.shape[0]
        if cov.shape != (n, n):
            raise ValueError("cov must be square")
        self.L = DeviceBuffer(ctx, (n, n))
        self.Lp = DeviceBuffer(ctx, (n, n))
        self.G = DeviceBuffer(ctx, (n, n))
        self.piv = DeviceBuffer(ctx, (n,), np.int32)
        scal = DeviceBuffer(ctx, (2,), np.int32)                    # [cholesky info, pivoted-Cholesky rank]
        M = DeviceBuffer(ctx, (n, n)).put(cov)                      # the only N x N upload
        try:
            self.L.copy_from(M)
            ctx.check(ctx.lib.gsum_cholesky(ctx.handle, self.L.ptr, n, 1, scal.ptr, None, MEM_DEVICE), "gsum_cholesky")
            self.status = ctx.check(ctx.lib.gsum_pivoted_cholesky(ctx.handle, M.ptr, n, self.Lp.ptr, self.piv.ptr, scal.ptr + 4,
                                                                  self.G.ptr, MEM_DEVICE), "gsum_pivoted_cholesky")
            self.chol_info, self.rank = (int(v) for v in scal.get())
        finally:
            M.free()
            scal.free()
        self._host = {}

    def _lazy(self, name, buf):
        if name not in self._host:
            self._host[name] = buf.get()
        return self._host[name]

    chol = property(lambda self: self._lazy("chol", self.L))
    pchol = property(lambda self: self._lazy("pchol", self.G))
    pchol_L = property(lambda self: self._lazy("pchol_L", self.Lp))
    piv_host = property(lambda self: self._lazy("piv", self.piv))


def credible_interval(Y, lower, upper, ctx=None):
    """Coverage (n_curves, n_alpha) of curves Y (n, n_curves) for interval bounds lower/upper (n_alpha, n)."""
    ctx = ctx or default_context()
    Y = as_f64(Y)
    n, k = Y.shape
    lower, upper = as_f64(np.atleast_2d(lower)), as_f64(np.atleast_2d(upper))
    out = np.empty((k, lower.shape[0]))
    ctx.check(ctx.lib.gsum_credible_interval(ctx.handle, _p(Y), n, k, _p(lower), _p(upper), lower.shape[0], _p(out), MEM_HOST),
              "gsum_credible_interval")
    return out


def _eigh_message(ctx):
    msg = ctx.lib.gsum_last_error(ctx.handle)
    return msg.decode() if msg else "Eigenvalues did not converge"


class ResidentEigen:
    """Eigendecomposition A = V diag(w) V^T computed on the device and kept in HBM (the `_eigh_tuple_` of
    gsum/models.py:714-716 and the `_eig` factor of gsum/diagnostics.py:63-68): A goes up once, every later
    `eig_solve` passes the resident (w, V) (GSUM_MEM_FACTOR_DEVICE).  `w` is a numpy array; `V` is copied back on
    first access only."""

    def __init__(self, A, ctx=None):
        self.ctx = ctx = ctx or default_context()
        A = as_f64(A)
        n = self.n = A.shape[0]
        if A.shape != (n, n):
            raise ValueError("A must be square")
        self.w_dev = DeviceBuffer(ctx, (n,))
        self.V_dev = DeviceBuffer(ctx, (n, n))
        M = DeviceBuffer(ctx, (n, n)).put(A)
        sweeps = C.c_int32(0)
        try:
            self.status = ctx.check(ctx.lib.gsum_eigh(ctx.handle, M.ptr, n, self.w_dev.ptr, self.V_dev.ptr, C.addressof(sweeps),
                                                      MEM_DEVICE), "gsum_eigh")
        finally:
            M.free()
        ctx.synchronize()
        if self.status:
            raise np.linalg.LinAlgError(_eigh_message(ctx))                       # numpy / scipy eigh: LinAlgError
        self.sweeps = int(sweeps.value)
        self.w = self.w_dev.get()
        self._V = None

    @property
    def V(self):
        if self._V is None:
            self._V = self.V_dev.get()
        return self._V

    def solve(self, Y, mean=None, mode=0):
        return eig_solve((self.w_dev, self.V_dev), Y, mean=mean, mode=mode, ctx=self.ctx)

    def conditional(self, R_on, D=None, want_var=False, want_cov=False):
        return eig_conditional((self.w_dev, self.V_dev), R_on, D, want_var=want_var, want_cov=want_cov, ctx=self.ctx)


def eigh(A, return_sweeps=False, ctx=None):
    """(w, V) with A = V diag(w) V^T, w ascending, eigenvectors in the columns of V — ``scipy.linalg.eigh(A)`` up to the
    sign of each eigenvector (largest-magnitude component positive here)."""
    ctx = ctx or default_context()
    A = as_f64(A)
    n = A.shape[0]
    if A.shape != (n, n):
        raise ValueError("A must be square")
    w, V = np.empty(n), np.empty((n, n))
    sweeps = C.c_int32(0)
    rc = ctx.check(ctx.lib.gsum_eigh(ctx.handle, _p(A), n, _p(w), _p(V), C.addressof(sweeps), MEM_HOST), "gsum_eigh")
    if rc:
        raise np.linalg.LinAlgError(_eigh_message(ctx))
    return (w, V, int(sweeps.value)) if return_sweeps else (w, V)


def eig_solve(eig, Y, mean=None, mode=0, ctx=None):
    """mode 0: V diag(1/w) V^T (Y - mean) (``solve_sqrt(..., decomposition='eig')``, gsum/models.py:480-484);
    mode 1: diag(|w|^-1/2) V^T (Y - mean) (``eigen_errors``, rows in the order of w).  `eig` = (w, V) as numpy arrays
    or as DeviceBuffers resident in HBM."""
    ctx = ctx or default_context()
    w, V = eig
    if isinstance(V, DeviceBuffer) != isinstance(w, DeviceBuffer):
        raise ValueError("w and V must both be numpy arrays or both DeviceBuffers")
    if isinstance(V, DeviceBuffer):
        wp, Vp, n, kind = w.ptr, V.ptr, V.shape[0], MEM_HOST | MEM_FACTOR_DEVICE
    else:
        w, V = as_f64(w), as_f64(V)
        wp, Vp, n, kind = _p(w), _p(V), V.shape[0], MEM_HOST
    Y = np.asarray(Y, dtype=np.float64)
    vec = Y.ndim == 1
    Yc = as_f64(Y[:, None] if vec else Y)
    if Yc.shape[0] != n:
        raise ValueError(f"Y must have {n} rows")
    mean = _vec(mean, n, "mean")
    X = np.empty_like(Yc)
    ctx.check(ctx.lib.gsum_eig_solve(ctx.handle, wp, Vp, n, _p(Yc), Yc.shape[1], _p(mean), _p(X), int(mode), kind), "gsum_eig_solve")
    return X[:, 0] if vec else X


def eig_conditional(eig, R_on, D=None, want_var=False, want_cov=False, ctx=None):
    """(R_no R^-1 D (m, k) | None, diag(R_no R^-1 R_on) (m,) | None, R_no R^-1 R_on (m, m) | None) with
    R^-1 = V diag(1/w) V^T — the products `predict` needs on the 'eig' route (gsum/models.py:826-836)."""
    ctx = ctx or default_context()
    w, V = eig
    if isinstance(V, DeviceBuffer):
        wp, Vp, n, kind = w.ptr, V.ptr, V.shape[0], MEM_HOST | MEM_FACTOR_DEVICE
    else:
        w, V = as_f64(w), as_f64(V)
        wp, Vp, n, kind = _p(w), _p(V), V.shape[0], MEM_HOST
    R_on = as_f64(R_on)
    m = R_on.shape[1]
    lin = var = cov = None
    k = 0
    if D is not None:
        D = as_f64(D)
        k = D.shape[1]
        lin = np.empty((m, k))
    if want_var:
        var = np.empty(m)
    if want_cov:
        cov = np.empty((m, m))
    ctx.check(ctx.lib.gsum_eig_conditional(ctx.handle, wp, Vp, n, _p(R_on), m, _p(D), k, _p(lin), _p(var), _p(cov), kind),
              "gsum_eig_conditional")
    return lin, var, cov


def lml_grid_device(ctx, X, dy, ref, orders, ls, Q, detf, ll_out, q_x_dependent=False, constant=1.0, noise=0.0,
                    nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0, scale0=1.0, student=False, logdet_out=None,
                    status_out=None):
    """Device-resident variant of :func:`lml_grid`: every argument is a contiguous torch CUDA tensor on the context's
    device (float64; `orders` / `status_out` int32) and the call only enqueues work on the context's stream."""
    n, d = X.shape
    n_c = dy.shape[1]
    n_ls, ls_dim = ls.shape
    n_q = Q.shape[0]
    rc = ctx.lib.gsum_lml_grid(ctx.handle, X.data_ptr(), n, d, dy.data_ptr(), n_c, ref.data_ptr(), orders.data_ptr(),
                               ls.data_ptr(), n_ls, ls_dim, Q.data_ptr(), n_q, 1 if q_x_dependent else 0,
                               None if detf is None else detf.data_ptr(), float(constant), float(noise), float(nugget),
                               float(center0), float(disp0), float(df0), float(scale0), 1 if student else 0,
                               ll_out.data_ptr(), None if logdet_out is None else logdet_out.data_ptr(),
                               None if status_out is None else status_out.data_ptr(), MEM_DEVICE)
    ctx.check(rc, "gsum_lml_grid")
    return ll_out


def grid_normalize_device(ctx, ll, post_out, lse_out):
    ctx.check(ctx.lib.gsum_grid_normalize(ctx.handle, ll.data_ptr(), ll.numel(), post_out.data_ptr(), lse_out.data_ptr(),
                                          MEM_DEVICE), "gsum_grid_normalize")
