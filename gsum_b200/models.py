"""Drop-in facade for the reference's conjugate-process and truncation classes (gsum/models.py).

Same class names, constructor arguments, methods, fitted attributes and error behaviour as the
reference on the hot path, but every linear-algebra step (kernel matrix, Cholesky, triangular solves,
conjugate updates, likelihood, posterior moments) is one call into the C ABI / CUDA library — the
Python here only marshals arguments and evaluates the O(n_c^2) closed forms of the gradient / 'eig' routes on the
small Gram matrices the device returns.  What the device path does not cover raises NotImplementedError
(custom `basis`, kernels other than [Constant*]RBF[+White]); nothing falls back to numpy/scipy.
`decomposition='eig'` runs on the device Jacobi eigensolver (csrc/eig.cuh).

New, additive: ``TruncationProcess.log_marginal_likelihood_grid`` evaluates the whole (Q, l) grid of
docs/notebooks/correlated_EFT_publication.ipynb cell 53 in one device call (optionally sharded over the
ranks of a torch.distributed process group, see ``gsum_b200.distributed``).
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.optimize import fmin_l_bfgs_b
from sklearn.base import clone
from sklearn.exceptions import ConvergenceWarning
from sklearn.gaussian_process.kernels import RBF, ConstantKernel
from sklearn.utils import check_random_state

from . import ops
from ._lib import PREDICT_COV, PREDICT_MEAN, PREDICT_VAR
from .helpers import _order_differences, coefficients, geometric_sum
from .kernels import flatten_kernel, theta_layout

__all__ = ["BaseConjugateProcess", "ConjugateGaussianProcess", "ConjugateStudentProcess", "TruncationProcess",
           "TruncationGP", "TruncationTP"]


def _scalar_prior(value, name):
    a = np.asarray(value, dtype=np.float64)
    if a.size != 1:
        raise NotImplementedError(f"gsum_b200: only the default single-column basis is supported, so `{name}` must be scalar")
    return float(a.reshape(-1)[0])


def _conjugate_from_gram(G, logdet, N, nc, pri, student):
    """Conjugate posterior and log-likelihood from the Gram G = RHS^T R^-1 RHS of RHS = [basis | y_1..y_nc] and log|R|.

    compute_center gsum/models.py:201-221, compute_disp 260-271, compute_df 296, compute_scale_sq 419-448,
    compute_cov_factor 501-503; Gaussian likelihood 1010-1040, Student-t evidence 1241-1258.  Every vector the
    reference solves against R is RHS a for a short coefficient vector a, so each term is a quadratic form in G."""
    eta0, V0, df0, scale0 = pri["center0"], pri["disp0"], pri["df0"], pri["scale0"]
    r = nc + 1
    eB = np.zeros(r); eB[0] = 1.0
    eavg = np.zeros(r); eavg[1:] = 1.0 / nc
    E = np.eye(r)[1:]
    if V0 == 0:
        V, eta = 0.0, eta0
    else:
        V = 1.0 / (1.0 / V0 + nc * G[0, 0])
        eta = V * (eta0 / V0 + nc * (eB @ G @ eavg))
    df = df0 + N * nc
    if np.isinf(df0):
        scale2 = scale0 ** 2
    else:
        Ec = E - eavg[None, :]
        quad = sum(e @ G @ e for e in Ec)
        ac = eavg - eta0 * eB
        aw = nc * (ac - nc * V * (eB @ G @ ac) * eB)
        scale2 = (df0 * scale0 ** 2 + quad + ac @ G @ aw) / df
    var = scale2 if np.isinf(df) else df * scale2 / (df - 2)
    if student:
        from scipy.special import loggamma

        def log_norm(df_, scale2_, disp_):
            norm = loggamma(df_ / 2.0) - df_ / 2.0 * np.log(df_ * scale2_ / 2.0)
            if disp_ > 0:
                norm += 0.5 * np.log(2 * np.pi * disp_)
            return norm

        with np.errstate(invalid='ignore'):
            ll = log_norm(df, scale2, V) - log_norm(df0, scale0 ** 2, V0) - nc / 2.0 * (N * np.log(2 * np.pi) + logdet)
    else:
        Bk = E - eta * eB[None, :]
        yKy = sum(b @ G @ b for b in Bk) / var
        ll = -0.5 * yKy - 0.5 * nc * (N * np.log(var) + logdet) - 0.5 * nc * N * np.log(2 * np.pi)
    return dict(center=float(eta), disp=float(V), df=df, scale_sq=float(scale2), cov_factor=float(var), lml=float(ll))


class BaseConjugateProcess:
    """Stochastic process with a normal-inverse-chi^2 conjugate prior (gsum/models.py:29-900).

    Parameters are those of the reference: kernel, center, disp, df, scale, sd, basis, nugget, optimizer,
    n_restarts_optimizer, copy_X_train, random_state, decomposition.
    """

    _student = False

    def __init__(self, kernel=None, center=0, disp=0, df=1, scale=1, sd=None, basis=None, nugget=1e-10,
                 optimizer='fmin_l_bfgs_b', n_restarts_optimizer=0, copy_X_train=True, random_state=None,
                 decomposition='cholesky'):
        self.kernel = kernel
        self._center_0 = np.atleast_1d(center)
        self._disp_0 = np.atleast_2d(disp)
        if sd is not None:
            self._df_0, self._scale_0 = np.inf, sd
        else:
            self._df_0, self._scale_0 = df, scale
        self._fit = False
        self.X_train_ = self.y_train_ = None
        self.center_ = self.disp_ = self.df_ = self.scale_ = None
        self.cov_factor_ = self.cbar_sq_mean_ = None
        self.kernel_ = None
        self._rng = None
        self._handle = None
        self._eig = None                # ops.ResidentEigen of corr_ + nugget I (decomposition='eig')
        self._corr_L = self._corr = None
        self.nugget = nugget
        self.copy_X_train = copy_X_train
        self.random_state = random_state
        self.n_restarts_optimizer = n_restarts_optimizer
        self.optimizer = optimizer
        self.decomposition = decomposition
        self._default_kernel = ConstantKernel(1.0, constant_value_bounds='fixed') * RBF(1.0, length_scale_bounds='fixed')
        if basis is not None:
            raise NotImplementedError("gsum_b200: a custom `basis` is not supported (the reference itself only ever "
                                      "installs the constant basis, gsum/models.py:149-150)")
        self.basis = lambda X: np.ones((X.shape[0], 1))
        self.basis_train_ = None

    # ---- priors (gsum/models.py:153-167) ----
    @property
    def center0(self):
        return self._center_0

    @property
    def disp0(self):
        return self._disp_0

    @property
    def df0(self):
        return self._df_0

    @property
    def scale0(self):
        return self._scale_0

    def _priors(self):
        return dict(center0=_scalar_prior(self._center_0, "center"), disp0=_scalar_prior(self._disp_0, "disp"),
                    df0=float(self._df_0), scale0=float(self._scale_0))

    def _check_decomposition(self):
        if self.decomposition == 'cholesky':
            return
        if self.decomposition == 'eig':
            return
        raise ValueError('decomposition must be "cholesky" or "eig"')

    def _eig_route(self):
        self._check_decomposition()
        return self.decomposition == "eig"

    def _active_kernel(self):
        if self.kernel_ is not None:
            return self.kernel_
        return self._default_kernel if self.kernel is None else self.kernel

    # ---- fitted quantities ----
    @property
    def corr_L_(self):
        """Lower Cholesky factor of corr_ + nugget*I (fetched from the device on first access); on the 'eig' route the
        symmetric-eigenvector square root Q diag(sqrt(eig)) (gsum/models.py:717)."""
        if self._eig is not None:
            if self._corr_L is None:
                self._corr_L = self._eig.V * np.sqrt(self._eig.w)[None, :]
            return self._corr_L
        if self._handle is None:
            return None
        if self._corr_L is None:
            self._refit(want_L=True)
        return self._corr_L

    corr_sqrt_ = corr_L_

    @property
    def _eigh_tuple_(self):
        """(eig, Q) of corr_ + nugget*I on the 'eig' route (gsum/models.py:714-716), else None."""
        return None if self._eig is None else (self._eig.w, self._eig.V)

    @property
    def corr_(self):
        if self._handle is None and self._eig is None:
            return None
        if self._corr is None:
            k = flatten_kernel(self.kernel_)
            self._corr = ops.kernel_matrix(self.X_train_, None, k.ls_for(self.X_train_.shape[1]), k.constant, k.noise)
        return self._corr

    def center(self):
        """Posterior regression coefficients of the mean (gsum/models.py:505-516)."""
        return self.center_

    def disp(self):
        return self.disp_

    def df(self):
        return self.df_

    def scale(self):
        return self.scale_

    def mean(self, X):
        """MAP mean of the process at X; does not interpolate (gsum/models.py:551-560)."""
        center = self.center_ if self._fit else self.center0
        return self.basis(X) @ center

    def _cov_factor_and_kernel(self):
        if not self._fit:
            if self.df0 <= 2:
                raise ValueError('df must be greater than 2 for the covariance to exist')
            var = self.scale0 ** 2 if self.df0 == np.inf else self.df0 * self.scale0 ** 2 / (self.df0 - 2)
            return var, (self._default_kernel if self.kernel is None else self.kernel)
        return self.cov_factor_, self.kernel_

    def cov(self, X, Xp=None):
        """cov_factor * kernel(X, Xp) (gsum/models.py:562-599); not the conditional covariance."""
        var, kernel = self._cov_factor_and_kernel()
        X = np.atleast_2d(X)
        k = flatten_kernel(kernel)
        return ops.process_cov(X, Xp, k.ls_for(X.shape[1]), k.constant, k.noise, factor=var)

    def underlying_properties(self, X, return_std=False, return_cov=False):
        y_mean = self.mean(X)
        if return_cov:
            return y_mean, self.cov(X)
        if return_std:
            return y_mean, np.sqrt(np.diag(self.cov(X)))
        return y_mean

    # ---- fit (gsum/models.py:671-738) ----
    def _refit(self, want_L=False):
        k = flatten_kernel(self.kernel_)
        X = np.atleast_2d(self.X_train_)
        h = ops.FitHandle(X, self.y_train_, k.ls_for(X.shape[1]), k.constant, k.noise, self.nugget, student=self._student,
                          want_L=want_L, **self._priors())
        if self._handle is not None:
            self._handle.close()
        self._handle = h
        if want_L:
            self._corr_L = h.L
        return h

    def fit(self, X, y):
        """Fit to (X, y): calibrate the kernel (if it has free hyperparameters and an optimizer is set), factor the
        correlation matrix once on the device and update center/disp/df/scale."""
        self._check_decomposition()
        self.kernel_ = clone(self._default_kernel if self.kernel is None else self.kernel)
        self._rng = check_random_state(self.random_state)
        X, y = np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64)
        self.X_train_ = X.copy() if self.copy_X_train else X
        self.y_train_ = y.copy() if self.copy_X_train else y
        self.basis_train_ = self.basis(self.X_train_)
        self._corr_L = self._corr = None
        self._fit = False
        self._calibrate_kernel()
        if self._eig_route():
            return self._fit_eig()
        self._eig = None
        h = self._refit()
        self.center_ = np.array([h.center])
        self.disp_ = np.array([[h.disp]])
        self.df_ = h.df
        self.scale_ = h.scale
        self.cov_factor_ = self.cbar_sq_mean_ = h.cov_factor
        self.log_marginal_likelihood_value_ = h.lml if self._lml_from_optimizer is None else self._lml_from_optimizer
        self._fit = True
        return self

    # ---- decomposition='eig' (gsum/models.py:713-717, 480-484, 973-974, 1016-1019, 1215-1216, 1251-1253) ----
    def _eig_gram(self, X, y, kernel):
        """(ResidentEigen of R = kernel(X) + nugget I, Gram G = RHS^T R^-1 RHS of RHS = [basis | y], logdet R).

        R is built (K1), diagonalised (Jacobi) and applied (two DMMA GEMMs) on the device; the (n_c+1)^2 Gram of the
        returned R^-1 RHS against RHS is the only product formed on the host."""
        k = flatten_kernel(kernel)
        R = ops.kernel_matrix(X, None, k.ls_for(X.shape[1]), k.constant, k.noise)
        R[np.diag_indices_from(R)] += self.nugget
        eig = ops.ResidentEigen(R)
        rhs = np.concatenate([np.ones((X.shape[0], 1)), y], axis=1)
        G = rhs.T @ eig.solve(rhs)
        with np.errstate(invalid='ignore', divide='ignore'):
            logdet = float(np.sum(np.log(eig.w)))                                 # models.py:1019, 1253
        return eig, 0.5 * (G + G.T), logdet

    def _fit_eig(self):
        X, y = np.atleast_2d(self.X_train_), self.y_train_
        y2 = y[:, None] if y.ndim == 1 else y
        if self._handle is not None:
            self._handle.close()
            self._handle = None
        self._eig, G, logdet = self._eig_gram(X, y2, self.kernel_)
        post = _conjugate_from_gram(G, logdet, X.shape[0], y2.shape[1], self._priors(), self._student)
        self.center_ = np.array([post["center"]])
        self.disp_ = np.array([[post["disp"]]])
        self.df_ = post["df"]
        self.scale_ = np.sqrt(post["scale_sq"])
        self.cov_factor_ = self.cbar_sq_mean_ = post["cov_factor"]
        self.log_marginal_likelihood_value_ = post["lml"] if self._lml_from_optimizer is None else self._lml_from_optimizer
        self._fit = True
        return self

    def _lml_eig(self, theta, X, y):
        kernel = self._active_kernel().clone_with_theta(theta)
        X = self.X_train_ if X is None else X
        y = self.y_train_ if y is None else y
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        _, G, logdet = self._eig_gram(X, y, kernel)
        return float(_conjugate_from_gram(G, logdet, X.shape[0], y.shape[1], self._priors(), self._student)["lml"])

    def _predict_parts_eig(self, X, want, Xc, y, pred_noise, want_cond_basis):
        """`_predict_parts` on the 'eig' route: R_no R^-1 [y - m | B], and R_no R^-1 R_on (or its diagonal), from one
        device call on the resident eigendecomposition (gsum/models.py:797-845, 1150-1174)."""
        k = flatten_kernel(self.kernel_)
        d = X.shape[1]
        ls = k.ls_for(d)
        center = float(self.center_[0])
        if Xc is None:
            Xc, eig = np.atleast_2d(self.X_train_), self._eig
            if y is None:
                y = self.y_train_
        else:
            Xc = np.atleast_2d(np.asarray(Xc, dtype=np.float64))
            R = ops.kernel_matrix(Xc, None, ls, k.constant, k.noise)
            R[np.diag_indices_from(R)] += self.nugget
            eig = ops.ResidentEigen(R)                                            # models.py:811
            if y is None:
                y = self.y_train_
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        ny = y.shape[1]
        D = y - center
        if want_cond_basis:
            D = np.concatenate([D, np.ones((Xc.shape[0], 1))], axis=1)
        R_on = ops.kernel_matrix(Xc, X, ls, k.constant, 0.0)
        lin, qd, qc = eig.conditional(R_on, D, want_var=want == PREDICT_VAR, want_cov=want == PREDICT_COV)
        mean = center + lin[:, :ny]
        cond_basis = (1.0 - lin[:, ny]) if want_cond_basis else None
        extra = self.nugget if pred_noise else 0.0
        var = None
        if want == PREDICT_VAR:
            var = self.cov_factor_ * ((k.constant + k.noise + extra) - qd)
        elif want == PREDICT_COV:
            qc *= -1.0
            qc += ops.kernel_matrix(X, None, ls, k.constant, k.noise + extra)
            qc *= self.cov_factor_
            var = qc
        return mean, var, cond_basis

    def _calibrate_kernel(self):
        """gsum/models.py:630-669: L-BFGS on -log_marginal_likelihood with the analytic gradient (device contractions,
        _lml_gradient), for the Gaussian likelihood and the Student-t evidence alike.  (The ragged-array crash of
        models.py:664 on numpy >= 1.24 does not exist here.)"""
        self._lml_from_optimizer = None
        if self.optimizer is None or self.kernel_.n_dims == 0:
            return

        def obj_func(theta, eval_gradient=True):
            if not eval_gradient:
                return -self.log_marginal_likelihood(theta)
            lml, grad = self.log_marginal_likelihood(theta, eval_gradient=True)
            return -lml, -grad

        optima = [self._constrained_optimization(obj_func, self.kernel_.theta, self.kernel_.bounds)]
        if self.n_restarts_optimizer > 0:
            if not np.isfinite(self.kernel_.bounds).all():
                raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds are finite.")
            bounds = self.kernel_.bounds
            for _ in range(self.n_restarts_optimizer):
                theta_initial = self._rng.uniform(bounds[:, 0], bounds[:, 1])
                optima.append(self._constrained_optimization(obj_func, theta_initial, bounds))
        best = int(np.argmin([o[1] for o in optima]))
        self.kernel_.theta = optima[best][0]
        self._lml_from_optimizer = -optima[best][1]

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        """gsum/models.py:884-900."""
        if self.optimizer == "fmin_l_bfgs_b":
            theta_opt, func_min, info = fmin_l_bfgs_b(obj_func, initial_theta, bounds=bounds)
            if info["warnflag"] != 0:
                warnings.warn("fmin_l_bfgs_b terminated abnormally with the  state: %s" % info, ConvergenceWarning)
        elif callable(self.optimizer):
            theta_opt, func_min = self.optimizer(obj_func, initial_theta, bounds=bounds)
        else:
            raise ValueError("Unknown optimizer %s." % self.optimizer)
        return theta_opt, func_min

    # ---- likelihood (gsum/models.py:912-1057 / 1184-1273) ----
    def _lml_gradient(self, theta, X, y):
        """(log-likelihood, gradient): the Gaussian conjugate likelihood, gsum/models.py:957-1056 (eval_gradient=True), or
        the Student-t evidence, models.py:1199-1271.

        The device returns the Gram G = RHS^T R^-1 RHS of RHS = [basis | curves], the contractions H_p = Z^T dR_p Z
        (Z = R^-1 RHS) and t_p = trace(R^-1 dR_p) for the derivative of R with respect to each log-hyperparameter
        (ops.lml_grad_terms); every vector the reference contracts with dR is R^-1 (RHS a) for a small coefficient
        vector a, so its formulas become quadratic forms in G and H_p:
          compute_center  models.py:201-230,  compute_scale_sq  419-455,  compute_cov_factor  501-503,
          dK = var dR + dvar R  1024-1025,  0.5 (alpha alpha^T - K^-1) : dK - dmean^T alpha  1041-1056."""
        self._check_decomposition()
        theta = np.asarray(theta, dtype=np.float64)
        kernel = self._active_kernel().clone_with_theta(theta)
        X = self.X_train_ if X is None else X
        y = self.y_train_ if y is None else y
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        N, nc = y.shape
        k = flatten_kernel(kernel)
        layout = theta_layout(kernel, X.shape[1])
        if len(layout) != len(theta):
            raise ValueError("theta does not match the kernel's free hyperparameters")
        rhs = np.concatenate([np.ones((N, 1)), y], axis=1)
        G, H, tr, logdet, info = ops.lml_grad_terms(X, rhs, k.ls_for(X.shape[1]), constant=k.constant, noise=k.noise, nugget=self.nugget,
                                                    decomposition=self.decomposition)
        if info != 0:
            return -np.inf, np.zeros_like(theta)                                  # models.py:970-972
        pri = self._priors()
        eta0, V0, df0, scale0 = pri["center0"], pri["disp0"], pri["df0"], pri["scale0"]
        r = nc + 1
        eB = np.zeros(r); eB[0] = 1.0
        eavg = np.zeros(r); eavg[1:] = 1.0 / nc
        E = np.eye(r)[1:]                                                         # coefficient vectors of the curves
        quad_form = lambda M, a, b: a @ M @ b
        # posterior dispersion / center and d(center) (models.py:201-230, 260-277)
        if V0 == 0:
            V, eta = 0.0, eta0
            dcenter = np.zeros(len(H))
        else:
            V = 1.0 / (1.0 / V0 + nc * G[0, 0])
            eta = V * (eta0 / V0 + nc * quad_form(G, eB, eavg))
            dcenter = np.array([nc * V * quad_form(Hp, eB, eta * eB - eavg) for Hp in H])
        df = df0 + N * nc
        # scale^2 and its derivative (models.py:419-455)
        if np.isinf(df0):
            scale2, dscale2 = scale0 ** 2, np.zeros(len(H))
        else:
            Ec = E - eavg[None, :]                                                # centred curves
            quad = sum(quad_form(G, e, e) for e in Ec)
            ac = eavg - eta0 * eB                                                 # avg_y - B center0
            aw = nc * (ac - nc * V * quad_form(G, eB, ac) * eB)                   # mat_invR_avg_yc = Z aw
            quad2 = quad_form(G, ac, aw)
            scale2 = (df0 * scale0 ** 2 + quad + quad2) / df
            dscale2 = np.array([-(sum(quad_form(Hp, e, e) for e in Ec) + quad_form(Hp, aw, aw) / nc) / df for Hp in H])
        if self._student:
            # exact normal-inverse-chi^2 evidence and its gradient (models.py:1241-1271); compute_disp's derivative is
            # dV = n_c V (B^T R^-1 dR R^-1 B) V  (models.py:270-277)
            from scipy.special import loggamma

            def log_norm(df_, scale2_, disp_):
                norm = loggamma(df_ / 2.0) - df_ / 2.0 * np.log(df_ * scale2_ / 2.0)
                if disp_ > 0:
                    norm += 0.5 * np.log(2 * np.pi * disp_)
                return norm

            ll = log_norm(df, scale2, V) - log_norm(df0, scale0 ** 2, V0) - nc / 2.0 * (N * np.log(2 * np.pi) + logdet)
            grad_dev = -(nc / 2.0) * tr - (df / 2.0) * dscale2 / scale2
            if V != 0:
                dV = np.array([nc * V * V * quad_form(Hp, eB, eB) for Hp in H])
                grad_dev = grad_dev + 0.5 * dV / V
            grad = np.array([w * grad_dev[slot] for slot, w in layout])
            return float(ll), grad
        cov_factor = (lambda s2: s2) if np.isinf(df) else (lambda s2: df * s2 / (df - 2))
        var, dvar = cov_factor(scale2), cov_factor(dscale2)
        Bk = E - eta * eB[None, :]                                                # y_train_k = RHS b_k
        yKy = sum(quad_form(G, b, b) for b in Bk) / var                           # sum_k y_k^T K^-1 y_k
        ll = -0.5 * yKy - 0.5 * nc * (N * np.log(var) + logdet) - 0.5 * nc * N * np.log(2 * np.pi)
        grad_dev = np.empty(len(H))
        for p_, Hp in enumerate(H):
            aKa = sum(quad_form(Hp, b, b) for b in Bk) / var + dvar[p_] * yKy / var          # sum_k alpha^T dK alpha
            trK = tr[p_] + dvar[p_] * N / var                                                 # trace(K^-1 dK)
            dmean_alpha = dcenter[p_] * sum(quad_form(G, eB, b) for b in Bk) / var            # sum_k dmean^T alpha_k
            grad_dev[p_] = 0.5 * aKa - 0.5 * nc * trK - dmean_alpha
        grad = np.array([w * grad_dev[slot] for slot, w in layout])
        return float(ll), grad

    def _lml(self, theta, eval_gradient, X, y):
        if eval_gradient:
            return self._lml_gradient(theta, X, y)
        if self._eig_route():
            return self._lml_eig(theta, X, y)
        kernel = self._active_kernel().clone_with_theta(theta)
        X = self.X_train_ if X is None else X
        y = self.y_train_ if y is None else y
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        k = flatten_kernel(kernel)
        ls = k.ls_for(X.shape[1])
        ll = ops.lml_grid(X, y, 1.0, np.zeros(y.shape[1], dtype=np.int32), ls[None, :], np.ones(1), constant=k.constant,
                          noise=k.noise, nugget=self.nugget, student=self._student, **self._priors())
        return float(ll[0, 0])

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, X=None, y=None):
        raise NotImplementedError

    # ---- predict (gsum/models.py:753-845) ----
    def _predict_parts(self, X, want, Xc, y, pred_noise, want_cond_basis=False):
        """One device call: (mean (m, n_y), var|cov|None, conditional basis|None).

        Xc=None conditions at the training inputs with the factor kept on the device by `fit` (gsum/models.py:797-803);
        an explicit Xc is factored afresh with the nugget (models.py:806-809).  y=None means the y of `fit`."""
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        if self._eig_route():
            return self._predict_parts_eig(X, want, Xc, y, pred_noise, want_cond_basis)
        m = X.shape[0]
        center = float(self.center_[0])
        if Xc is not None and y is None:
            y = self.y_train_
        n_old = self.X_train_.shape[0] if Xc is None else np.atleast_2d(Xc).shape[0]
        return self._handle.predict(X, want=want, Xc=Xc, yc=y, mean_old=np.full(n_old, center), mean_new=np.full(m, center),
                                    basis_old=np.ones(n_old) if want_cond_basis else None,
                                    basis_new=np.ones(m) if want_cond_basis else None,
                                    pred_noise=pred_noise, want_cond_basis=want_cond_basis)

    def predict(self, X, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False):
        """Posterior mean / std / cov at X (gsum/models.py:753-845); the GP prior if `fit` has not been called."""
        if return_std and return_cov:
            raise RuntimeError('Only one of return_std or return_cov may be True')
        if not self._fit:
            return self.underlying_properties(X=X, return_std=return_std, return_cov=return_cov)
        self._check_decomposition()
        want = PREDICT_COV if return_cov else (PREDICT_VAR if return_std else PREDICT_MEAN)
        mean, var, _ = self._predict_parts(X, want, Xc, y, pred_noise)
        m_pred = np.squeeze(mean)
        if return_std:
            return m_pred, np.sqrt(var)
        if return_cov:
            return m_pred, np.squeeze(var)
        return m_pred

    def sample_y(self, X, n_samples=1, random_state=0, underlying=False):
        """Draws from the (posterior or underlying) process at X (gsum/models.py:847-879).

        The reference samples through numpy's SVD-based ``multivariate_normal``; here the covariance is factored
        by the device pivoted Cholesky (rank revealing, so the singular posterior covariance at training points is
        fine) and the draws are mean + G z on the device.  Same distribution, different random stream."""
        rng = check_random_state(random_state)
        if underlying:
            y_mean, y_cov = self.underlying_properties(X=X, return_cov=True)
        else:
            y_mean, y_cov = self.predict(X, return_cov=True)
        y_cov = np.atleast_2d(y_cov)
        n = y_cov.shape[0]
        _, Lp, piv, rank, _ = ops.pivoted_cholesky(y_cov)
        Lp[:, rank:] = 0.0
        inv = np.empty(n, dtype=np.int64)
        inv[piv] = np.arange(n)

        def draw(mean_vec):
            z = rng.standard_normal((n, n_samples))
            d, _ = ops.draws(Lp, np.zeros(n), Z=z)
            return mean_vec[:, None] + d[inv]

        if y_mean.ndim == 1:
            return draw(y_mean)
        return np.hstack([draw(y_mean[:, i])[:, np.newaxis] for i in range(y_mean.shape[1])])


class ConjugateGaussianProcess(BaseConjugateProcess):
    """Conjugacy-based Gaussian process (gsum/models.py:903-1087)."""

    _student = False

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, X=None, y=None):
        """Gaussian log-likelihood at the plug-in posterior-mean variance (gsum/models.py:912-1057)."""
        if theta is None and self._fit:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        return self._lml(theta, eval_gradient, X, y)


class ConjugateStudentProcess(BaseConjugateProcess):
    """Conjugacy-based Student-t process (gsum/models.py:1090-1273)."""

    _student = True

    def cov(self, X, Xp=None):
        """var * (corr + B V Bᵀ) (gsum/models.py:1099-1125)."""
        if not self._fit:
            df, scale, disp = self.df0, self.scale0, float(self.disp0[0, 0])
            kernel = self._default_kernel if self.kernel is None else self.kernel
        else:
            df, scale, disp, kernel = self.df_, self.scale_, float(self.disp_[0, 0]), self.kernel_
        if df <= 2:
            raise ValueError('df must be greater than 2 for the covariance to exist')
        var = scale ** 2 if df == np.inf else df * scale ** 2 / (df - 2)
        X = np.atleast_2d(X)
        k = flatten_kernel(kernel)
        return ops.process_cov(X, Xp, k.ls_for(X.shape[1]), k.constant, k.noise, factor=var, kernel_add=disp)

    def predict(self, X, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False):
        """gsum/models.py:1128-1182: the Gaussian prediction plus the mean-uncertainty term var * b V bᵀ with the
        conditional basis b = B* - R_*o R^-1 B (std is *added*, not combined in quadrature — reference behaviour)."""
        if return_std and return_cov:
            raise RuntimeError('Only one of return_std or return_cov may be True')
        if not self._fit:
            pred = self.underlying_properties(X=X, return_std=return_std, return_cov=return_cov)
            if not (return_std or return_cov):
                return pred
            disp = float(self.disp0[0, 0])
            var = self.scale0 ** 2 if self.df0 == np.inf else self.df0 * self.scale0 ** 2 / (self.df0 - 2)
            basis = np.ones(np.atleast_2d(X).shape[0])
        else:
            self._check_decomposition()
            want = PREDICT_COV if return_cov else (PREDICT_VAR if return_std else PREDICT_MEAN)
            mean, v, basis = self._predict_parts(X, want, Xc, y, pred_noise, want_cond_basis=return_std or return_cov)
            m_pred = np.squeeze(mean)
            if not (return_std or return_cov):
                return m_pred
            pred = (m_pred, np.sqrt(v)) if return_std else (m_pred, np.squeeze(v))
            disp, var = float(self.disp_[0, 0]), self.cov_factor_
        if return_std:
            return pred[0], pred[1] + np.sqrt(var * disp * basis * basis)
        return pred[0], pred[1] + var * disp * np.outer(basis, basis)

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, X=None, y=None):
        """Exact normal-inverse-chi^2 evidence (gsum/models.py:1184-1273).  NB: like the reference, this does not
        short-circuit on theta=None."""
        return self._lml(theta, eval_gradient, X, y)


class TruncationProcess:
    """EFT truncation-error model on top of a conjugate coefficient process (gsum/models.py:1284-1507).

    Parameters: kernel, ratio (scalar or callable), ref (scalar or callable), excluded, ratio_kws, and any
    BaseConjugateProcess keyword.
    """

    _process_class = BaseConjugateProcess

    def __init__(self, kernel=None, ratio=0.5, ref=1, excluded=None, ratio_kws=None, **kwargs):
        self.ref = ref if callable(ref) else (lambda X, ref=ref: ref * np.ones(X.shape[0]))
        self.ratio = ratio if callable(ratio) else (lambda X, ratio=ratio: ratio * np.ones(X.shape[0]))
        self.coeffs_process = self._process_class(kernel=kernel, **kwargs)
        self.kernel = kernel
        self.excluded = excluded
        self.ratio_kws = {} if ratio_kws is None else ratio_kws
        self._fit = False
        self.X_train_ = self.y_train_ = self.orders_ = None
        self.dX_ = self.dy_ = None
        self.coeffs_ = None

    # ---- scaled moments of the underlying process (gsum/models.py:1337-1365) ----
    def mean(self, X, start=0, end=np.inf):
        coeff_mean = self.coeffs_process.mean(X=X)
        ratio_sum = geometric_sum(x=self.ratio(X, **self.ratio_kws), start=start, end=end, excluded=self.excluded)
        return self.ref(X) * ratio_sum * coeff_mean

    def _kernel_add(self):
        cp = self.coeffs_process
        if not cp._student:
            return 0.0
        return float((cp.disp_ if cp._fit else cp.disp0)[0, 0])

    def cov(self, X, Xp=None, start=0, end=np.inf):
        cp = self.coeffs_process
        if cp._student:
            df = cp.df_ if cp._fit else cp.df0
            if df <= 2:
                raise ValueError('df must be greater than 2 for the covariance to exist')
            scale = cp.scale_ if cp._fit else cp.scale0
            var = scale ** 2 if df == np.inf else df * scale ** 2 / (df - 2)
            kernel = cp.kernel_ if cp._fit else (cp._default_kernel if cp.kernel is None else cp.kernel)
        else:
            var, kernel = cp._cov_factor_and_kernel()
        X = np.atleast_2d(X)
        k = flatten_kernel(kernel)
        q1, s1 = self.ratio(X, **self.ratio_kws), self.ref(X)
        q2 = s2 = None
        if Xp is not None:
            Xp = np.atleast_2d(Xp)
            q2, s2 = self.ratio(Xp, **self.ratio_kws), self.ref(Xp)
        return ops.process_cov(X, Xp, k.ls_for(X.shape[1]), k.constant, k.noise, factor=var, sc1=s1, sc2=s2, q1=q1, q2=q2,
                               gs_start=start, gs_end=end, excluded=self.excluded, kernel_add=self._kernel_add())

    def basis(self, X, start=0, end=np.inf):
        cn_basis = self.coeffs_process.basis(X=X)
        ratio_sum = geometric_sum(x=self.ratio(X, **self.ratio_kws)[:, None], start=start, end=end, excluded=self.excluded)
        return self.ref(X)[:, None] * ratio_sum * cn_basis

    def underlying_properties(self, X, order, return_std=False, return_cov=False):
        y_mean = self.mean(X, start=order + 1)
        if return_cov:
            return y_mean, self.cov(X, start=order + 1)
        if return_std:
            return y_mean, np.sqrt(np.diag(self.cov(X, start=order + 1)))
        return y_mean

    # ---- fit (gsum/models.py:1367-1387) ----
    def fit(self, X, y, orders, dX=None, dy=None):
        self.X_train_, self.y_train_, self.orders_ = X, y, orders
        # the grid caches are keyed by object identity: a later array may reuse the address of one that has been freed
        self._grid_inputs_cache = self._grid_ref_cache = None
        orders_mask = ~np.isin(orders, self.excluded)
        self.dX_, self.dy_ = dX, dy
        ratio, ref = self.ratio(X, **self.ratio_kws), self.ref(X)
        if np.atleast_1d(ratio).ndim > 1:
            raise ValueError('ratio must return a 1d array or a scalar')
        if np.atleast_1d(ref).ndim > 1:
            raise ValueError('ref must return a 1d array or a scalar')
        self.coeffs_ = coefficients(y=y, ratio=ratio, ref=ref, orders=orders)[:, orders_mask]
        self.coeffs_process.fit(X=X, y=self.coeffs_)
        self._fit = True
        return self

    # ---- predict (gsum/models.py:1389-1483) ----
    def _conditional(self, X, Xc, yc, start, end, want, want_cond_basis=False):
        """One scaled GP conditional on the device: returns (mean (m,), var|cov|None, cond_basis|None)."""
        cp = self.coeffs_process
        X, Xc = np.atleast_2d(X), np.atleast_2d(Xc)
        kw = self.ratio_kws
        q_old, q_new = self.ratio(Xc, **kw), self.ratio(X, **kw)
        s_old, s_new = self.ref(Xc), self.ref(X)
        b_old = b_new = None
        if want_cond_basis:
            b_old, b_new = self.basis(Xc, start=start, end=end)[:, 0], self.basis(X, start=start, end=end)[:, 0]
        try:
            if cp._handle is None:
                # 'eig' route: this predict is decomposition-independent in the reference (LU, models.py:1449).  The truncation branch
                # of gsum_predict builds its own factor of K_oo; the Cholesky fit here only provides the handle, and a training
                # correlation matrix without a Cholesky factor — the case the eig route exists for — goes to the eigen path below
                # (ADVICE r1), which needs nothing but the posterior the eig-route fit already holds.
                cp._refit()
            mean, var, cb = cp._handle.predict(
                X, want=want, Xc=Xc, yc=np.asarray(yc, dtype=np.float64), mean_old=self.mean(Xc, start=start, end=end),
                mean_new=self.mean(X, start=start, end=end), basis_old=b_old, basis_new=b_new, sc_old=s_old, sc_new=s_new,
                q_old=q_old, q_new=q_new, gs_start=start, gs_end=end, excluded=self.excluded, truncation=True,
                want_cond_basis=want_cond_basis, kernel_add=self._kernel_add())
            return mean[:, 0], var, cb
        except np.linalg.LinAlgError:
            # K_oo carries neither white noise nor a nugget here (gsum/models.py:1443-1449): beyond a few tens of smooth points
            # it is not numerically positive definite and has no Cholesky factor, while the reference's LU solve still returns
            # a result.  Same products through the symmetric eigendecomposition of K_oo on the device instead.
            return self._conditional_eig(X, Xc, yc, start, end, want, b_old, b_new)

    def _conditional_eig(self, X, Xc, yc, start, end, want, b_old, b_new):
        """`_conditional` for a K_oo without a Cholesky factor: K_no K_oo^-1 [y - m | basis | K_on] with
        K_oo^-1 = V diag(1/w) V^T from the device eigensolver (gsum_eigh + gsum_eig_conditional)."""
        K_oo = self.cov(Xc, Xc, start=start, end=end)
        K_on = np.ascontiguousarray(self.cov(Xc, X, start=start, end=end))
        eig = ops.ResidentEigen(K_oo)
        # eigenvalues at the rounding level of the largest one carry no information (they come out with either sign): they are
        # left out of the inverse — w = inf makes their 1/w vanish on the device — which is the pseudo-inverse the LU result
        # scatters around (measured on 60 points, l = 0.5: the data are reproduced to 2e-4 where the reference's LU gives 3e-3)
        w = eig.w.copy()
        w[w <= w.shape[0] * np.finfo(np.float64).eps * np.max(w)] = np.inf
        eig.w_dev.put(w)
        yc = np.asarray(yc, dtype=np.float64)
        D = (yc if yc.ndim == 2 else yc[:, None]) - self.mean(Xc, start=start, end=end)[:, None]
        n_y = D.shape[1]
        if b_old is not None:
            D = np.concatenate([D, np.asarray(b_old, dtype=np.float64)[:, None]], axis=1)
        lin, vterm, cterm = eig.conditional(K_on, np.ascontiguousarray(D), want_var=(want == PREDICT_VAR), want_cov=(want == PREDICT_COV))
        mean = self.mean(X, start=start, end=end) + lin[:, 0]
        cb = None if b_old is None else np.asarray(b_new, dtype=np.float64) - lin[:, n_y]
        var = None
        if want == PREDICT_VAR:
            var = self._prior_part(X, start, end, PREDICT_VAR)[1] - vterm
        elif want == PREDICT_COV:
            var = self.cov(X, X, start=start, end=end) - cterm
        return mean, var, cb

    def _prior_part(self, X, start, end, want):
        """Unconditioned truncation-error process: mean and (diagonal of the) covariance with Xp = X given
        explicitly, i.e. without white noise (gsum/models.py:1460-1461, 1474-1477)."""
        m = self.mean(X, start=start, end=end)
        if want == PREDICT_MEAN:
            return m, None
        if want == PREDICT_COV:
            return m, self.cov(X, Xp=X, start=start, end=end)
        cp = self.coeffs_process
        k = flatten_kernel(cp.kernel_)
        q, s = self.ratio(X, **self.ratio_kws), self.ref(X)
        gs = geometric_sum(q * q, start, end, self.excluded)
        return m, ((s * s) * gs) * (cp.cov_factor_ * (k.constant + self._kernel_add()))

    def predict(self, X, order, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False, kind='both'):
        """Interpolating GP of the order-`order` partial sum plus the truncation-error GP (gsum/models.py:1389-1483).

        As in the reference, K_oo carries no white noise / nugget here (kernel called with both arguments); the
        reference solves with LU, the device path with a Cholesky factor of the same symmetric matrix."""
        if not self._fit:
            return self.underlying_properties(X, order, return_cov=return_cov, return_std=return_std)
        if Xc is None:
            Xc = self.X_train_
        if y is None:
            if order not in self.orders_:
                raise ValueError('order must be in orders passed to `fit`')
            y = self.y_train_ if self.y_train_.ndim == 1 else np.squeeze(self.y_train_[:, self.orders_ == order])
        if kind not in ['both', 'interp', 'trunc']:
            raise ValueError('kind must be one of "both", "interp" or "trunc"')
        want = PREDICT_COV if return_cov else (PREDICT_VAR if return_std else PREDICT_MEAN)
        m_pred, K_pred = 0, 0
        if kind in ('both', 'interp'):
            m, K, _ = self._conditional(X, Xc, y, 0, order, want)
            m_pred = m_pred + m
            if K is not None:
                K_pred = K_pred + K
        if kind in ('both', 'trunc'):
            if self.dX_ is not None:
                m, K, _ = self._conditional(X, self.dX_, self.dy_, order + 1, np.inf, want)
            else:
                m, K = self._prior_part(X, order + 1, np.inf, want)
            m_pred = m_pred + m
            if K is not None:
                K_pred = K_pred + K
        if return_cov:
            return m_pred, K_pred
        if return_std:
            return m_pred, np.sqrt(K_pred)
        return m_pred

    # ---- likelihood (gsum/models.py:1485-1507) ----
    def _grid_inputs(self, X, y, orders):
        if X is None and y is None and orders is None:
            # the training data: converted once per fit (the grid is usually evaluated many times on the same data)
            key = (id(self.X_train_), id(self.y_train_), id(self.orders_))
            cached = getattr(self, "_grid_inputs_cache", None)
            if cached is None or cached[0] != key:
                cached = (key, self._grid_inputs(self.X_train_, self.y_train_, self.orders_))
                self._grid_inputs_cache = cached
            return cached[1]
        X = self.X_train_ if X is None else X
        y = self.y_train_ if y is None else y
        orders = self.orders_ if orders is None else orders
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        y = np.asarray(y, dtype=np.float64)
        orders = np.asarray(orders)
        if y.ndim != 2:
            raise ValueError('y must be 2d')
        if len(orders) != y.shape[-1]:
            raise ValueError('partials and orders must have the same length')
        mask = ~np.isin(orders, self.excluded)
        dy = np.ascontiguousarray(_order_differences(y)[:, mask])
        return X, dy, np.ascontiguousarray(orders[mask], dtype=np.int32)

    def log_marginal_likelihood(self, theta, eval_gradient=False, X=None, y=None, orders=None, **ratio_kws):
        """ll of the partial sums for kernel hyperparameters `theta` and ratio keyword(s) (gsum/models.py:1485-1507).

        Like the reference (models.py:1498-1507), only the scalar is returned even when eval_gradient=True is
        requested: the reference computes the coefficient process' gradient and drops it."""
        cp = self.coeffs_process
        X, dy, orders_in = self._grid_inputs(X, y, orders)
        kernel = cp._active_kernel().clone_with_theta(theta)
        k = flatten_kernel(kernel)
        ref = np.asarray(self.ref(X), dtype=np.float64)
        ratio = np.asarray(self.ratio(X, **ratio_kws), dtype=np.float64)
        det_factor = np.sum(len(orders_in) * np.log(np.abs(ref)) + np.sum(orders_in) * np.log(np.abs(ratio)))
        if cp._eig_route():
            # coefficients (gsum/helpers.py:71-101) on the host, then the coefficient process' 'eig' likelihood
            coeffs = dy / (ref[:, None] * ratio[:, None] ** orders_in[None, :])
            return cp._lml_eig(theta, X, coeffs) - float(det_factor)
        ll = ops.lml_grid(X, dy, ref, orders_in, k.ls_for(X.shape[1])[None, :], ratio[None, :], q_x_dependent=True,
                          detf=np.array([det_factor]), constant=k.constant, noise=k.noise, nugget=cp.nugget,
                          student=cp._student, **cp._priors())
        return float(ll[0, 0])

    def log_marginal_likelihood_grid(self, ls_vals, ratio_vals=None, ratio_kws_list=None, X=None, y=None, orders=None,
                                     group=None, return_status=False):
        """The whole (Q, l) likelihood surface in one device call — what the nested list comprehension of
        docs/notebooks/correlated_EFT_publication.ipynb cell 53 computes cell by cell.

        ls_vals : (n_ls,) or (n_ls, d) length scales (the kernel's other hyperparameters stay as they are).
        ratio_vals : (n_q,) scalar expansion parameters (each length scale is factored once and reused for every Q), or
        ratio_kws_list : list of keyword dicts for a callable `ratio` (x-dependent Q: one right-hand-side block per entry).
        group : optional torch.distributed process group; the length scales are then sharded round-robin over its
            ranks and the blocks all-gathered (see gsum_b200.distributed.lml_grid_sharded).
        Returns ll with shape (n_q, n_ls), i.e. ``ll[i_ratio][i_ls]`` as in the notebook.
        """
        cp = self.coeffs_process
        on_training_data = X is None and y is None and orders is None
        X, dy, orders_in = self._grid_inputs(X, y, orders)
        n = X.shape[0]
        k = flatten_kernel(cp._active_kernel())
        ls = np.asarray(ls_vals, dtype=np.float64)
        ls = ls.reshape(ls.shape[0], -1)
        if ls.shape[1] not in (1, X.shape[1]):
            raise ValueError("ls_vals must have shape (n_ls,) or (n_ls, n_features)")
        n_c, so = len(orders_in), float(np.sum(orders_in))
        # ref(X) and its log-Jacobian term on the training data: evaluated once per fit, like the order differences
        rc = getattr(self, "_grid_ref_cache", None) if on_training_data else None
        if rc is not None and rc[0] is X and rc[1] == n_c and rc[4] is self.ref:
            ref, ref_logsum = rc[2], rc[3]
        else:
            ref = np.asarray(self.ref(X), dtype=np.float64)
            ref_logsum = np.sum(n_c * np.log(np.abs(ref)))
            if on_training_data:
                self._grid_ref_cache = (X, n_c, ref, ref_logsum, self.ref)
        if (ratio_vals is None) == (ratio_kws_list is None):
            raise ValueError("give exactly one of ratio_vals and ratio_kws_list")
        if ratio_vals is not None:
            Q = np.asarray(ratio_vals, dtype=np.float64)
            detf = ref_logsum + n * so * np.log(np.abs(Q))
            xdep = False
        else:
            Q = np.stack([np.asarray(self.ratio(X, **kw), dtype=np.float64) for kw in ratio_kws_list])
            detf = ref_logsum + so * np.sum(np.log(np.abs(Q)), axis=1)
            xdep = True
        if cp._eig_route():
            detf_q = np.broadcast_to(detf, (Q.shape[0],))
            if group is not None:
                # same sharding as the Cholesky grid: length scales dealt round-robin, one all-gather
                from .distributed import lml_grid_sharded
                return lml_grid_sharded(X, dy, ref, orders_in, ls, Q, group=group,
                                        _evaluator=lambda X_, dy_, ref_, o_, ls_, Q_, **_kw:
                                        self._lml_grid_eig(X_, dy_, ref_, o_, ls_, Q_, xdep, detf_q, k))
            ll = self._lml_grid_eig(X, dy, ref, orders_in, ls, Q, xdep, detf_q, k)
            return (ll, np.zeros(ls.shape[0], dtype=np.int32)) if return_status else ll
        kw = dict(q_x_dependent=xdep, detf=detf, constant=k.constant, noise=k.noise, nugget=cp.nugget, student=cp._student,
                  **cp._priors())
        if group is not None:
            from .distributed import lml_grid_sharded
            return lml_grid_sharded(X, dy, ref, orders_in, ls, Q, group=group, **kw)
        return ops.lml_grid(X, dy, ref, orders_in, ls, Q, return_status=return_status, **kw)


    def _lml_grid_eig(self, X, dy, ref, orders_in, ls, Q, xdep, detf, k):
        """The (Q, l) surface on the 'eig' route: per length scale ONE device eigendecomposition and ONE device solve, reused
        for every Q — for a scalar Q the coefficients are dy / (ref Q^order) (gsum/helpers.py:71-101), so the Gram of
        [basis | c(Q)] is S(Q) G S(Q) with G the Gram of [basis | dy / ref] and S = diag(1, Q^-order); for an x-dependent
        Q all right-hand-side blocks go through the same solve.  The O(n_c^2) cell algebra is `_conjugate_from_gram`."""
        cp = self.coeffs_process
        n, nc = dy.shape
        n_q = Q.shape[0]
        pri = cp._priors()
        base = dy / ref[:, None]
        if xdep:
            blocks = [base / Q[i][:, None] ** orders_in[None, :] for i in range(n_q)]
            rhs = np.concatenate([np.ones((n, 1))] + blocks, axis=1)
        else:
            rhs = np.concatenate([np.ones((n, 1)), base], axis=1)
        ll = np.empty((n_q, ls.shape[0]))
        for j in range(ls.shape[0]):
            R = ops.kernel_matrix(X, None, ls[j], k.constant, k.noise)
            R[np.diag_indices_from(R)] += cp.nugget
            eig = ops.ResidentEigen(R)
            W = eig.solve(rhs)
            with np.errstate(invalid='ignore', divide='ignore'):
                logdet = float(np.sum(np.log(eig.w)))
            if not xdep:
                G0 = rhs.T @ W
                G0 = 0.5 * (G0 + G0.T)
            for i in range(n_q):
                if xdep:
                    cols = np.r_[0, 1 + i * nc:1 + (i + 1) * nc]
                    G = rhs[:, cols].T @ W[:, cols]
                    G = 0.5 * (G + G.T)
                else:
                    sc = np.r_[1.0, Q[i] ** (-orders_in.astype(np.float64))]
                    G = G0 * np.outer(sc, sc)
                ll[i, j] = _conjugate_from_gram(G, logdet, n, nc, pri, cp._student)["lml"] - detf[i]
        return ll


class TruncationGP(TruncationProcess):
    """Gaussian-process truncation model (gsum/models.py:1510-1516)."""
    _process_class = ConjugateGaussianProcess


class TruncationTP(TruncationProcess):
    """Student-t-process truncation model (gsum/models.py:1519-1570)."""
    _process_class = ConjugateStudentProcess

    def predict(self, X, order, return_std=False, return_cov=False, Xc=None, y=None, pred_noise=False, kind='both'):
        """gsum/models.py:1527-1570.  As in the reference, the Gaussian part is always evaluated with kind='both'
        (the reference does not forward `kind` to the parent) and the mean-uncertainty std is added linearly."""
        pred = super().predict(X=X, order=order, return_std=return_std, return_cov=return_cov, Xc=Xc, y=y,
                               pred_noise=pred_noise)
        if not return_std and not return_cov:
            return pred
        if Xc is None:
            Xc = self.X_train_
        cp = self.coeffs_process
        var, disp = cp.cov_factor_, float(cp.disp_[0, 0])
        m = np.atleast_2d(X).shape[0]
        basis_lower, basis_trunc = np.zeros(m), np.zeros(m)
        if kind in ('both', 'interp'):
            yy = np.zeros(np.atleast_2d(Xc).shape[0])
            _, _, basis_lower = self._conditional(X, Xc, yy, 0, order, PREDICT_MEAN, want_cond_basis=True)
        if kind in ('both', 'trunc'):
            if self.dX_ is not None:
                yy = np.zeros(np.atleast_2d(self.dX_).shape[0])
                _, _, basis_trunc = self._conditional(X, self.dX_, yy, order + 1, np.inf, PREDICT_MEAN, want_cond_basis=True)
            else:
                basis_trunc = self.basis(start=order + 1, end=np.inf, X=X)[:, 0]
        b = basis_lower + basis_trunc
        if return_std:
            return pred[0], pred[1] + np.sqrt(var * disp * b * b)
        return pred[0], pred[1] + var * disp * np.outer(b, b)
