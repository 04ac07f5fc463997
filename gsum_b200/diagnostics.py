"""Model-checking diagnostics with the reference's `Diagnostic` API (gsum/diagnostics.py:21-194), Gaussian case.

Construction factors the covariance twice on the device (Cholesky and LAPACK-dpstrf-style pivoted
Cholesky); every method is then a forward solve, a batched `m + L z` draw or a coverage count on the GPU.
`df=` selects the Student-t variant (gsum/diagnostics.py:51-55): multivariate-t draws are the Gaussian draws
scaled per draw by sqrt(df / chi2_df) and the interval end points come from `scipy.stats.t`.
`eigen_errors` runs on a device Jacobi eigendecomposition of `cov`, computed at its first call.
Out of scope here: `variogram`, plotting.
"""
from __future__ import annotations

import numpy as np
import scipy.stats as stats

from . import ops

__all__ = ["Diagnostic"]


class Diagnostic:
    """Diagnostics of curves y against N(mean, cov) (Bastos & O'Hagan).

    Parameters
    ----------
    mean : (n_samples,) array
    cov : (n_samples, n_samples) array
    df : None (Gaussian) or the degrees of freedom of a multivariate t with covariance `cov` (df > 2)
    random_state : int seed used by `samples`
    """

    def __init__(self, mean, cov, df=None, random_state=1):
        self.mean = np.asarray(mean, dtype=np.float64)
        self.cov = np.asarray(cov, dtype=np.float64)
        self.sd = np.sqrt(np.diag(self.cov))
        self.random_state = random_state
        self.df = None if df is None else float(df)
        if df is None:
            self.udist = stats.norm(loc=self.mean, scale=self.sd)   # interval end points only (host, O(N))
            self.std_udist = stats.norm(loc=0., scale=1.)
        else:
            # gsum/diagnostics.py:51-55: MVT(mean, sigma = cov (df - 2) / df, df); the marginals the reference uses for
            # the intervals are t(loc=mean, scale=sd, df) (scale sd, not sqrt(sigma_ii): mirrored as is)
            if not self.df > 2:
                raise ValueError("df must be greater than 2 for the covariance to exist")
            self.udist = stats.t(loc=self.mean, scale=self.sd, df=self.df)
            self.std_udist = stats.t(loc=0., scale=1., df=self.df)
        # one upload of cov; Cholesky (diagnostics.py:60) and pivoted Cholesky (diagnostics.py:61 -> helpers.py:185-199)
        # on the device copy; the factors stay in HBM for every later call
        self._eigen = None          # eigendecomposition for eigen_errors, computed on first use
        self._factors = f = ops.ResidentFactors(self.cov)
        if f.chol_info:
            raise np.linalg.LinAlgError("Matrix is not positive definite")        # numpy.linalg.cholesky
        if f.status > 0:
            raise np.linalg.LinAlgError('M is not positive-semidefinite')         # helpers.py:189-190

    # factors as numpy arrays (copied back from HBM on first access)
    _chol = property(lambda self: self._factors.chol)
    _pchol = property(lambda self: self._factors.pchol)
    _pchol_L = property(lambda self: self._factors.pchol_L)
    _piv = property(lambda self: self._factors.piv_host)

    # -- draws -------------------------------------------------------------------------------------
    def _draw_scale(self, n, rs):
        """Per-draw factor turning N(0, cov) draws into multivariate-t draws with scale matrix cov (df - 2) / df:
        sqrt((df - 2) / df) / sqrt(chi2_df / df)  (statsmodels `multivariate_t_rvs`: m + z / sqrt(x), x = chi2/df)."""
        if self.df is None:
            return None
        if np.isinf(self.df):
            return np.ones(int(n))
        x = rs.chisquare(self.df, int(n)) / self.df
        return np.sqrt((self.df - 2.0) / self.df) / np.sqrt(x)

    def samples(self, n, device_rng=False):
        """(n_samples, n) draws from N(mean, cov) — or the multivariate t — as mean + s L z on the device
        (gsum/diagnostics.py:70-82).

        z comes from numpy's RandomState(random_state) (or the device Philox generator with `device_rng`);
        the reference's scipy/numpy SVD sampler uses a different factor, so streams differ by construction."""
        rs = np.random.RandomState(self.random_state)
        if device_rng:
            d, _ = ops.draws(self._factors.L, self.mean, n_draws=int(n), seed=int(self.random_state or 0),
                             draw_scale=self._draw_scale(n, rs))
            return d
        z = rs.standard_normal((self.mean.shape[0], int(n)))
        d, _ = ops.draws(self._factors.L, self.mean, Z=z, draw_scale=self._draw_scale(n, rs))
        return d

    def sample_coverage(self, n, intervals, seed=None, first_draw=0, n_total=None, counts=False, per_draw=True):
        """Coverage (n, n_intervals) of `n` fresh device draws, fused with the draw so the (N, n) sample matrix is
        never copied back (the GraphicalDiagnostic reference bands of gsum/diagnostics.py:557-584).

        `first_draw` / `n_total` select draws first_draw .. first_draw + n - 1 of a run of n_total draws (one shard of
        the draw axis, gsum_b200.distributed.sample_coverage_sharded); with `counts` the int64 (n_intervals,) totals of
        (draw, point) pairs inside each interval are returned as well (alone, without the (n, n_intervals) matrix, when
        `per_draw` is False: nothing but n_intervals integers then leaves the device)."""
        lower, upper = self._bounds(intervals)
        seed = int(self.random_state or 0) if seed is None else int(seed)
        scale = None
        if self.df is not None:                                     # the whole run's chi-square stream, then this shard's slice
            total = int(first_draw) + int(n) if n_total is None else int(n_total)
            scale = self._draw_scale(total, np.random.RandomState(seed))[int(first_draw):int(first_draw) + int(n)]
        res = ops.draws(self._factors.L, self.mean, n_draws=int(n), seed=seed, lower=lower, upper=upper, want_draws=False,
                        first_draw=int(first_draw), draw_scale=scale, want_counts=counts, want_coverage=per_draw or not counts)
        if counts:
            return (res[1], res[2]) if per_draw else res[2]
        return res[1]

    # -- errors ------------------------------------------------------------------------------------
    def individual_errors(self, y):
        """(y - mean) / sd (gsum/diagnostics.py:84-98); elementwise."""
        return ((np.asarray(y).T - self.mean) / self.sd).T

    def _as_columns(self, y):
        y = np.asarray(y, dtype=np.float64)
        return (y[:, None], True) if y.ndim == 1 else (y, False)

    def cholesky_errors(self, y):
        """L^{-1}(y - mean) (gsum/diagnostics.py:100-101)."""
        Y, single = self._as_columns(y)
        E, _ = ops.cholesky_errors(self._factors.L, self.mean, Y)
        return E[:, 0] if single else E

    def pivoted_cholesky_errors(self, y):
        """solve(G, y - mean) (gsum/diagnostics.py:103-104) via permutation + forward substitution."""
        Y, single = self._as_columns(y)
        E = ops.pc_errors(self._factors.Lp, self._factors.piv, self.mean, Y)
        return E[:, 0] if single else E

    def eigen_errors(self, y):
        """solve(Q diag(sqrt(eig)), y - mean) with the eigenvalues ordered from largest to smallest
        (gsum/diagnostics.py:63-68, 106-107) = diag(eig^-1/2) Q^T (y - mean): the eigendecomposition of `cov` is
        computed on the device at the first call (Jacobi) and kept in HBM.  Each row carries the arbitrary sign of its
        eigenvector (numpy's `eigh` fixes none either)."""
        if self._eigen is None:
            self._eigen = ops.ResidentEigen(self.cov)
        Y, single = self._as_columns(y)
        E = self._eigen.solve(Y, mean=self.mean, mode=1)[::-1]
        return np.ascontiguousarray(E[:, 0] if single else E)

    @property
    def _eig(self):
        """Q diag(sqrt(eig)), largest eigenvalue first (gsum/diagnostics.py:63-68)."""
        if self._eigen is None:
            self._eigen = ops.ResidentEigen(self.cov)
        return (self._eigen.V * np.sqrt(self._eigen.w)[None, :])[:, ::-1]

    def chi2(self, y):
        return np.sum(self.individual_errors(y), axis=0)

    def md_squared(self, y):
        """Squared Mahalanobis distance of each curve (gsum/diagnostics.py:112-114)."""
        Y, single = self._as_columns(y)
        _, md2 = ops.cholesky_errors(self._factors.L, self.mean, Y, want_errors=False, want_md2=True)
        return md2[0] if single else md2

    def kl(self, mean, cov):
        """gsum/diagnostics.py:116-146, term by term as the reference evaluates them: tr(cov_1^{-1} cov_0) from a device
        cho_solve, the Mahalanobis term, and `logs = 2 sum(log diag(c1)) - logdet(c0)` — the reference takes the
        diagonal of the covariance c1 itself there (not of its factor), which is kept.  logdet(c0) comes from a device
        Cholesky of c0 (numpy's `slogdet` is LU based; for a covariance the two agree)."""
        c0 = np.asarray(cov, dtype=np.float64)
        tr = float(np.trace(ops.cho_solve(self._factors.L, c0)))
        dist = float(self.md_squared(np.asarray(mean, dtype=np.float64)))
        k = self.cov.shape[-1]
        _, info, logdet0 = ops.cholesky(c0, return_info=True)
        if info:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        logs = 2.0 * float(np.sum(np.log(np.diag(self.cov)))) - float(logdet0)
        return 0.5 * (tr + dist - k + logs)

    # -- credible intervals ------------------------------------------------------------------------
    def _bounds(self, intervals):
        """End points (n_intervals, n_samples) of the central credible intervals of the marginals — what the reference gets
        from `self.udist.interval(np.atleast_2d(intervals).T)` (gsum/diagnostics.py:161).  scipy evaluates
        `ppf(q) * scale + loc` element by element; the quantile depends on the level only, so it is computed once per level
        (n_intervals calls of the inverse CDF instead of n_intervals * n_samples) and expanded with the same multiply and
        add: bit-identical end points (tests/test_host_logic.py::test_interval_bounds_match_scipy)."""
        zl, zu = self.std_udist.interval(np.atleast_1d(np.asarray(intervals, dtype=np.float64)))
        lower = zl[:, None] * self.sd[None, :] + self.mean[None, :]
        upper = zu[:, None] * self.sd[None, :] + self.mean[None, :]
        return np.ascontiguousarray(lower), np.ascontiguousarray(upper)

    def credible_interval(self, y, intervals):
        """Fraction of points of each curve inside each central credible interval (gsum/diagnostics.py:148-171).

        y : (n_samples, [n_curves]); returns ([n_curves], n_intervals)."""
        Y, single = self._as_columns(y)
        lower, upper = self._bounds(intervals)
        dci = ops.credible_interval(Y, lower, upper)
        return np.squeeze(dci) if single else dci
