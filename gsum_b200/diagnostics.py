"""Model-checking diagnostics with the reference's `Diagnostic` API (gsum/diagnostics.py:21-194), Gaussian case.

Construction factors the covariance twice on the device (Cholesky and LAPACK-dpstrf-style pivoted
Cholesky); every method is then a forward solve, a batched `m + L z` draw or a coverage count on the GPU.
Out of scope here: `df=` (Student-t), `eigen_errors`, `kl`, `variogram`, plotting.
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
    df : must be None (Gaussian)
    random_state : int seed used by `samples`
    """

    def __init__(self, mean, cov, df=None, random_state=1):
        if df is not None:
            raise NotImplementedError("gsum_b200: Student-t diagnostics (df != None) are not implemented on the device path")
        self.mean = np.asarray(mean, dtype=np.float64)
        self.cov = np.asarray(cov, dtype=np.float64)
        self.sd = np.sqrt(np.diag(self.cov))
        self.random_state = random_state
        self.udist = stats.norm(loc=self.mean, scale=self.sd)       # interval end points only (host, O(N))
        self.std_udist = stats.norm(loc=0., scale=1.)
        self._chol = ops.cholesky(self.cov)                          # raises LinAlgError like numpy (diagnostics.py:60)
        G, Lp, piv, rank, status = ops.pivoted_cholesky(self.cov)    # diagnostics.py:61 -> helpers.py:185-199
        if status > 0:
            raise np.linalg.LinAlgError('M is not positive-semidefinite')
        self._pchol, self._pchol_L, self._piv = G, Lp, piv

    # -- draws -------------------------------------------------------------------------------------
    def samples(self, n, device_rng=False):
        """(n_samples, n) draws from N(mean, cov): mean + L z on the device (gsum/diagnostics.py:70-82).

        z comes from numpy's RandomState(random_state) (or the device Philox generator with `device_rng`);
        the reference's scipy/numpy SVD sampler uses a different factor, so streams differ by construction."""
        if device_rng:
            d, _ = ops.draws(self._chol, self.mean, n_draws=int(n), seed=int(self.random_state or 0))
            return d
        z = np.random.RandomState(self.random_state).standard_normal((self.mean.shape[0], int(n)))
        d, _ = ops.draws(self._chol, self.mean, Z=z)
        return d

    def sample_coverage(self, n, intervals, seed=None):
        """Coverage (n, n_intervals) of `n` fresh device draws, fused with the draw so the (N, n) sample matrix is
        never copied back (the GraphicalDiagnostic reference bands of gsum/diagnostics.py:557-584)."""
        lower, upper = self._bounds(intervals)
        seed = int(self.random_state or 0) if seed is None else int(seed)
        _, cov = ops.draws(self._chol, self.mean, n_draws=int(n), seed=seed, lower=lower, upper=upper, want_draws=False)
        return cov

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
        E, _ = ops.cholesky_errors(self._chol, self.mean, Y)
        return E[:, 0] if single else E

    def pivoted_cholesky_errors(self, y):
        """solve(G, y - mean) (gsum/diagnostics.py:103-104) via permutation + forward substitution."""
        Y, single = self._as_columns(y)
        E = ops.pc_errors(self._pchol_L, self._piv, self.mean, Y)
        return E[:, 0] if single else E

    def eigen_errors(self, y):
        raise NotImplementedError("gsum_b200: eigen_errors needs an eigensolver (SURVEY.md §8f item 2)")

    def chi2(self, y):
        return np.sum(self.individual_errors(y), axis=0)

    def md_squared(self, y):
        """Squared Mahalanobis distance of each curve (gsum/diagnostics.py:112-114)."""
        Y, single = self._as_columns(y)
        _, md2 = ops.cholesky_errors(self._chol, self.mean, Y, want_errors=False, want_md2=True)
        return md2[0] if single else md2

    def kl(self, mean, cov):
        raise NotImplementedError("gsum_b200: kl is not implemented on the device path")

    # -- credible intervals ------------------------------------------------------------------------
    def _bounds(self, intervals):
        lower, upper = self.udist.interval(np.atleast_2d(intervals).T)
        return np.ascontiguousarray(lower), np.ascontiguousarray(upper)

    def credible_interval(self, y, intervals):
        """Fraction of points of each curve inside each central credible interval (gsum/diagnostics.py:148-171).

        y : (n_samples, [n_curves]); returns ([n_curves], n_intervals)."""
        Y, single = self._as_columns(y)
        lower, upper = self._bounds(intervals)
        dci = ops.credible_interval(Y, lower, upper)
        return np.squeeze(dci) if single else dci
