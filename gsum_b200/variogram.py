"""`VariogramFourthRoot` — empirical semivariogram with uncertainties via the fourth-root transformation, drop-in for
gsum/helpers.py:525-730 (Bowman & Crujeiras 2013; Cressie & Hawkins 1980).

The O(N^2) pair pass (distances, bins, sqrt|z_i - z_j|, per-bin sums) and the O(N^4) covariance of two bins (a sum over all
pairs of pairs of a correlation that needs 2F1(3/4, 3/4; 1/2; rho^2)) run on the device (csrc/pointwise.cuh)."""
from __future__ import annotations

import numpy as np
from scipy.special import gamma, digamma

from . import ops

__all__ = ["VariogramFourthRoot"]

_NT = 56


def _hyp_tables():
    """Series coefficients of F(z) = (1 - z) 2F1(3/4, 3/4; 1/2; z) = 2F1(-1/4, -1/4; 1/2; z): Maclaurin (z <= 1/2) and the
    logarithmic expansion in 1 - z (Abramowitz & Stegun 15.3.11 with m = 1); layout [a_n | A_n | B_n | K0, K1]."""
    a, A, B = np.empty(_NT), np.empty(_NT), np.empty(_NT)
    a[0] = A[0] = 1.0
    for n in range(_NT - 1):
        a[n + 1] = a[n] * (n - 0.25) ** 2 / ((n + 0.5) * (n + 1))
    for n in range(_NT):
        if n > 0:
            A[n] = A[n - 1] * (0.75 + n - 1) ** 2 / (n * (n + 1))
        B[n] = -digamma(n + 1) - digamma(n + 2) + 2 * digamma(0.75 + n)
    return np.concatenate([a, A, B, [np.sqrt(np.pi) / gamma(0.75) ** 2, gamma(0.5) / gamma(-0.25) ** 2]])


class VariogramFourthRoot:
    """X (n_samples, n_features), z (n_samples,) or (n_curves, n_samples), bin_bounds (n_bins - 1,) — as the reference."""

    mean_factor = np.sqrt(2 / np.pi) * gamma(0.75)
    var_factor = 2. / np.pi * (np.sqrt(np.pi) - gamma(0.75)**2)
    corr_factor = gamma(0.75)**2 / (np.sqrt(np.pi) - gamma(0.75)**2)
    _tab = None

    def __init__(self, X, z, bin_bounds):
        X = np.asarray(X, dtype=np.float64)
        bin_bounds = np.asarray(bin_bounds, dtype=np.float64)
        N = len(X)
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        Ncurves = z.shape[0]
        bin_grid, hij, bin_idx, dij, counts, hsum, dsum = ops.variogram_bins(X, z, bin_bounds)

        Nb = len(bin_bounds) + 1
        bin_labels = np.arange(Nb)
        gamma_star_hat = np.full((Nb, Ncurves), np.nan)
        # bin locations: midpoints of the boundaries, the overflow bins one bin length over (gsum/helpers.py:583-588),
        # moved to the average distance within the bin where the bin has data
        bin_locations = np.zeros(Nb)
        bin_locations[1:-1] = (bin_bounds[1:] + bin_bounds[:-1]) / 2
        bin_locations[0] = 2 * bin_bounds[0] - bin_locations[1]
        bin_locations[-1] = 2 * bin_bounds[-1] - bin_locations[-2]
        has = counts > 0
        bin_locations[has] = hsum[has] / counts[has]
        gamma_star_hat[has] = dsum[has] / counts[has][:, None]
        gamma_tilde = self.variogram_scale(gamma_star_hat)

        tri = np.tril_indices(N, -1)
        inputs = np.recarray((len(hij),), dtype=[('hij', float), ('bin_idxs', int), ('i', int), ('j', int)])
        inputs.hij, inputs.bin_idxs, inputs.i, inputs.j = hij, bin_idx, tri[0], tri[1]
        data = np.recarray((len(hij), Ncurves), dtype=[('dij', float), ('zi', float), ('zj', float)])
        data.dij, data.zi, data.zj = dij, z.T[tri[0]], z.T[tri[1]]

        self.N = N
        self.Nb = Nb
        self.Ncurves = Ncurves
        self.inputs = inputs
        self.data = data
        self.bin_idx = bin_idx.astype(int)
        self.bin_mask = bin_labels[:, None] == self.bin_idx
        self.bin_labels = bin_labels
        self.bin_counts = counts.astype(int)
        self.bin_locations = bin_locations
        self.gamma_star_hat = gamma_star_hat
        self.gamma_star_mean = self.mean_factor * gamma_star_hat
        self.gamma_tilde = gamma_tilde
        self._bin_grid = bin_grid
        if VariogramFourthRoot._tab is None:
            VariogramFourthRoot._tab = _hyp_tables()

    @property
    def gamma_tilde_grid(self):
        """gamma_tilde[bin of (i, j)]: (N, N, Ncurves), built on demand."""
        return self.gamma_tilde[self._bin_grid]

    def rho_ijkl(self, i, j, k, l):
        """Correlation between (Z_i - Z_j) and (Z_k - Z_l), estimated by gamma tilde (gsum/helpers.py:613-623)."""
        gam = self.gamma_tilde_grid
        return (gam[j, k] + gam[i, l] - gam[i, k] - gam[j, l]) / (2 * np.sqrt(gam[i, j] * gam[k, l]))

    def _pairs_cov(self, i, j, k, l, same_is_one=True):
        """cov_ijkl for explicit index lists through the device kernel: one (ij, kl) combination per call row."""
        out = np.empty((len(i), self.Ncurves))
        for c0 in range(0, self.Ncurves, 8):
            g = np.ascontiguousarray(self.gamma_tilde[:, c0:c0 + 8])
            for m in range(len(i)):
                out[m, c0:c0 + 8] = ops.variogram_cov([i[m]], [j[m]], [k[m]], [l[m]], self._bin_grid, g, self._tab,
                                                      self.var_factor, self.corr_factor, same_is_one)
        return out

    def var_ij(self, i, j):
        """Variance of sqrt|Z_i - Z_j|, estimated by gamma tilde (gsum/helpers.py:660-662)."""
        return self.var_factor * np.sqrt(self.gamma_tilde_grid[i, j])

    def cov_ijkl(self, i, j, k, l):
        """Covariance between sqrt|Z_i - Z_j| and sqrt|Z_k - Z_l| (gsum/helpers.py:645-658)."""
        i, j, k, l = np.atleast_1d(i, j, k, l)
        if not (i.shape == j.shape == k.shape == l.shape):
            raise ValueError(i.shape == j.shape == k.shape == l.shape, 'i, j, k, l must have the same shape')
        return self._pairs_cov(i, j, k, l)

    def corr_ijkl(self, i, j, k, l):
        """Correlation between sqrt|Z_i - Z_j| and sqrt|Z_k - Z_l| (gsum/helpers.py:625-643): the covariance over the two
        standard deviations; (i, j) == (k, l) is NOT special-cased there (the formula's value is returned)."""
        i, j, k, l = np.atleast_1d(i, j, k, l)
        return self._pairs_cov(i, j, k, l, same_is_one=False) / np.sqrt(self.var_ij(i, j) * self.var_ij(k, l))

    def cov(self, bin1, bin2=None):
        """Covariance of the fourth-root estimates of two bins (gsum/helpers.py:664-696) — the O(N^4) reduction, on the device."""
        nb1 = self.bin_counts[bin1]
        if bin2 is None or bin2 == bin1:
            bin2, nb2 = bin1, nb1
        else:
            nb2 = self.bin_counts[bin2]
        if (nb1 * nb2) == 0:
            return 0.
        m1, m2 = self.bin_mask[bin1], self.bin_mask[bin2]
        i1, j1, i2, j2 = self.inputs.i[m1], self.inputs.j[m1], self.inputs.i[m2], self.inputs.j[m2]
        out = np.empty(self.Ncurves)
        for c0 in range(0, self.Ncurves, 8):
            g = np.ascontiguousarray(self.gamma_tilde[:, c0:c0 + 8])
            out[c0:c0 + 8] = ops.variogram_cov(i1, j1, i2, j2, self._bin_grid, g, self._tab, self.var_factor, self.corr_factor)
        return out

    def variogram_scale(self, x):
        return (x / self.mean_factor) ** 4

    def fourth_root_scale(self, x):
        return self.mean_factor * x ** 0.25

    def compute(self, rt_scale=False):
        """Mean semivariogram and approximate 68% bands, on the 4th-root scale or the variogram scale (default)
        (gsum/helpers.py:704-730)."""
        gam = self.gamma_star_mean if rt_scale else self.gamma_tilde
        sd = np.zeros((self.Nb, self.Ncurves))
        for i in range(self.Nb):
            sd[i] = np.sqrt(self.cov(i))
        lower = self.gamma_star_mean - sd
        upper = self.gamma_star_mean + sd
        if not rt_scale:
            lower = self.variogram_scale(lower)
            upper = self.variogram_scale(upper)
        return gam, lower, upper
