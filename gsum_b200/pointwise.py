"""`TruncationPointwise` — the pointwise convergence model of Furnstahl et al. (2015), drop-in for gsum/models.py:1573-1836.

The arithmetic of `fit` (coefficients, posterior scale, truncation-error scales) and of `log_likelihood` (the sums over
points) runs on the device through the C ABI (csrc/pointwise.cuh); the frozen `scipy.stats.t` objects `dist_` /
`coeffs_dist_` and what hangs off them (`interval`, `pdf`, `logpdf`, `std`) are the reference's own third-party calls and
are built from the device results.  `log_likelihood_grid` is additive: the Lambda_b-style scan over many expansion
parameters in one device call."""
from __future__ import annotations

import numpy as np
import scipy.stats as st
from scipy.special import loggamma

from . import ops

__all__ = ["TruncationPointwise"]


class TruncationPointwise:
    """y_k = y_ref sum_{n<=k} c_n Q^n with iid c_n | cbar^2 ~ N(0, cbar^2), cbar^2 ~ Inv-chi^2(df, scale^2).

    Parameters as in the reference: `df`, `scale` (prior hyperparameters), `excluded` (orders left out of the updating and
    of the truncation error)."""

    def __init__(self, df=1, scale=1, excluded=None):
        self.df0 = df
        self.scale0 = scale
        self.excluded = excluded

        self._fit = False
        self.y_ = None
        self.ratio_ = None
        self.ref_ = None
        self.orders_ = None
        self.orders_mask_ = None
        self._orders_masked = None
        self.coeffs_ = None
        self.coeffs_dist_ = None
        self.df_ = None
        self.scale_ = None
        self.y_masked_ = None
        self.dist_ = None

    @classmethod
    def _compute_df(cls, c, df0):
        return df0 + c.shape[-1]

    @staticmethod
    def _num_orders(y):
        if y.ndim == 1:
            return 1
        elif y.ndim == 2:
            return y.shape[-1]

    def _compute_order_indices(self, orders):
        if orders is None:
            return slice(None)
        orders = np.atleast_1d(orders)
        return np.squeeze([np.nonzero(self._orders_masked == order) for order in orders])

    def _excluded_array(self):
        return np.array([], dtype=np.int32) if self.excluded is None else np.atleast_1d(self.excluded).astype(np.int32)

    def fit(self, y, ratio, ref=1, orders=None):
        """gsum/models.py:1651-1690."""
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        ratio, ref = np.atleast_1d(ratio, ref)
        self.y_ = y
        self.ratio_ = ratio
        self.ref_ = ref
        if orders is None:
            orders = np.arange(y.shape[-1])
        orders = np.asarray(orders)
        if y.shape[-1] != orders.size:
            raise ValueError('The last dimension of `y` must have the same size as `orders`')
        self.orders_ = orders
        self.orders_mask_ = orders_mask = ~ np.isin(orders, self.excluded)
        n = y.shape[0]
        coeffs, scale, trunc_scale = ops.pointwise_fit(y, orders, orders_mask, self._excluded_array(), np.broadcast_to(ratio, (n,)),
                                                       np.broadcast_to(ref, (n,)), self.df0, self.scale0)
        self.coeffs_ = coeffs
        self.df_ = self._compute_df(c=coeffs, df0=self.df0)
        self.scale_ = scale
        self.y_masked_ = y[:, orders_mask]
        self._orders_masked = orders[orders_mask]
        self.coeffs_dist_ = st.t(loc=0, scale=self.scale_, df=self.df_)
        self.dist_ = st.t(loc=self.y_masked_, scale=trunc_scale, df=self.df_)
        self._fit = True
        return self

    def interval(self, alpha, orders=None):
        alpha = np.array(alpha)
        if alpha.ndim == 1:
            alpha = alpha[:, None, None]
        interval = np.array(self.dist_.interval(alpha))
        idx = self._compute_order_indices(orders)
        return interval[..., idx]

    def pdf(self, y, orders=None):
        y = np.atleast_1d(y)
        if y.ndim == 1:
            y = y[:, None, None]
        idx = self._compute_order_indices(orders)
        return self.dist_.pdf(y)[..., idx]

    def logpdf(self, y, orders=None):
        y = np.atleast_1d(y)
        if y.ndim == 1:
            y = y[:, None, None]
        idx = self._compute_order_indices(orders)
        return self.dist_.logpdf(y)[..., idx]

    def std(self):
        return self.dist_.std()

    def _loglike_from_sums(self, S1, S2):
        """Assemble gsum/models.py:1796-1803 from the two device sums (per ratio set)."""
        n = int(np.sum(self.orders_mask_))
        df0, scale0 = self.df0, self.scale0
        df = df0 + n
        log_like = loggamma(df / 2.) - 0.5 * n * np.log(2 * np.pi)
        if df0 > 0:  # the reference drops this infinite constant for the scale-invariant prior, df0 == 0
            log_like += 0.5 * np.sum(df0 * np.log(df0 * scale0 ** 2 / 2.)) - loggamma(df0 / 2.)
        return log_like - 0.5 * df * S1 - S2

    def log_likelihood(self, ratio=None, ref=None):
        """Log likelihood of ratio and ref given the data passed to `fit` (gsum/models.py:1762-1804), including the
        reference's broadcasting of the change-of-variables term (one term for a scalar ratio AND ref, n_points otherwise)."""
        if not self._fit:
            raise ValueError('Must call fit before calling log_likelihood')
        if ratio is None:
            ratio = self.ratio_
        if ref is None:
            ref = self.ref_
        ratio, ref = np.atleast_1d(np.asarray(ratio, dtype=np.float64)), np.atleast_1d(np.asarray(ref, dtype=np.float64))
        S1, S2 = ops.pointwise_loglike_sums(self.y_, self.orders_, self.orders_mask_, ratio[None, :], ref, self.df0, self.scale0)
        return self._loglike_from_sums(S1, S2)[0]

    def log_likelihood_grid(self, ratio_vals, ref=None):
        """`log_likelihood` for every row of `ratio_vals` ((n_r,) scalars or (n_r, n_points)) in one device call."""
        if not self._fit:
            raise ValueError('Must call fit before calling log_likelihood')
        if ref is None:
            ref = self.ref_
        ratio_vals = np.asarray(ratio_vals, dtype=np.float64)
        if ratio_vals.ndim == 1:
            ratio_vals = ratio_vals[:, None]
        S1, S2 = ops.pointwise_loglike_sums(self.y_, self.orders_, self.orders_mask_, ratio_vals, np.atleast_1d(ref), self.df0, self.scale0)
        return self._loglike_from_sums(S1, S2)

    def credible_diagnostic(self, data, dobs, band_intervals=None, band_dobs=None, beta=True):
        """gsum/models.py:1806-1836 (the `band_intervals` option needs the reference's legacy `hpd` helper for beta=True)."""
        dist = self.dist_
        dobs = np.atleast_1d(dobs)
        if data.ndim == 1:
            data = data[:, None]
        lower, upper = dist.interval(dobs[:, None, None])
        indicator = (lower < data) & (data < upper)
        D_CI = np.average(indicator, axis=1)
        if band_intervals is not None:
            if band_dobs is None:
                band_dobs = dobs
            band_dobs = np.atleast_1d(band_dobs)
            N = self.y_.shape[0]
            if beta:
                raise NotImplementedError("gsum_b200: the beta-HPD bands use the reference's legacy `hpd` helper "
                                          "(gsum/helpers.py:202-501), which is out of scope; pass beta=False")
            band_dist = st.binom(n=N, p=band_dobs)
            band_intervals = np.atleast_2d(band_intervals)
            bands = np.asarray(band_dist.interval(band_intervals.T)) / N
            bands = np.transpose(bands, [1, 0, 2])
            return D_CI, bands
        return D_CI
