"""SURVEY.md 8(f).4 on the device: TruncationPointwise (gsum/models.py:1573-1836) and VariogramFourthRoot
(gsum/helpers.py:525-730) against golden vectors produced by the real reference (tests/golden/make_golden_pointwise.py)."""
import os

import numpy as np
import pytest

import gsum_b200 as gb

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pointwise_variogram.npz"))
RTOL = 1e-10


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    scale = np.maximum(np.abs(b), 1e-300)
    assert np.all(both_nan | (np.abs(a - b) <= rtol * scale)), np.nanmax(np.abs(a - b) / scale)


@pytest.mark.parametrize("case", ["scalar", "xdep", "df0"])
def test_truncation_pointwise(case):
    p = "pw_" + case + "_"
    orders = G["pw_orders"]
    ratio, ref = G[p + "ratio"], G[p + "ref"]
    ratio = ratio[0] if ratio.size == 1 else ratio
    ref = ref[0] if ref.size == 1 else ref
    df0, scale0 = G[p + "prior"]
    excluded = G[p + "excluded"].tolist() or None
    tp = gb.TruncationPointwise(df=df0, scale=scale0, excluded=excluded).fit(G[p + "y"], ratio, ref, orders)
    close(tp.coeffs_, G[p + "coeffs"])
    assert tp.df_ == float(G[p + "df"])
    close(tp.scale_, G[p + "scale"])
    close(tp.dist_.kwds["scale"], G[p + "dist_scale"])
    close(tp.interval(G["pw_alpha"]), G[p + "interval"])
    close(tp.interval(G["pw_alpha"], orders=tp._orders_masked[-2:]), G[p + "interval_sel"])
    close(tp.pdf(G["pw_ygrid"]), G[p + "pdf"])
    close(tp.logpdf(G["pw_ygrid"], orders=tp._orders_masked[:1]), G[p + "logpdf"])
    close(tp.std(), G[p + "std"])
    close(tp.log_likelihood(), G[p + "ll"])
    grid = G["pw_ratio_grid"]
    close([tp.log_likelihood(ratio=q) for q in grid], G[p + "ll_grid"])
    close(tp.log_likelihood_grid(grid), G[p + "ll_grid"])                       # the scan in one device call
    n = G[p + "y"].shape[0]
    x = np.linspace(0.1, 1.0, n)
    ref_full = np.atleast_1d(ref) * np.ones(n)
    close(tp.log_likelihood_grid(grid[:, None] * (0.5 + x)[None, :], ref=ref_full), G[p + "ll_grid_x"])
    close(tp.credible_diagnostic(G[p + "data"], G["pw_dobs"]), G[p + "dci"], rtol=0)
    with pytest.raises(ValueError):
        gb.TruncationPointwise().log_likelihood()
    with pytest.raises(ValueError):
        gb.TruncationPointwise().fit(G[p + "y"], 0.5, 1.0, orders[:-1])


@pytest.mark.parametrize("case", ["1d", "2d"])
def test_variogram_fourth_root(case):
    p = "vg_" + case + "_"
    z = G[p + "z"]
    vg = gb.VariogramFourthRoot(G[p + "X"], z if z.shape[0] > 1 else z[0], G[p + "bounds"])
    assert np.array_equal(vg.bin_counts, G[p + "bin_counts"]) and np.array_equal(vg.bin_idx, G[p + "bin_idx"])   # exact
    close(vg.bin_locations, G[p + "bin_locations"])
    close(vg.gamma_star_hat, G[p + "gamma_star_hat"])
    close(vg.gamma_tilde, G[p + "gamma_tilde"])
    idx = G[p + "ijkl"]
    close(vg.rho_ijkl(*idx.T), G[p + "rho"])
    close(vg.corr_ijkl(*idx.T), G[p + "corr"], rtol=1e-9)
    close(vg.cov_ijkl(*idx.T), G[p + "cov_ijkl"], rtol=1e-9)
    nc = vg.Ncurves
    got = np.array([np.atleast_1d(vg.cov(b)) * np.ones(nc) for b in range(vg.Nb)])
    close(got, G[p + "cov_diag"], rtol=1e-9)                                     # sums of ~1e3..1e4 signed terms
    close(np.atleast_1d(vg.cov(1, 2)), G[p + "cov_01"], rtol=1e-9)
    for rt in (False, True):
        with np.errstate(invalid="ignore"):
            g, lo, up = vg.compute(rt_scale=rt)
        close(np.stack([g, lo, up]), G[p + f"compute_{int(rt)}"], rtol=1e-8)


def test_variogram_hypergeometric_series():
    """The device's F(z) = (1 - z) 2F1(3/4, 3/4; 1/2; z) against scipy over the whole range, through single pairs of pairs."""
    from scipy.special import hyp2f1
    from gsum_b200.variogram import _hyp_tables, _NT
    tab = _hyp_tables()
    z = np.concatenate([np.linspace(0, 0.999999, 401), 1 - np.logspace(-12, -1, 30)])

    def F(zz):                                    # the same two series on the host, to pin the tables themselves
        out = np.empty_like(zz)
        lo = zz <= 0.5
        out[lo] = np.polyval(tab[:_NT][::-1], zz[lo])
        w = 1 - zz[~lo]
        s = np.zeros_like(w)
        for n in range(_NT - 1, -1, -1):
            s = s * w + tab[_NT + n] * (np.log(w) + tab[2 * _NT + n])
        out[~lo] = tab[3 * _NT] + tab[3 * _NT + 1] * w * s
        return out
    ref = (1 - z) * hyp2f1(0.75, 0.75, 0.5, z)
    assert np.max(np.abs(F(z) - ref) / np.abs(ref)) < 1e-13


def test_pointwise_and_variogram_against_the_oracle_on_fresh_inputs():
    """Beyond the fixtures: seeded inputs of another size through the device classes and through the oracle's restatements
    (oracle/gsum_oracle.py: pointwise_*, VariogramOracle — themselves pinned on the real reference's outputs in tests/test_oracle.py)."""
    from oracle import gsum_oracle as o
    rs = np.random.RandomState(101)
    n, n_o = 300, 8
    orders = np.arange(n_o)
    x = np.linspace(0.05, 0.95, n)
    ratio, ref = 0.25 + 0.5 * x, 1.5 - x
    y = o.partials(rs.randn(n, n_o) * 0.8, ratio, ref, orders)
    tp = gb.TruncationPointwise(df=2.5, scale=1.3, excluded=[2]).fit(y, ratio, ref, orders)
    f = o.pointwise_fit(y, ratio, ref, orders, 2.5, 1.3, [2])
    close(tp.coeffs_, f["coeffs"])
    close(tp.scale_, f["scale"])
    close(tp.dist_.kwds["scale"], f["trunc_scale"])
    alpha = np.array([0.3, 0.9])
    close(tp.interval(alpha), o.pointwise_interval(f, alpha))
    yg = np.linspace(-1.0, 2.0, 5)
    close(tp.pdf(yg), o.pointwise_pdf(f, yg))
    close(tp.log_likelihood(), o.pointwise_log_likelihood(f))
    grid = np.linspace(0.3, 0.7, 9)
    close(tp.log_likelihood_grid(grid), [o.pointwise_log_likelihood(f, ratio=q) for q in grid])
    data = y[:, -1] + 0.2 * rs.randn(n)
    dobs = np.linspace(0.1, 0.9, 7)
    assert np.array_equal(tp.credible_diagnostic(data, dobs), o.pointwise_credible_diagnostic(f, data, dobs))
    # variogram: 60 points in 2-D, two curves
    X = rs.rand(60, 2)
    z = rs.randn(2, 60)
    bounds = np.linspace(0.15, 0.95, 7)
    vg, vo = gb.VariogramFourthRoot(X, z, bounds), o.VariogramOracle(X, z, bounds)
    assert np.array_equal(vg.bin_counts, vo.bin_counts) and np.array_equal(vg.bin_idx, vo.bin_idx)
    close(vg.gamma_star_hat, vo.gamma_star_hat)
    close(vg.gamma_tilde, vo.gamma_tilde)
    for b in (0, 2, 5):
        close(np.atleast_1d(vg.cov(b)), np.atleast_1d(vo.cov(b)), rtol=1e-8)
    close(np.atleast_1d(vg.cov(1, 3)), np.atleast_1d(vo.cov(1, 3)), rtol=1e-8)
