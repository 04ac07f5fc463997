"""Dev tool: wall-clock of the other BASELINE.json configurations at full size through the public API (numpy in / out),
with the oracle (reference algorithm on the host) timed beside it on a bounded sample.

    python tools/perf_configs.py [c2] [c3] [c4x] [c5]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
import scipy.stats as st
import gsum_b200 as gb
from gsum_b200 import ops, _lib
from oracle import gsum_oracle as o

which = [a for a in sys.argv[1:] if a in ("c2", "c3", "c4x", "c5")] or ["c2", "c3", "c4x", "c5"]
ctx = _lib.default_context()

def timed(f, reps=3):
    f(); ctx.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r

if "c2" in which:
    rs = np.random.RandomState(1); n = 200
    X = np.linspace(0, 1, n)[:, None]
    coeffs = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, 6)
    orders = np.arange(6); y = o.partials(coeffs, 0.5, 1.0, orders)
    ls_vals, q_vals = np.linspace(0.02, 0.5, 64), np.linspace(0.3, 0.7, 64)
    kern = RBF(0.2) + WhiteKernel(1e-6, 'fixed')
    gp = gb.TruncationGP(kern, ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
    ms, ll = timed(lambda: gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals))
    t0 = time.perf_counter()
    ref = np.array([[o.truncation_lml(kern, [np.log(l)], X, y, orders, q * np.ones(n), np.ones(n), o.Priors(0, 0, 1, 1)) for l in ls_vals[::8]] for q in q_vals[::8]])
    cpu = (time.perf_counter() - t0) / 64 * 4096 * 1e3
    print(f"C2 64x64 grid N=200: device {ms:.2f} ms ({4096 / ms * 1e3:.3g} evals/s) | reference algorithm on host ~{cpu:.0f} ms (64 cells timed) | max rel err on those cells {np.max(np.abs(ll[::8, ::8] - ref) / np.abs(ref)):.2e}")

if "c3" in which:
    rs = np.random.RandomState(2)
    g1 = np.linspace(0, 1, 50); X = o.cartesian(g1, g1); n = len(X)
    Xt = rs.rand(10000, 2)
    kern = RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-6, 'fixed')
    Ktrue = RBF([0.02, 0.03])(X) + 1e-8 * np.eye(n)
    coeffs = np.linalg.cholesky(Ktrue) @ rs.randn(n, 6)
    orders = np.arange(6); y = o.partials(coeffs, 0.4, 1.0, orders)
    gp = gb.TruncationGP(kern, ratio=0.4, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
    ms_fit, _ = timed(lambda: gp.fit(X, y, orders=orders))
    ms_cp, (m1, s1) = timed(lambda: gp.coeffs_process.predict(Xt, return_std=True))
    ms_tp, (m2, s2) = timed(lambda: gp.predict(Xt, order=5, return_std=True, kind='both'))
    ms_cov, (m3, c3) = timed(lambda: gp.coeffs_process.predict(Xt[:4096], return_cov=True), reps=2)
    # oracle on a sub-sample of the test points
    f = o.fit_conjugate(kern, X, o.coefficients(y, 0.4, 1.0, orders), o.Priors(0, 0, 1, 1))
    t0 = time.perf_counter(); mr, sr = o.predict_conjugate(f, Xt[:1000], return_std=True); cpu_cp = time.perf_counter() - t0
    e_m = np.max(np.abs(m1[:1000] - mr)) / np.max(np.abs(mr)); e_s = np.max(np.abs(s1[:1000] - sr)) / np.max(np.abs(sr))
    flops_std = 1.0 * n * n * len(Xt)
    print(f"C3 N=2500 -> M=10000: fit {ms_fit:.1f} ms | coeffs_process.predict(std) {ms_cp:.1f} ms ({flops_std / ms_cp * 1e-9:.2f} TFLOP/s on the N^2 M forward solve) | "
          f"TruncationGP.predict(both, std) {ms_tp:.1f} ms | predict(cov) M=4096 {ms_cov:.1f} ms")
    print(f"   oracle predict(std) on 1000 of the points {cpu_cp * 1e3:.0f} ms (reference forms the M x M matrix: ~{cpu_cp * 1e3 * 100:.0f} ms extrapolated to 10000 by M^2) | rel err mean {e_m:.2e} std {e_s:.2e}")

if "c4x" in which:
    # second C4 variant of SURVEY.md 8(d): x-dependent Q(x) = q(x) / Lambda — one block of 6 right-hand sides per Lambda, so the
    # border of every factorisation carries 256 * 6 + 1 = 1537 rows (F_trsm = 128 * 1024^2 * 1537 = 206 GFLOP, GEMM shaped)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from util import c4_inputs
    X, y, orders = c4_inputs()
    n = len(X)
    ls_vals, lams = np.geomspace(0.005, 0.5, 128), np.linspace(0.8, 1.6, 256)
    kern = RBF(0.05) + WhiteKernel(1e-6, 'fixed')
    ratio = lambda X, lam: np.linspace(0.2, 0.6, len(X)) / lam
    gp = gb.TruncationGP(kern, ratio=ratio, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None, ratio_kws=dict(lam=1.0)).fit(X, y, orders=orders)
    kws = [dict(lam=l) for l in lams]
    ctx.profile(True)
    ms, ll = timed(lambda: gp.log_marginal_likelihood_grid(ls_vals, ratio_kws_list=kws))
    ctx.profile_read()
    gp.log_marginal_likelihood_grid(ls_vals, ratio_kws_list=kws); ctx.synchronize()
    fact_ms, fact_flops, n_br = ctx.profile_read()
    ctx.profile(False)
    t0 = time.perf_counter()
    cells = [(a, b) for a in (0, 100, 255) for b in (10, 64, 120)]
    ref = np.array([o.truncation_lml(kern, [np.log(ls_vals[b])], X, y, orders, ratio(X, lams[a]), np.ones(n), o.Priors(0, 0, 1, 1)) for a, b in cells])
    cpu = (time.perf_counter() - t0) / len(cells)
    got = np.array([ll[a, b] for a, b in cells])
    print(f"C4x 128 l x 256 Lambda, x-dependent Q, N=1024 (1537 border rows): grid {ms:.1f} ms through the numpy API ({128 * 256 / ms * 1e3:.3g} evals/s) | "
          f"factorisation + forward solves {fact_ms:.2f} ms over {n_br} brackets, {fact_flops / fact_ms * 1e-9:.2f} TFLOP/s algorithmic | "
          f"reference algorithm on host {cpu * 1e3:.0f} ms per cell ({1 / cpu:.2f} evals/s) | max rel err on {len(cells)} cells {np.max(np.abs(got - ref) / np.abs(ref)):.2e}")

if "c5" in which:
    n = 4096
    Xd = np.linspace(0, 1, n)[:, None]
    cov = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n)); mean = np.zeros(n)
    ms_init, d = timed(lambda: gb.Diagnostic(mean, cov, random_state=1), reps=2)
    ms_chol, ch = timed(lambda: ops.cholesky(cov), reps=2)
    ms_pc, pc = timed(lambda: ops.pivoted_cholesky(cov), reps=2)
    Y = d.samples(64)
    ms_md, md2 = timed(lambda: d.md_squared(Y))
    ms_pce, E = timed(lambda: d.pivoted_cholesky_errors(Y))
    iv = np.linspace(0, 1, 101)
    ms_draw, cv = timed(lambda: d.sample_coverage(100000, iv), reps=2)
    ms_cnt, cnt = timed(lambda: d.sample_coverage(100000, iv, counts=True, per_draw=False), reps=2)
    assert np.array_equal(cnt, np.rint(cv * n).astype(np.int64).sum(0))
    t0 = time.perf_counter(); Lr = np.linalg.cholesky(cov); cpu_chol = time.perf_counter() - t0
    from scipy.linalg.lapack import dpstrf
    t0 = time.perf_counter(); c_, p_, r_, i_ = dpstrf(cov, lower=True); cpu_pc = time.perf_counter() - t0
    piv_ok = np.array_equal(pc[2], p_ - 1)
    t0 = time.perf_counter(); z = np.random.RandomState(0).standard_normal((n, 2000)); Dr = Lr @ z; cpu_draw = time.perf_counter() - t0
    print(f"C5 N=4096: Diagnostic() {ms_init:.0f} ms (cholesky {ms_chol:.1f} ms incl. 2x128 MB PCIe; pivoted cholesky {ms_pc:.0f} ms, pivots == dpstrf: {piv_ok}) | md_squared(64) {ms_md:.1f} ms | pc_errors(64) {ms_pce:.1f} ms | "
          f"1e5 draws + coverage(101) {ms_draw:.0f} ms with the (1e5, 101) matrix copied back, {ms_cnt:.0f} ms counts only ({1.0 * n * n * 1e5 / ms_cnt * 1e-9:.1f} TFLOP/s triangular), max |coverage - nominal| {np.max(np.abs(cv.mean(0) - iv)):.4f}")
    print(f"   host: numpy cholesky {cpu_chol * 1e3:.0f} ms | LAPACK dpstrf {cpu_pc * 1e3:.0f} ms | L @ z for 2000 draws {cpu_draw * 1e3:.0f} ms (x50 for 1e5) | sum pc_err^2 vs md2 rel {np.max(np.abs((E ** 2).sum(0) - md2) / md2):.2e}")
