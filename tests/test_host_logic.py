"""CPU: host-side logic of the facade — kernel flattening, series helpers, grid sharding over a gloo process group."""
import os
import sys

import numpy as np
import pytest
from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, DotProduct, Matern, WhiteKernel

from gsum_b200 import helpers
from gsum_b200.distributed import assemble_blocks, shard_indices
from gsum_b200.kernels import flatten_kernel
from oracle import gsum_oracle as o
from util import relerr


def test_flatten_supported_kernels():
    k = flatten_kernel(C(1.5) * RBF([0.2, 0.3]) + WhiteKernel(1e-6))
    assert k.constant == 1.5 and k.noise == 1e-6 and k.length_scale.tolist() == [0.2, 0.3]
    k = flatten_kernel(RBF(0.2))
    assert k.constant == 1.0 and k.noise == 0.0 and k.length_scale.tolist() == [0.2]
    k = flatten_kernel(WhiteKernel(1e-3) + RBF(0.4) * C(2.0) * C(3.0))
    assert k.constant == 6.0 and k.noise == 1e-3
    # theta round trip as the facade does it (models.py: clone_with_theta then flatten)
    kern = RBF(0.2) + WhiteKernel(1e-6, 'fixed')
    assert flatten_kernel(kern.clone_with_theta([np.log(0.37)])).length_scale[0] == pytest.approx(0.37)
    with pytest.raises(ValueError):
        k = flatten_kernel(RBF([0.1, 0.2, 0.3])); k.ls_for(2)


@pytest.mark.parametrize("kern", [Matern(0.3), DotProduct(), RBF(0.1) + RBF(0.2), RBF(0.1) * RBF(0.2), C(1.0) + RBF(0.2)])
def test_flatten_rejects_unsupported(kern):
    with pytest.raises(NotImplementedError):
        flatten_kernel(kern)


def test_series_helpers_match_oracle():
    rs = np.random.RandomState(1)
    y = rs.randn(30, 6).cumsum(1)
    q, ref, orders = 0.2 + 0.5 * rs.rand(30), 1 + rs.rand(30), np.array([0, 1, 3, 4, 5, 8])
    assert np.array_equal(helpers.coefficients(y, q, ref, orders), o.coefficients(y, q, ref, orders))
    assert np.array_equal(helpers.coefficients(y, 0.4, 2.0), o.coefficients(y, 0.4, 2.0))
    c = rs.randn(30, 6)
    assert np.array_equal(helpers.partials(c, q, ref, orders), o.partials(c, q, ref, orders))
    x = rs.rand(5, 4) * 0.9
    for s, e, ex in [(0, 3, None), (2, np.inf, None), (0, np.inf, [0, 2]), (1, 4, 3)]:
        assert np.array_equal(helpers.geometric_sum(x, s, e, ex), o.geometric_sum(x, s, e, ex))
    assert np.array_equal(helpers.cartesian(np.arange(3), np.arange(2)), o.cartesian(np.arange(3), np.arange(2)))
    with pytest.raises(ValueError):
        helpers.coefficients(y[:, 0], 0.5)
    with pytest.raises(ValueError):
        helpers.coefficients(y, 0.5, orders=np.arange(3))
    with pytest.raises(ValueError):
        helpers.geometric_sum(x, 4, 2)


def test_shard_indices_cover_grid_exactly_once():
    for n_ls in (1, 7, 128, 130):
        for world in (1, 2, 3, 8):
            seen = np.concatenate([shard_indices(n_ls, world, r) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(n_ls))
            per = -(-n_ls // world)
            blocks = []
            for r in range(world):
                idx = shard_indices(n_ls, world, r)
                b = np.full((3, per), -np.inf)
                b[:, :len(idx)] = idx[None, :] + 1000.0 * np.arange(3)[:, None]
                blocks.append(b)
            full = assemble_blocks(blocks, n_ls, world)
            assert np.array_equal(full, np.arange(n_ls)[None, :] + 1000.0 * np.arange(3)[:, None])


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gsum_b200.distributed import lml_grid_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def evaluator(X, dy, ref, orders, ls, Q, **kw):          # stand-in for the device grid: a known function of (Q, ls)
        return np.asarray(Q)[:, None] * 10.0 + np.asarray(ls)[:, 0][None, :]

    ls = np.linspace(0.1, 1.3, 13)
    Q = np.linspace(0.2, 0.8, 5)
    full = lml_grid_sharded(None, None, None, None, ls, Q, _evaluator=evaluator)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_grid_gloo_world2():
    """World-size-2 gloo run of the sharding + all-gather path (the NCCL run on GPUs uses the same code)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.linspace(0.2, 0.8, 5)[:, None] * 10.0 + np.linspace(0.1, 1.3, 13)[None, :]
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)


class _FakeDiagnostic:
    mean = np.zeros(7)


class _FakeProcess:
    """predict = a known function of the test points (stands in for a fitted replica)."""

    def predict(self, X, return_std=False, order=None):
        mean = np.stack([X[:, 0] * 2.0 + (order or 0), X[:, 0] ** 2], axis=1)
        return (mean, np.abs(X[:, 0]) + 0.5) if return_std else mean


def _draws_predict_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gsum_b200.distributed import predict_sharded, sample_coverage_sharded, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_draws, intervals = 11, np.linspace(0.1, 0.9, 4)

    def evaluator(lo, n):                     # counts a 1-GPU run would produce for draws lo .. lo + n - 1
        j = np.arange(lo, lo + n)
        return np.array([(j % (a + 2)).sum() for a in range(4)], dtype=np.int64)

    cov = sample_coverage_sharded(_FakeDiagnostic(), n_draws, intervals, _evaluator=evaluator)
    X = np.linspace(-1, 1, 9)[:, None]
    m, s = predict_sharded(_FakeProcess(), X, return_std=True, order=3)
    m_only = predict_sharded(_FakeProcess(), X[:1])                 # fewer points than ranks: rank 1 owns nothing
    q.put((rank, cov, m, s, m_only, shard_range(n_draws, world, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_draws_and_predict_gloo_world2():
    """World-size-2 gloo run of the draw-axis all-reduce and the test-point all-gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_draws_predict_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    j = np.arange(11)
    want_cov = np.array([(j % (a + 2)).sum() for a in range(4)], dtype=float) / (11 * 7)
    X = np.linspace(-1, 1, 9)
    assert res[0][4] == (0, 6) and res[1][4] == (6, 11)
    for r in range(2):
        cov, m, s, m_only, _ = res[r]
        assert np.array_equal(cov, want_cov)
        assert np.array_equal(m, np.stack([X * 2.0 + 3, X ** 2], axis=1)) and np.array_equal(s, np.abs(X) + 0.5)
        assert m_only.shape == (1, 2) and np.array_equal(m_only[0], [-2.0, 1.0])


@pytest.mark.parametrize("df", [None, 7.0])
def test_interval_bounds_match_scipy(df):
    """Diagnostic._bounds (one inverse-CDF call per level, expanded by multiply + add) is bit-identical to the reference's
    `udist.interval(np.atleast_2d(intervals).T)` (gsum/diagnostics.py:161), Gaussian and Student-t, end levels included."""
    import scipy.stats as st
    from gsum_b200.diagnostics import Diagnostic
    rs = np.random.RandomState(0)
    d = Diagnostic.__new__(Diagnostic)                      # host-side pieces only: no device factorisation here
    d.mean, d.sd = rs.randn(300), 0.5 + rs.rand(300)
    d.std_udist = st.norm(loc=0., scale=1.) if df is None else st.t(loc=0., scale=1., df=df)
    udist = st.norm(loc=d.mean, scale=d.sd) if df is None else st.t(loc=d.mean, scale=d.sd, df=df)
    iv = np.concatenate([np.linspace(0, 1, 101), [0.6827, 0.9545]])
    lower, upper = d._bounds(iv)
    want_lower, want_upper = udist.interval(np.atleast_2d(iv).T)
    assert np.array_equal(lower, want_lower) and np.array_equal(upper, want_upper)


def test_facade_rejects_out_of_scope_options():
    from gsum_b200 import ConjugateGaussianProcess, Diagnostic, TruncationGP
    with pytest.raises(NotImplementedError):
        ConjugateGaussianProcess(RBF(0.2), basis=lambda X: X)
    gp = ConjugateGaussianProcess(RBF(0.2, 'fixed'), decomposition='lu')
    with pytest.raises(ValueError):
        gp.fit(np.zeros((3, 1)), np.zeros(3))
    with pytest.raises(ValueError):
        Diagnostic(np.zeros(3), np.eye(3), df=2)                           # multivariate t without a covariance
    with pytest.raises(RuntimeError):
        ConjugateGaussianProcess(RBF(0.2, 'fixed')).predict(np.zeros((2, 1)), return_std=True, return_cov=True)
    with pytest.raises(ValueError):
        ConjugateGaussianProcess(RBF(0.2), df=1).cov(np.zeros((2, 1)))      # df <= 2: covariance does not exist


@pytest.mark.parametrize("student", [False, True])
@pytest.mark.parametrize("prior", [(0, 0, 1, 1), (0.3, 1, 3, 0.7), (0.1, 0, np.inf, 1.3), (-0.2, 0.5, 5, 2.0)])
def test_conjugate_from_gram_matches_oracle(student, prior):
    """The host half of the 'eig' route (gsum_b200.models._conjugate_from_gram: posterior and likelihood as quadratic forms
    in the Gram of R^-1 [1 | y]) against the oracle's step-by-step restatement of the reference (no device involved: the
    Gram is formed here with numpy)."""
    from gsum_b200.models import _conjugate_from_gram
    rs = np.random.RandomState(3)
    n, nc = 40, 4
    X = np.linspace(0, 1, n)[:, None]
    kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
    y = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, nc) + 0.2
    R = kern(X) + 1e-10 * np.eye(n)
    rhs = np.concatenate([np.ones((n, 1)), y], axis=1)
    G = rhs.T @ np.linalg.solve(R, rhs)
    logdet = np.linalg.slogdet(R)[1]
    center, disp, df, scale = prior
    pri = dict(center0=float(center), disp0=float(disp), df0=float(df), scale0=float(scale))
    got = _conjugate_from_gram(0.5 * (G + G.T), logdet, n, nc, pri, student)
    p = o.Priors(center, disp, df, scale)
    f = o.fit_conjugate(kern, X, y, p, nugget=1e-10, student=student)
    want = dict(center=f["center"][0], disp=f["disp"][0, 0], df=f["df"], scale_sq=f["scale"] ** 2, cov_factor=f["cov_factor"], lml=f["lml"])
    for key, w in want.items():
        if np.isnan(w):
            assert np.isnan(got[key]), key
        elif w == 0 or np.isinf(w):
            assert got[key] == w, key
        else:
            assert abs(got[key] - w) <= 1e-9 * abs(w), (key, got[key], w)


class _NumpyEigen:
    """Test stand-in for ops.ResidentEigen (numpy `eigh`): lets the HOST half of the 'eig' route — Gram assembly, closed
    forms, conditioning algebra, the per-length-scale reuse of the grid — run without a device."""

    def __init__(self, A, ctx=None):
        self.w, self.V = np.linalg.eigh(A)
        self.sweeps = 0

    def solve(self, Y, mean=None, mode=0):
        Y = np.asarray(Y, dtype=float)
        vec = Y.ndim == 1
        Y2 = Y[:, None] if vec else Y
        if mean is not None:
            Y2 = Y2 - mean[:, None]
        T = self.V.T @ Y2
        X = self.V @ (T / self.w[:, None]) if mode == 0 else T / np.sqrt(self.w)[:, None]
        return X[:, 0] if vec else X

    def conditional(self, R_on, D=None, want_var=False, want_cov=False):
        U = self.V.T @ R_on
        lin = U.T @ ((self.V.T @ D) / self.w[:, None]) if D is not None else None
        cov = U.T @ (U / self.w[:, None])
        return lin, (np.diag(cov).copy() if want_var else None), (cov if want_cov else None)


def _numpy_kernel_matrix(X1, X2, ls, constant=1.0, noise=0.0, ctx=None):
    X1, ls = np.atleast_2d(X1), np.atleast_1d(ls)
    k = RBF(ls if len(ls) > 1 else ls[0])
    if X2 is None:
        return constant * k(X1) + noise * np.eye(len(X1))
    return constant * k(X1, np.atleast_2d(X2))


@pytest.mark.parametrize("tag", ["g", "t"])
def test_eig_route_host_logic_against_reference_golden(monkeypatch, golden, tag):
    """fit / likelihood / predict / grid of the facade's 'eig' route with the two device calls replaced by numpy stand-ins,
    against the golden vectors of the real reference (the device kernels themselves are covered by tests/test_gpu_eig.py)."""
    import gsum_b200 as gb
    from gsum_b200 import ops
    from util import prior_kwargs
    monkeypatch.setattr(ops, "kernel_matrix", _numpy_kernel_matrix)
    monkeypatch.setattr(ops, "ResidentEigen", _NumpyEigen)
    g = golden("eig_route")
    cls = gb.ConjugateGaussianProcess if tag == "g" else gb.ConjugateStudentProcess
    for ip in range(3):
        pri = prior_kwargs(g["priors"][ip])
        kern = C(1.5, 'fixed') * RBF(0.2, 'fixed') + WhiteKernel(1e-4, 'fixed')
        gp = cls(kern, nugget=1e-10, decomposition='eig', **pri).fit(g["X"], g["y"])
        post = np.array([gp.center_[0], gp.disp_[0, 0], gp.df_, gp.scale_, gp.cov_factor_])
        want = g[f"{tag}{ip}_post"]
        ok = np.isfinite(want) & (want != 0)
        assert np.array_equal(np.isnan(post), np.isnan(want)) and np.max(np.abs(post[ok] - want[ok]) / np.abs(want[ok])) < 1e-9
        kfree = C(1.5, 'fixed') * RBF(0.2) + WhiteKernel(1e-4, 'fixed')
        gpf = cls(kfree, nugget=1e-10, optimizer=None, decomposition='eig', **pri).fit(g["X"], g["y"])
        lml = np.array([gpf.log_marginal_likelihood(theta=[t]) for t in g["thetas"]])
        wl = g[f"{tag}{ip}_lml"]
        assert np.array_equal(np.isnan(lml), np.isnan(wl))
        if np.isfinite(wl).all():
            assert relerr(lml, wl) < 1e-9
        if f"{tag}{ip}_mean" in g:
            m, s = gp.predict(g["Xn"], return_std=True)
            assert relerr(m, g[f"{tag}{ip}_mean"]) < 1e-9 and relerr(s, g[f"{tag}{ip}_std"]) < 1e-6
            _, cv = gp.predict(g["Xn"][::4], return_cov=True, pred_noise=True)
            assert relerr(cv, g[f"{tag}{ip}_cov"]) < 2e-6
            m, s = gp.predict(g["Xn"], return_std=True, Xc=g["Xc"], y=g["yc"])
            assert relerr(m, g[f"{tag}{ip}_mean_c"]) < 1e-9 and relerr(s, g[f"{tag}{ip}_std_c"]) < 1e-6
    if tag == "g":
        tgp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1,
                              optimizer=None, decomposition='eig').fit(g["Xt"], g["yt"], orders=g["orders"])
        grid = tgp.log_marginal_likelihood_grid(g["ls_vals"], g["q_vals"])
        assert relerr(grid, g["t_ll"]) < 1e-7
        cell = tgp.log_marginal_likelihood(theta=[np.log(g["ls_vals"][1])], ratio=g["q_vals"][2])
        assert abs(cell - g["t_ll"][2, 1]) < 1e-7 * abs(g["t_ll"][2, 1])
        tq = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=lambda X, q=0.5: q * np.ones(len(X)), ref=1, center=0,
                             disp=0, df=1, scale=1, optimizer=None, decomposition='eig').fit(g["Xt"], g["yt"], orders=g["orders"])
        gx = tq.log_marginal_likelihood_grid(g["ls_vals"], ratio_kws_list=[dict(q=q) for q in g["q_vals"]])
        assert relerr(gx, g["t_ll"]) < 1e-7


def _jacobi_pair(np_, r, k):
    """Python mirror of `jacobi_pair` in gsum_b200/csrc/eig.cuh (round-robin / circle method)."""
    m = np_ - 1
    p, q = (m, r) if k == 0 else ((r + k) % m, (r - k + m) % m)
    return (p, q) if p < q else (q, p)


@pytest.mark.parametrize("n", [2, 3, 4, 7, 16, 33])
def test_jacobi_schedules_cover_every_pair_once_with_disjoint_rounds(n):
    """The two pair schedules of gsum_eigh (eig.cuh: round-robin tournament; `jacobi_select`'s modulus ordering i + j = s
    mod n): every round's pairs are disjoint (CTAs of one launch never share a row) and a sweep visits each pair once."""
    want = {(i, j) for i in range(n) for j in range(i + 1, n)}
    np_ = n + (n & 1)
    seen = []
    for r in range(np_ - 1):
        pairs = [_jacobi_pair(np_, r, k) for k in range(np_ // 2)]
        pairs = [pq for pq in pairs if pq[1] < n]                     # the padding index of an odd n sits out
        flat = [x for pq in pairs for x in pq]
        assert len(flat) == len(set(flat))
        seen += pairs
    assert sorted(seen) == sorted(want)
    seen = []
    for s in range(n):
        pairs = [(i, (s - i + n) % n) for i in range(n) if i < (s - i + n) % n]
        flat = [x for pq in pairs for x in pq]
        assert len(flat) == len(set(flat))
        seen += pairs
    assert sorted(seen) == sorted(want)


def test_factorisation_claim_lists_are_topologically_ordered(tmp_path):
    """The deadlock-freedom argument of chol_hetero_tma_kernel (csrc/hetero.cuh): workers claim in list order and a task waits only
    for tasks EARLIER in the joint order.  tests/native/task_order_check.cu rebuilds the host-side lists (many-matrices schedule with
    and without thin border rows / diagonal delay, and the chain-mode lists with their `pre` tasks) for 270 shapes and checks exactly
    that, plus: every tile once, column-0 diagonal tiles at the head of the factor list, the split preserves the order."""
    import shutil
    import subprocess
    nvcc = shutil.which(os.environ.get("NVCC", "nvcc"))
    if nvcc is None:
        pytest.skip("nvcc not on PATH")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "task_order_check")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                        os.path.join(here, "native", "task_order_check.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "0 violations" in r.stdout, r.stdout[-2000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver divides by) runs without a GPU and prints ONE JSON line with the contract's
    keys: same metric / unit / config as the GPU arm, `impl: reference`, a `cpu_baseline` describing the run and a zero-copy `e2e`."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("(l,Q) grid log-likelihood evals/sec at N=1024") and d["config"]["workload"].startswith("C4")
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1 BLAS thread" in d["cpu_baseline"]["sample"]


# ---- free helpers and data generators either side of the path (gsum/helpers.py:202-368, gsum/datasets.py) ---------------------
def test_host_only_helpers_match_reference_golden(golden):
    """`stabilize`, `predictions`, `hpd`, `hpd_pdf`, `median_pdf`, `rbf(ls=0)`: no device work; outputs of the real reference."""
    import scipy.stats as st
    import gsum_b200 as gb
    g = golden("helpers_datasets")
    assert np.array_equal(gb.stabilize(g["stabilize_in"]), g["stabilize_out"])
    x, pdf, alphas = g["pdf_x"], g["pdf_vals"], g["pdf_alphas"]
    assert np.array_equal(np.array([gb.hpd_pdf(pdf, a, x) for a in alphas]), g["hpd_pdf"])
    assert gb.median_pdf(pdf, x) == float(g["median_pdf"])
    assert gb.median_pdf(0.01 * np.ones(5), np.arange(5.0)) == 4.0                       # never reaches 1/2: the last grid point
    assert np.allclose(np.array([gb.hpd(st.norm(0.3, 1.1), a) for a in alphas]), g["hpd_norm"], rtol=0, atol=1e-12)
    assert np.allclose(np.array([gb.hpd(st.gamma, a, 2.0) for a in alphas]), g["hpd_gamma"], rtol=0, atol=1e-12)
    assert np.allclose(np.array([gb.hpd(st.t(4.5, loc=-1.0, scale=0.6), a) for a in alphas]), g["hpd_t"], rtol=0, atol=1e-12)
    dist = st.norm(np.linspace(-1, 1, 7), np.linspace(0.5, 2.0, 7))
    assert np.array_equal(gb.predictions(dist), g["pred_mean"])
    m, iv = gb.predictions(dist, dob=[0.68, 0.95])
    assert np.array_equal(m, g["pred_mean"]) and np.array_equal(iv, g["pred_intervals"])
    assert np.array_equal(gb.predictions(dist, dob=0.5)[1], g["pred_interval_single"])
    assert np.array_equal(gb.rbf(g["rbf_ls0_X"], ls=0), g["rbf_ls0"])
    with pytest.raises(ValueError):
        gb.kl_gauss(0.0, 1.0, 0.0)                                                        # neither cov1 nor chol1
    with pytest.raises(ValueError):
        gb.kl_gauss(0.0, 1.0, 0.0, cov1=1.0, chol1=1.0)


def _numpy_pivoted_cholesky(M):
    """Stand-in for ops.pivoted_cholesky: LAPACK dpstrf itself, (G, Lp, piv, rank, status) with the device call's conventions."""
    from scipy.linalg.lapack import dpstrf
    c, p, rank, info = dpstrf(np.array(M, dtype=float), lower=1)
    Lp, piv = np.tril(c), (p - 1).astype(np.int32)
    G = np.empty_like(Lp)
    G[piv] = np.where(np.arange(len(piv))[None, :] < rank, Lp, 0.0)
    return G, Lp, piv, int(rank), int(info)


def _numpy_draws(L, mean, Z=None, **kw):
    return np.asarray(mean)[:, None] + np.tril(L) @ Z, None


def test_partial_sum_generators_host_logic(monkeypatch, golden):
    """The generators with the three device calls replaced by numpy stand-ins: shapes, the order bookkeeping, the un-pivoting
    of the factor (the draws must have covariance K, not P^T K P) and the inputs of the `_uniform` / `_on_grid` variants.
    The device calls themselves: tests/test_gpu_helpers.py."""
    import gsum_b200 as gb
    from gsum_b200 import datasets as ds, ops
    g = golden("helpers_datasets")
    monkeypatch.setattr(ops, "kernel_matrix", _numpy_kernel_matrix)
    monkeypatch.setattr(ops, "pivoted_cholesky", _numpy_pivoted_cholesky)
    monkeypatch.setattr(ops, "draws", _numpy_draws)
    kern = C(1.5) * RBF(0.25) + WhiteKernel(1e-3)
    X, K = g["ds_X"], g["ds_K"]
    n_draw = 4000
    big = gb.make_gaussian_partial_sums(X, orders=n_draw, kernel=kern, mean=lambda X: 0.5 * np.ones(len(X)), ratio=1.0, ref=1.0,
                                        nugget=1e-4, random_state=5)
    coeffs = gb.coefficients(big, 1.0, 1.0)
    sd = np.sqrt((K ** 2 + np.outer(np.diag(K), np.diag(K))) / (n_draw - 1))
    assert np.max(np.abs(np.cov(coeffs) - K) / sd) < 5.0                                   # same bound the reference's draws meet
    assert np.max(np.abs(coeffs.mean(axis=1) - 0.5) / np.sqrt(np.diag(K) / n_draw)) < 5.0
    ratio_fn, ref_fn = (lambda X: 0.3 + 0.2 * X[:, 0]), (lambda X: 2.0 - X[:, 0])
    y = gb.make_gaussian_partial_sums(X, orders=g["ds_orders"], kernel=kern, ratio=ratio_fn, ref=ref_fn, nugget=1e-4, random_state=5)
    assert y.shape == tuple(g["ds_y_shape"])
    y2 = gb.make_gaussian_partial_sums(X, orders=g["ds_orders"], kernel=kern, ratio=ratio_fn, ref=ref_fn, nugget=1e-4, random_state=5)
    assert np.array_equal(y, y2)                                                           # seeded: reproducible
    c = gb.coefficients(y, ratio_fn(X), ref_fn(X), g["ds_orders"])
    z = np.random.RandomState(5).standard_normal((len(X), 4))
    G = o.pivoted_cholesky(K)
    assert relerr(c, G @ z) < 1e-9                                                         # coefficients = G z for the seed's normals
    Xu, yu = gb.make_gaussian_partial_sums_uniform(n_samples=12, n_features=2, orders=3, random_state=9)
    assert np.array_equal(Xu, g["ds_uniform_X"]) and yu.shape == tuple(g["ds_uniform_y_shape"])
    Xg, yg = gb.make_gaussian_partial_sums_on_grid(n_samples=9, n_features=1, orders=4, random_state=9)
    assert np.array_equal(Xg, g["ds_grid_X"]) and yg.shape == tuple(g["ds_grid_y_shape"])
    Xg2, yg2 = gb.make_gaussian_partial_sums_on_grid(n_samples=5, n_features=2, orders=2, nugget=1e-6)
    assert Xg2.shape == (25, 2) and yg2.shape == (25, 2) and np.array_equal(Xg2, gb.cartesian(np.linspace(0, 1, 5), np.linspace(0, 1, 5)))
    # a singular covariance (two coinciding inputs, no noise): accepted by default, scipy's LinAlgError when not allowed
    Xs = np.array([[0.0], [0.5], [0.5], [1.0]])
    ys = gb.make_gaussian_partial_sums(Xs, orders=3, kernel=RBF(0.3))
    assert ys.shape == (4, 3) and np.allclose(ys[1], ys[2], rtol=0, atol=1e-7)
    with pytest.raises(np.linalg.LinAlgError):
        gb.make_gaussian_partial_sums(Xs, orders=3, kernel=RBF(0.3), allow_singular=False)
    with pytest.raises(NotImplementedError):
        gb.make_gaussian_partial_sums(X, kernel=Matern(0.3))
    with pytest.raises(ValueError):
        ds.make_gaussian_partial_sums(np.linspace(0, 1, 5))                               # X must be 2d


def test_device_backed_helpers_host_logic(monkeypatch, golden):
    """`gaussian` / `rbf` / `kl_gauss` with the device calls replaced by numpy stand-ins, against the real reference's outputs:
    the argument handling (the un-rescaled Xp of `gaussian`, the broadcast prior mean and the `stabilize`d cov1 of `kl_gauss`)."""
    import scipy.linalg as sl
    import gsum_b200 as gb
    from gsum_b200 import ops
    g = golden("helpers_datasets")
    monkeypatch.setattr(ops, "kernel_matrix", _numpy_kernel_matrix)
    monkeypatch.setattr(ops, "cholesky", lambda A, return_info=False, ctx=None:
                        (np.linalg.cholesky(A), 0, 2 * np.sum(np.log(np.diag(np.linalg.cholesky(A))))) if return_info else np.linalg.cholesky(A))
    monkeypatch.setattr(ops, "cho_solve", lambda L, B, forward_only=False, ctx=None: sl.cho_solve((L, True), B))

    def errors(L, mean, Y, want_errors=True, want_md2=False, ctx=None):
        E = sl.solve_triangular(L, Y - np.broadcast_to(mean, (L.shape[0],))[:, None], lower=True)
        return (E if want_errors else None), (np.sum(E * E, axis=0) if want_md2 else None)
    monkeypatch.setattr(ops, "cholesky_errors", errors)
    for i, ls in enumerate(g["corr_ls"]):
        for tag, X, Xp in (("1d", g["corr_X1"], None), ("3d", g["corr_X2"], None), ("3d_cross", g["corr_X2"], g["corr_Xp2"])):
            assert np.allclose(gb.rbf(X, Xp, ls=ls), g[f"rbf_{tag}_{i}"], rtol=1e-10, atol=1e-300)
            assert np.allclose(gb.gaussian(X, Xp, ls=ls), g[f"gauss_{tag}_{i}"], rtol=1e-10, atol=1e-300)
    assert gb.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], cov1=g["kl_cov1"]) == pytest.approx(float(g["kl_from_cov"]), rel=1e-11)
    assert gb.kl_gauss(g["kl_mu0"], g["kl_cov0"], g["kl_mu1"], chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_from_chol"]), rel=1e-11)
    assert gb.kl_gauss(0.2, 1.3, -0.4, cov1=0.9) == pytest.approx(float(g["kl_scalar"]), rel=1e-12)
    assert gb.kl_gauss(np.zeros(60), g["kl_cov0"], 0.25, chol1=g["kl_chol1"]) == pytest.approx(float(g["kl_scalar_mean1"]), rel=1e-11)


def test_decorator_helpers():
    """`lazy_property` / `default_attributes` (gsum/helpers.py:371-386, 416-501): the docstring example of the reference."""
    import gsum_b200 as gb

    class T:
        calls = 0

        def __init__(self, x, y):
            self.x, self._y = x, y

        @gb.lazy_property
        def total(self):
            """doc"""
            T.calls += 1
            return self.x + self._y

        @gb.default_attributes(x='x', y='_y')
        def add(self, x=None, y=None):
            return x + y

        @gb.default_attributes(args='x', kw='_y')
        def star(self, *args, **kw):
            return args, kw

    t = T(2, 3)
    assert t.total == 5 and t.total == 5 and T.calls == 1 and t._cache_total == 5 and T.total.__doc__ == "doc"
    assert t.add() == 5 and t.add(10) == 13 and t.add(y=1) == 3
    t.x = 20
    assert t.add() == 23 and t.total == 5                               # defaults are read at call time; the cache is not
    assert t.add(np.zeros(2), np.ones(2)).tolist() == [1.0, 1.0]        # arrays are never "missing"
    t.x, t._y = (1, 2), dict(a=1)
    assert t.star() == ((1, 2), dict(a=1)) and t.star(7) == ((7,), dict(a=1)) and t.star(b=2) == ((1, 2), dict(b=2))


def test_legacy_generators_host_logic(monkeypatch):
    """`generate_coefficients` / `toy_data` (gsum/helpers.py:36-68) with numpy stand-ins for the device calls."""
    import gsum_b200 as gb
    from gsum_b200 import ops
    monkeypatch.setattr(ops, "kernel_matrix", _numpy_kernel_matrix)
    monkeypatch.setattr(ops, "pivoted_cholesky", _numpy_pivoted_cholesky)
    monkeypatch.setattr(ops, "draws", _numpy_draws)
    X = np.linspace(0, 1, 12)[:, None]
    np.random.seed(4)
    c = gb.generate_coefficients(X, size=3000, beta=0.7, sd=1.5, noise=0.05, ls=0.3)
    assert c.shape == (3000, 12)
    K = 1.5 ** 2 * o.rbf_corr(X, ls=0.3) + 0.05 ** 2 * np.eye(12)
    sd = np.sqrt((K ** 2 + np.outer(np.diag(K), np.diag(K))) / 2999)
    assert np.max(np.abs(np.cov(c.T) - K) / sd) < 5.0 and np.max(np.abs(c.mean(axis=0) - 0.7) / np.sqrt(np.diag(K) / 3000)) < 5.0
    np.random.seed(4)
    c2 = gb.generate_coefficients(X, size=2, corr=lambda X, ls: o.rbf_corr(X, ls=ls), basis=lambda X: np.c_[np.ones(len(X)), X[:, 0]],
                                  beta=[0.7, 0.0], sd=1.5, noise=0.05, ls=0.3)
    assert c2.shape == (2, 12) and np.isfinite(c2).all()
    np.random.seed(4)
    y = gb.toy_data(X, orders=np.arange(12), ls=0.3)                    # curves are rows: square case only, as in the reference
    assert y.shape == (12, 12)


def test_grid_input_caches_do_not_outlive_a_fit(monkeypatch):
    """The per-fit caches of `log_marginal_likelihood_grid` (order differences, ref(X)) are keyed by object identity; `fit` must
    drop them, because a later array can reuse the address of a freed one.  Device calls replaced by recorders."""
    import gsum_b200 as gb
    from gsum_b200 import ops

    class _Handle:
        center = disp = 0.0
        df = scale = cov_factor = 1.0
        lml = 0.0

        def __init__(self, *a, **k):
            pass

        def close(self):
            pass
    seen = []
    monkeypatch.setattr(ops, "FitHandle", _Handle)
    monkeypatch.setattr(ops, "lml_grid", lambda X, dy, ref, orders, ls, Q, **kw: seen.append((dy.copy(), np.array(ref, copy=True))) or np.zeros((len(Q), len(ls))))
    X = np.linspace(0, 1, 7)[:, None]
    gp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=lambda X: 1.0 + X[:, 0], optimizer=None)
    y1 = np.cumsum(np.ones((7, 3)), axis=1)
    gp.fit(X, y1, orders=np.arange(3))
    gp.log_marginal_likelihood_grid([0.1, 0.2], ratio_vals=[0.4, 0.5])
    cached = gp._grid_inputs_cache
    gp.log_marginal_likelihood_grid([0.1, 0.2], ratio_vals=[0.4, 0.5])
    assert gp._grid_inputs_cache is cached                                 # reused between calls on the same fit
    y2 = np.cumsum(2.0 * np.ones((7, 3)), axis=1)
    gp.fit(X, y2, orders=np.arange(3))
    assert gp._grid_inputs_cache is None and gp._grid_ref_cache is None
    gp.log_marginal_likelihood_grid([0.1, 0.2], ratio_vals=[0.4, 0.5])
    assert np.array_equal(seen[0][0], np.ones((7, 3))) and np.array_equal(seen[-1][0], 2.0 * np.ones((7, 3)))
    assert np.array_equal(seen[-1][1], 1.0 + X[:, 0])
