"""bench.py — (l, Q)-grid log-likelihood evaluations per second at N = 1024, 6 orders (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path over THE grid BASELINE.json names (configs[3]): 128 length scales x 256 expansion
parameters at N = 1024, 6 orders.  With N GPUs that fixed grid is sharded (strong scaling): the 128 length scales are dealt
round-robin, 128/N per rank, one all-gather of the FP64 blocks and an on-device max-shift normalisation per step.  `value`
times the device-resident call with CUDA events; `e2e` times the public API (TruncationGP.log_marginal_likelihood_grid) with
host buffers in and out.  The weak-scaling figure (128 length scales PER GPU) is reported beside it under "weak".  One JSON
line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_ORDERS, N_LS, N_Q = 1024, 6, 128, 256
N_LS_PER_GPU = N_LS                                  # the weak-scaling side figure
METRIC = "(l,Q) grid log-likelihood evals/sec at N=1024, 6 orders"
WORKLOAD = "C4: N=1024, 6 orders, 128 l x 256 Q grid (configs[3] of BASELINE.json), length scales sharded over the GPUs"
# DRAM bytes of ONE factorisation launch at 128 length scales, recorded from an `ncu --set full` capture (not measured in
# this run): {profile file: dram__bytes_read.sum + dram__bytes_write.sum}
NCU_TRAFFIC = {"file": "profiles/r02_ncu_hetero_tma.txt", "bytes": 4.291450e9 + 660.052736e6, "n_ls": 128}


def make_inputs(n_ls):
    """X = linspace(0,1,1024); coefficients ~ GP(RBF(0.05) + 1e-6 I), Q = 0.5, ref = 1, seed 3."""
    from sklearn.gaussian_process.kernels import RBF
    rs = np.random.RandomState(3)
    X = np.linspace(0, 1, N_POINTS)[:, None]
    Lt = np.linalg.cholesky(RBF(0.05)(X) + 1e-6 * np.eye(N_POINTS))
    coeffs = Lt @ rs.standard_normal((N_POINTS, N_ORDERS))
    orders = np.arange(N_ORDERS)
    y = np.cumsum(coeffs * 0.5 ** orders, axis=-1)
    ls_vals = np.geomspace(0.005, 0.5, n_ls)
    q_vals = np.linspace(0.2, 0.8, N_Q)
    return X, y, orders, ls_vals, q_vals


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions (B200_PROFILING.md recipe).  The sampler runs
    from before the warm-up (nvidia-smi needs ~0.1 s to start) at a 20 ms period; every row is stamped when it is read
    and only rows that fall inside a timed window (`mark_start` / `mark_end`) are reported."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows, self._t0 = index, [], None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self._t0 = time.perf_counter()

    def mark_end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, time.perf_counter()))
            self._t0 = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for ts, r in self.rows if any(a <= ts <= b + 0.02 for a, b in self.windows)]
        rows = inside if inside else [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '', 1).isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '', 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": "timed regions (device-resident steps and end-to-end steps)" if inside else
                "whole run (no sample fell inside the timed regions)"}


def cpu_reference_cells(X, y, orders, ls_vals, q_vals, cells):
    """The reference's per-cell path (oracle port: same numpy/scipy/sklearn calls as gsum/models.py:1485 -> 912) on host cores."""
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    from oracle import gsum_oracle as o
    kern = RBF(0.05) + WhiteKernel(1e-6, 'fixed')
    pri = o.Priors(0, 0, 1, 1)
    n = X.shape[0]
    t0 = time.perf_counter()
    out = [o.truncation_lml(kern, [np.log(ls_vals[b])], X, y, orders, q_vals[a] * np.ones(n), np.ones(n), pri) for a, b in cells]
    return time.perf_counter() - t0, out


# ---- cell-parallel CPU arm: the cells are independent, so the strongest way to run the reference algorithm on a multi-core
# host is one worker process per core, each evaluating whole cells with a single BLAS thread (at N = 1024 a threaded
# LAPACK Cholesky scales poorly, a process per cell scales linearly).  Workers rebuild the deterministic inputs themselves.
_W = {}


_ONE_THREAD_ENV = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS")


def _cpu_worker_init(n_ls):
    # the parent exported *_NUM_THREADS=1 before the spawn; import everything that owns a thread pool FIRST, then limit
    # whatever is loaded and verify that every pool really is at one thread (one worker process per core)
    import scipy.linalg  # noqa: F401
    import sklearn.gaussian_process.kernels  # noqa: F401
    from oracle import gsum_oracle  # noqa: F401
    _W["inputs"] = make_inputs(n_ls)
    from threadpoolctl import threadpool_limits, threadpool_info
    _W["limit"] = threadpool_limits(limits=1)
    _W["pools"] = [int(p.get("num_threads", 1)) for p in threadpool_info()]
    assert all(t == 1 for t in _W["pools"]), f"CPU worker BLAS pools not limited to one thread: {_W['pools']}"


def _cpu_worker_cells(cells):
    X, y, orders, ls_vals, q_vals = _W["inputs"]
    return cpu_reference_cells(X, y, orders, ls_vals, q_vals, cells)[1]


def _cpu_worker_pools(_):
    return _W.get("pools", [])


class CellPool:
    """`workers` spawned processes (spawn, not fork: the parent may hold a CUDA context); `run(cells)` deals the cells
    round-robin and returns the wall time of the whole map."""

    def __init__(self, n_ls, workers):
        import multiprocessing as mp
        self.workers = workers
        saved = {k: os.environ.get(k) for k in _ONE_THREAD_ENV}
        os.environ.update({k: "1" for k in _ONE_THREAD_ENV})             # inherited by the spawned workers
        try:
            self.pool = mp.get_context("spawn").Pool(workers, initializer=_cpu_worker_init, initargs=(n_ls,))
            self.pool.map(_cpu_worker_cells, [[(0, 0)]] * workers)       # imports, inputs and one warm cell per worker
            pools = self.pool.map(_cpu_worker_pools, range(workers))
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        self.blas_threads_per_worker = max([t for p in pools for t in p] + [1])
        assert self.blas_threads_per_worker == 1

    def run(self, cells):
        chunks = [cells[i::self.workers] for i in range(self.workers) if cells[i::self.workers]]
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_worker_cells, chunks, chunksize=1)
        return time.perf_counter() - t0, out

    def close(self):
        self.pool.close()
        self.pool.join()


def stratified_cells(n_q, n_ls, k):
    qa = np.linspace(0, n_q - 1, k).round().astype(int)
    lb = np.linspace(0, n_ls - 1, k).round().astype(int)
    return [(int(a), int(b)) for a in qa for b in lb]


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


class best_blas_setting:
    """The CPU legs run the reference algorithm the way it is fastest on this host: the BLAS pool at all cores (numpy's
    default) or at one thread (torchrun exports OMP_NUM_THREADS=1, and at N = 1024 OpenBLAS' fork/join overhead can
    make the threaded Cholesky slower than the serial one) — whichever wins on a 4-cell probe.  Context manager."""

    def __init__(self, probe):
        self.probe, self.ctx, self.threads = probe, None, blas_threads()

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
        except Exception:
            return self
        best = None
        for n in sorted({os.cpu_count() or 1, 1}, reverse=True):
            with threadpool_limits(limits=n):
                self.probe()                       # warm
                t = self.probe()
            if best is None or t < best[0]:
                best = (t, n)
        self.ctx = threadpool_limits(limits=best[1])
        self.ctx.__enter__()
        self.threads = best[1]
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def base_config(world):
    return {"workload": WORKLOAD, "n_points": N_POINTS, "n_orders": N_ORDERS, "n_ls": N_LS, "n_q": N_Q,
            "n_ls_per_gpu": N_LS // world, "grid_cells_per_step": N_LS * N_Q}


def run_reference(args, rank, world):
    """The reference algorithm alone on the host cores (rank 0 only; the other ranks exit without work).  `value` is the
    cell-parallel arm (one process per core, one BLAS thread each — verified inside the workers); the serial loop of the
    notebook (one cell after the other, BLAS threads at their best setting) is reported beside it."""
    if rank != 0:
        return
    X, y, orders, ls_vals, q_vals = make_inputs(N_LS)
    workers = os.cpu_count() or 1
    per_step = 2 * workers                                        # cells per step: two per core (~0.1-0.2 s of wall time)
    k = int(np.ceil(np.sqrt(per_step)))
    cells = stratified_cells(N_Q, N_LS, k)[:per_step]
    pool = CellPool(N_LS, workers)
    try:
        for _ in range(args.warmup):
            pool.run(cells[:workers])
        total = 0.0
        for _ in range(args.steps):
            dt, _ = pool.run(cells)
            total += dt
    finally:
        pool.close()
    value = len(cells) * args.steps / total
    serial_cells = stratified_cells(N_Q, N_LS, 3)[:8]
    with best_blas_setting(lambda: cpu_reference_cells(X, y, orders, ls_vals, q_vals, serial_cells[:4])[0]) as blas:
        dt_serial, _ = cpu_reference_cells(X, y, orders, ls_vals, q_vals, serial_cells)
    sample = (f"{len(cells)} cells/step of the 256x128 grid (stratified), {args.steps} steps, {workers} worker processes x "
              f"{pool.blas_threads_per_worker} BLAS thread (threadpool_info checked in every worker); one Cholesky + 4 cho_solve per "
              f"cell as in the reference")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_config(max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": workers, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(),
                         "serial_loop": {"value": len(serial_cells) / dt_serial, "unit": "evals/s", "blas_threads": blas.threads,
                                         "note": "the notebook's own cell-after-cell loop (docs/notebooks/correlated_EFT_publication.ipynb cell 53)"}},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def measure_fp64_peak(torch, n=4096, reps=6):
    """cuBLAS DGEMM throughput (TFLOP/s), best of `reps` — the FP64 roofline denominator (MEASURED_PEAKS.json has none)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best * 1e-9


_REAL_STDOUT = None


def _emit(line):
    """Write the result line to the process' original stdout (see run_ours)."""
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


class DeviceGrid:
    """Device-resident inputs and outputs of this rank's share of an (n_ls_total x N_Q) grid: one `step()` = K1..K4 on the
    rank's length scales, the all-gather of the blocks and the on-device normalisation."""

    def __init__(self, torch, dist, ctx, dev, rank, world, n_ls_total):
        from gsum_b200 import ops
        from gsum_b200.helpers import _order_differences
        self.torch, self.dist, self.ops, self.ctx, self.world = torch, dist, ops, ctx, world
        self.inputs = make_inputs(n_ls_total)
        X, y, orders, ls_vals, q_vals = self.inputs
        self.mine = np.arange(rank, n_ls_total, world)
        self.dy = np.ascontiguousarray(_order_differences(y))
        detf = N_POINTS * float(orders.sum()) * np.log(np.abs(q_vals))
        t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
        self.dX, self.ddy, self.dref, self.dord = t(X), t(self.dy), t(np.ones(N_POINTS)), t(orders.astype(np.int32), torch.int32)
        self.dls, self.dQ, self.ddetf = t(ls_vals[self.mine][:, None]), t(q_vals), t(detf)
        self.ll_block = torch.empty((N_Q, len(self.mine)), dtype=torch.float64, device=dev)
        self.gathered = torch.empty((world * N_Q, len(self.mine)), dtype=torch.float64, device=dev)
        self.post = torch.empty_like(self.gathered)
        self.lse = torch.empty(1, dtype=torch.float64, device=dev)
        self.kw = dict(constant=1.0, noise=1e-6, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0, scale0=1.0)

    def step(self):
        self.ops.lml_grid_device(self.ctx, self.dX, self.ddy, self.dref, self.dord, self.dls, self.dQ, self.ddetf, self.ll_block, **self.kw)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered, self.ll_block)
            self.ops.grid_normalize_device(self.ctx, self.gathered, self.post, self.lse)
        else:
            self.ops.grid_normalize_device(self.ctx, self.ll_block, self.post, self.lse)


def other_configs():
    """The other single-GPU configurations of BASELINE.json (C2, C3, C5) at full size through the public API —
    numpy in, numpy out, host<->device copies included, wall clock with the call's own synchronisation; best of three after one
    warm-up call — so that the driver's run records them with the same clock evidence as the headline.
    Outside the headline's timed regions; a failure is reported in place and never costs the main line.  tools/perf_configs.py
    is the long form with the host's timings beside them."""
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    import gsum_b200 as gb
    out = {}

    def best_ms(f, reps=3):
        f()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t0)
        return 1e3 * min(ts)

    def section(name, body):
        try:
            out[name] = body()
        except Exception as e:                                     # noqa: BLE001 — report, keep the headline
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    def c2():
        rs = np.random.RandomState(1)
        n, orders = 200, np.arange(6)
        X = np.linspace(0, 1, n)[:, None]
        coeffs = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, 6)
        y = np.cumsum(coeffs * 0.5 ** orders, axis=1)
        gp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
        gp.fit(X, y, orders=orders)
        ls_vals, q_vals = np.linspace(0.02, 0.5, 64), np.linspace(0.3, 0.7, 64)
        ms = best_ms(lambda: gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals))
        return {"workload": "C2: N=200, 6 orders, 64 l x 64 Q grid (small-N one-CTA path)", "grid_ms": ms, "evals_per_s": 4096 / (ms * 1e-3)}

    def c3():
        rs = np.random.RandomState(2)
        g1, orders = np.linspace(0, 1, 50), np.arange(6)
        X = np.stack(np.meshgrid(g1, g1, indexing="ij"), -1).reshape(-1, 2)
        Xt = rs.rand(10000, 2)
        coeffs = np.linalg.cholesky(RBF([0.02, 0.03])(X) + 1e-8 * np.eye(len(X))) @ rs.randn(len(X), 6)
        y = np.cumsum(coeffs * 0.4 ** orders, axis=1)
        gp = gb.TruncationGP(RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-6, 'fixed'), ratio=0.4, ref=1, center=0, disp=0, df=1, scale=1,
                             optimizer=None)
        fit_ms = best_ms(lambda: gp.fit(X, y, orders=orders))
        std_ms = best_ms(lambda: gp.coeffs_process.predict(Xt, return_std=True))
        both_ms = best_ms(lambda: gp.predict(Xt, order=5, return_std=True, kind='both'))
        return {"workload": "C3: 2-D 50 x 50 = 2500 training points, 10 000 test points", "fit_ms": fit_ms,
                "coeffs_process_predict_std_ms": std_ms, "truncation_predict_both_std_ms": both_ms,
                "forward_solve_tflops": 2500.0 ** 2 * 10000 / (std_ms * 1e-3) * 1e-12}

    def c5():
        n = 4096
        Xd = np.linspace(0, 1, n)[:, None]
        cov, mean = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n)), np.zeros(n)
        holder = {}

        def construct():
            holder["d"] = gb.Diagnostic(mean, cov, random_state=1)
        init_ms = best_ms(construct, reps=2)
        d = holder["d"]
        Y = d.samples(64)
        md_ms = best_ms(lambda: d.md_squared(Y))
        pc_ms = best_ms(lambda: d.pivoted_cholesky_errors(Y))
        iv = np.linspace(0, 1, 101)
        draw_ms = best_ms(lambda: d.sample_coverage(100000, iv, counts=True, per_draw=False), reps=2)
        return {"workload": "C5: N=4096 covariance; Cholesky + pivoted Cholesky, 64 held-out curves, 1e5 draws x 101 intervals",
                "diagnostic_constructor_ms": init_ms, "md_squared_64_ms": md_ms, "pivoted_cholesky_errors_64_ms": pc_ms,
                "draws_1e5_coverage_counts_ms": draw_ms, "draws_tflops_triangular": float(n) * n * 1e5 / (draw_ms * 1e-3) * 1e-12}

    section("C2", c2)
    section("C3", c3)
    section("C5", c5)
    out["timing"] = "wall clock around the public API call (numpy in / out, copies included), best of 3 (2 for the C5 constructor and draws) after one warm-up call"
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    import gsum_b200 as gb
    from gsum_b200 import _lib

    if N_LS % world:
        raise SystemExit(f"--gpus must divide {N_LS}")
    # NCCL's communicator lines stay visible (INFO unless the caller chose otherwise).  NCCL logs on the C-level stdout, so
    # the process' fd 1 is pointed at stderr for the whole run and the ONE JSON line goes to a saved duplicate of the real
    # stdout (_emit): the log is on stderr, the result line alone on stdout.
    # (forced: an image-level NCCL_DEBUG=WARN/VERSION would hide them; GSUM_NCCL_DEBUG overrides)
    os.environ["NCCL_DEBUG"] = os.environ.get("GSUM_NCCL_DEBUG", "INFO")
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)                     # one non-default stream for torch, NCCL and the library
    torch.cuda.set_stream(stream)
    ctx = _lib.Context(local_rank, stream.cuda_stream)         # library work is enqueued on torch's current stream
    assert stream.cuda_stream != 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(vals):
        tt = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return list(vals), list(vals)
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return [float(v) for v in mx], [float(v) for v in sm]

    def time_device(grid, steps, warmup, sampler=None):
        """`steps` device-resident steps, each bracketed by CUDA events on the launch stream, L2 flushed before each (untimed);
        returns (ms summed over the steps: max over ranks, launches: sum over ranks, this rank's factorisation bracket)."""
        for _ in range(warmup):
            grid.step()
        barrier()
        ctx.profile(True)
        launches0 = ctx.launch_count
        evs = []
        barrier()
        if sampler:
            sampler.mark_start()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); grid.step(); e1.record(stream)
            evs.append((e0, e1))
        barrier()
        if sampler:
            sampler.mark_end()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = ctx.launch_count - launches0
        fact_ms, fact_flops, _ = ctx.profile_read()
        ctx.profile(False)
        mx, sm = reduce_max_sum([ms, float(launches)])
        return mx[0], int(sm[1]), fact_ms, fact_flops

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- the metric: the fixed 128 x 256 grid, length scales sharded over the ranks (strong scaling) ----
    grid = DeviceGrid(torch, dist, ctx, dev, rank, world, N_LS)
    X, y, orders, ls_vals, q_vals = grid.inputs
    ms_total, launches, fact_ms, fact_flops = time_device(grid, args.steps, max(args.warmup, 3), sampler)
    cells_per_step = N_Q * N_LS
    value = cells_per_step * args.steps / (ms_total * 1e-3)

    # ---- side figure: weak scaling, 128 length scales PER GPU (the round-1 bench line) ----
    weak = None
    if world > 1:
        wgrid = DeviceGrid(torch, dist, ctx, dev, rank, world, N_LS_PER_GPU * world)
        wsteps = max(3, min(args.steps, 10))
        wms, _, wf_ms, wf_flops = time_device(wgrid, wsteps, 3)
        weak = {"value": N_Q * N_LS_PER_GPU * world * wsteps / (wms * 1e-3), "unit": "evals/s", "ms_per_step": wms / wsteps, "steps": wsteps,
                "n_ls_per_gpu": N_LS_PER_GPU, "grid": f"{N_LS_PER_GPU * world} l x {N_Q} Q",
                "factorisation_tflops_rank0": wf_flops / (wf_ms * 1e-3) * 1e-12 if wf_ms > 0 else None}
        del wgrid

    # ---- end-to-end arm: the public API with host buffers (H2D of inputs and D2H of the grid inside the timed region) ----
    gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
    gp.fit(X, y, orders=orders)
    group = dist.group.WORLD if world > 1 else None
    for _ in range(3):
        ll_host = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=group)
    barrier()
    sampler.mark_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_host = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=group)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    others = None
    if world == 1 and not args.no_other_configs:
        # side figures with a clock record of their own: the headline's windows and clock summary stay exactly what they were
        side = type(sampler)(local_rank)
        side.start()
        time.sleep(0.15)
        side.mark_start()
        others = other_configs()
        side.mark_end()
        others["clocks"] = side.stop()
        others["clocks"]["window"] = "the other-configs section"
    e2e_s = reduce_max_sum([e2e_s])[0][0]
    e2e_value = cells_per_step * args.steps / e2e_s
    n_mine = len(grid.mine)
    # per rank and step: the packed inputs go up once; every rank reads the WHOLE gathered grid back (world > 1), or its
    # grid + log-determinants + status (1 GPU, C ABI with host buffers)
    h2d_rank = 8 * (X.size + grid.dy.size + N_POINTS + n_mine + N_Q + N_Q) + 4 * N_ORDERS + (4 if world > 1 and N_ORDERS % 2 else 0)
    d2h_rank = 8 * N_Q * N_LS if world > 1 else 8 * (N_Q * N_LS + N_LS) + 4 * N_LS
    assert np.isfinite(ll_host).all() and ll_host.shape == (N_Q, N_LS)

    # ---- untimed: the sharded grid against the same grid computed by ONE GPU alone (whole grid, bit for bit) ----
    sharded_equal = None
    if world > 1 and rank == 0:
        single = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)
        sharded_equal = bool(np.array_equal(single, ll_host))

    if rank == 0:
        # parity spot check of the timed configuration (not timed)
        dt_cpu, want = cpu_reference_cells(X, y, orders, ls_vals, q_vals, [(0, 0), (100, 40), (255, 64)])
        got = [ll_host[0, 0], ll_host[100, 40], ll_host[255, 64]]
        parity = max(abs(g - w) / abs(w) for g, w in zip(got, want))
        # roofline of the dominant kernel (the bordered Cholesky launch, chol_hetero_tma_kernel), this rank's launches
        peak = measure_fp64_peak(torch)
        achieved = fact_flops / (fact_ms * 1e-3) * 1e-12 if fact_ms > 0 else 0.0
        # CPU baseline on a bounded sample (~10-20 s of CPU work in total): cell-parallel over all host cores, and the
        # notebook's serial loop beside it
        workers = os.cpu_count() or 1
        cells = stratified_cells(N_Q, N_LS, 16)
        pool = CellPool(N_LS, workers)
        try:
            cpu_s, _ = pool.run(cells)
        finally:
            pool.close()
        serial_cells = stratified_cells(N_Q, N_LS, 5)
        with best_blas_setting(lambda: cpu_reference_cells(X, y, orders, ls_vals, q_vals, serial_cells[:4])[0]) as blas:
            serial_s, _ = cpu_reference_cells(X, y, orders, ls_vals, q_vals, serial_cells)
        same_launch = (n_mine == NCU_TRAFFIC["n_ls"])
        cfg = base_config(world)
        cfg.update({"timing": "CUDA events on the launch stream per step, max over ranks; L2 flushed (256 MB memset) between steps",
                    "collective": "one NCCL all-gather of the FP64 blocks + device logsumexp" if world > 1 else "none (1 GPU); device logsumexp"})
        out = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(h2d_rank) * world, "d2h_bytes_per_step": int(d2h_rank) * world,
                    "h2d_bytes_per_step_per_rank": int(h2d_rank), "d2h_bytes_per_step_per_rank": int(d2h_rank),
                    "api": "TruncationGP.log_marginal_likelihood_grid (numpy in, numpy out on every rank; staged through pinned host buffers)",
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": NCU_TRAFFIC["bytes"] if same_launch else None,
                         "traffic_source": (f"RECORDED constant, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                            f"factorisation launch at {NCU_TRAFFIC['n_ls']} length scales, ncu --set full ({NCU_TRAFFIC['file']})")
                         if same_launch else f"no ncu capture of the {n_mine}-length-scale launch",
                         "kernel": "chol_hetero_tma_kernel (FP64 DMMA bordered Cholesky + forward solves, K2+K3; one cooperative launch per step "
                                   f"over this rank's {n_mine} length scales)",
                         "flops_per_step": fact_flops / max(args.steps, 1), "kernel_ms_per_step": fact_ms / max(args.steps, 1),
                         "peak_source": "cuBLAS DGEMM 4096^3 measured live in this run (MEASURED_PEAKS.json has no FP64 figure; "
                                        "profiles/r01_dgemm_peak.json: 35.5 TFLOP/s at 8192^3)"},
            "cpu_baseline": {"value": len(cells) / cpu_s, "unit": "evals/s", "cores": workers, "kind": "port", "host_cpus": os.cpu_count(),
                             "sample": f"{len(cells)} stratified cells of the {N_Q}x{N_LS} grid, per-cell reference algorithm "
                                       f"(numpy/scipy/sklearn), {workers} worker processes x {pool.blas_threads_per_worker} BLAS thread "
                                       f"(threadpool_info checked in every worker), {cpu_s:.1f} s wall",
                             "serial_loop": {"value": len(serial_cells) / serial_s, "unit": "evals/s", "blas_threads": blas.threads,
                                             "sample": f"{len(serial_cells)} cells, {serial_s:.1f} s"}},
            "clocks": clocks, "parity_spot_check_rel": parity,
        }
        if weak is not None:
            out["weak"] = weak
        if others is not None:
            out["other_configs"] = others
        if sharded_equal is not None:
            out["sharded_equals_single_gpu"] = sharded_equal
        _emit(json.dumps(out))
    if world > 1:
        barrier()
        from gsum_b200 import distributed as gdist
        gdist.release_graphs()          # the captured graphs hold NCCL kernels: they go before the communicator does
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C2 / C3 / C5 side figures of the 1-GPU line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
