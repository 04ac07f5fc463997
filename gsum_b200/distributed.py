"""Sharding of the path over the GPUs of one node: the (Q, l) likelihood grid by length scale, the
posterior draws of the credible-interval diagnostic by draw, `predict` by test point.

Cells are independent given (X, y) and all the cost is per length scale (one factorisation each), so the
length scales are dealt round-robin to the ranks of a ``torch.distributed`` process group, every rank
evaluates its (n_q, n_ls/P) block with the single-GPU path, and ONE all-gather of the FP64 blocks rebuilds
the grid on every rank; the max-shift normalisation then runs on the device.  There is no other exchange.
Each cell is computed by exactly one rank with the same kernels, so the gathered grid is bit-identical to
the single-GPU grid.


Draws (config C5): every rank holds a replica of the factor, takes a contiguous slice of the draw axis (the device
generator is counter based, so slice j reproduces exactly the draws the unsharded call would produce there) and the
ranks exchange ONE all-reduce of the int64 coverage counts — integers, so the sum does not depend on the order.
Predict (config C3): the fit is replicated, test points are dealt in contiguous blocks and one all-gather rebuilds the
(M, n_curves) mean and the (M,) standard deviation; a full (M, M) covariance is not sharded.

torch is used for the plumbing only (process group, device buffers for NCCL).
"""
from __future__ import annotations

import numpy as np

import os

from . import ops

__all__ = ["shard_indices", "assemble_blocks", "lml_grid_sharded", "shard_range", "sample_coverage_sharded", "predict_sharded",
           "release_graphs"]


def shard_indices(n_ls, world_size, rank):
    """Length-scale indices owned by `rank`: i with i % world_size == rank."""
    return np.arange(rank, n_ls, world_size)


def assemble_blocks(blocks, n_ls, world_size):
    """Inverse of the round-robin deal: blocks[r] is (n_q, padded) holding columns shard_indices(n_ls, P, r)."""
    n_q = blocks[0].shape[0]
    out = np.empty((n_q, n_ls), dtype=blocks[0].dtype)
    for r in range(world_size):
        idx = shard_indices(n_ls, world_size, r)
        out[:, idx] = blocks[r][:, :len(idx)]
    return out


def lml_grid_sharded(X, dy, ref, orders, ls, Q, group=None, normalize=False, **kw):
    """Evaluate ops.lml_grid on this rank's length scales and all-gather the blocks.

    Works with the NCCL backend (device buffers, one all-gather over NVLink) and with gloo (host buffers; used by the
    CPU-side tests of the sharding logic with a stub evaluator).  Returns the full (n_q, n_ls) grid on every rank
    (and the normalised posterior + logsumexp if `normalize`)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ls = np.asarray(ls, dtype=np.float64).reshape(len(ls), -1)
    n_ls, n_q = ls.shape[0], np.asarray(Q).shape[0]
    mine = shard_indices(n_ls, world, rank)
    per = -(-n_ls // world)                                   # ceil: every rank sends the same count
    backend = dist.get_backend(group)
    if backend == "nccl" and "_evaluator" not in kw:
        return _sharded_device(X, dy, ref, orders, ls, Q, mine, per, world, group, normalize, kw)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    block = np.full((n_q, per), -np.inf)
    if len(mine):
        evaluator = kw.pop("_evaluator", ops.lml_grid)
        block[:, :len(mine)] = evaluator(X, dy, ref, orders, ls[mine], Q, **kw)
    send = torch.from_numpy(block).to(dev)
    recv = torch.empty((world * n_q, per), dtype=torch.float64, device=dev)   # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)     # the single collective of the path
    full = assemble_blocks(list(recv.view(world, n_q, per).cpu().numpy()), n_ls, world)
    if normalize:
        post, lse = ops.grid_normalize(full)
        return full, post, lse
    return full


_stream_ctx = {}
_side_streams = {}        # device index -> the stream the sharded grid runs on when torch's current stream is the default one
_grid_plans = {}
_USE_GRAPH = os.environ.get("GSUM_B200_GRID_GRAPH", "1") != "0"       # replay the sharded grid's device sequence as one CUDA graph (GSUM_B200_GRID_GRAPH=0 disables)
_PROF = None            # dev probe (tools/e2e_sharded_breakdown.py): dict of accumulated host seconds per section


def release_graphs():
    """Drop the cached staging plans and their CUDA graphs.  Call it BEFORE `torch.distributed.destroy_process_group()`:
    a captured graph holds the NCCL communicator's kernels, and destroying the communicator first blocks (observed with
    NCCL 2.28 / torch 2.11 at world size 2)."""
    if _grid_plans:
        import torch
        torch.cuda.synchronize()
        for plan in list(_grid_plans.values()):
            plan.graphs.clear()
        _grid_plans.clear()
        torch.cuda.synchronize()


def _tick(name, t0):
    import time
    t1 = time.perf_counter()
    if _PROF is not None:
        _PROF[name] = _PROF.get(name, 0.0) + (t1 - t0)
    return t1


class _GridPlan:
    """Persistent staging of one sharded-grid shape: a pinned host buffer and its device twin for the packed inputs, the
    send / receive buffers of the all-gather, the grid in its final layout and a pinned host buffer for the result.
    Nothing is allocated and no pageable buffer is touched by the device on the per-call path."""

    def __init__(self, dev, sizes, n_q, per, world, n_ls, n_mine):
        import torch
        self.offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        words = int(self.offs[-1])
        self.h_in = torch.empty(words, dtype=torch.float64, pin_memory=True)
        self.h_in_np = self.h_in.numpy()
        self.d_in = torch.empty(words, dtype=torch.float64, device=dev)
        self.v = [self.d_in[int(self.offs[i]):int(self.offs[i + 1])] for i in range(len(sizes))]
        self.send = torch.full((n_q, per), float("-inf"), dtype=torch.float64, device=dev)
        self.out = self.send if n_mine == per else torch.empty((n_q, max(n_mine, 1)), dtype=torch.float64, device=dev)
        self.recv = torch.empty((world * n_q, per), dtype=torch.float64, device=dev)   # rank-major concatenation along dim 0
        self.full = torch.empty((n_q, per * world), dtype=torch.float64, device=dev)
        self.post = torch.empty((n_q, n_ls), dtype=torch.float64, device=dev)
        self.lse = torch.empty(1, dtype=torch.float64, device=dev)
        self.h_out = torch.empty((2, n_q, per * world), dtype=torch.float64, pin_memory=True)
        self.h_lse = torch.empty(1, dtype=torch.float64, pin_memory=True)
        self.graphs = {}        # scalar arguments -> 1 (seen once, eager) | torch.cuda.CUDAGraph | -1 (capture failed)


def _sharded_device(X, dy, ref, orders, ls, Q, mine, per, world, group, normalize, kw):
    """NCCL path: this rank's block never leaves the device between the likelihood kernels and the all-gather — the inputs
    go up in ONE copy from a persistent pinned buffer, the kernels and the collective are enqueued on torch's current
    stream, the round-robin deal is undone by one strided copy on the device into the final layout, and ONE device-to-host
    copy into pinned memory returns the grid (`_GridPlan`: no allocation on the per-call path).

    From the third call with the same shapes and scalar arguments on, the whole device sequence (upload, kernels, all-gather,
    permute, download) is replayed as ONE CUDA graph captured from the second call: with the grid sharded eight ways the
    kernels take less time than the host needs to enqueue them one by one."""
    import time
    import torch
    import torch.distributed as dist
    from . import _lib

    t0 = time.perf_counter()
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream(dev)
    if stream.cuda_stream == 0:
        # The legacy default stream cannot be handed to the library (gsum_ctx_create takes 0 as "create your own", and that
        # stream does not synchronise with torch's): the kernels would race the all-gather that torch enqueues on ITS current
        # stream (seen at world size 2: the first call of a shape gathered a block whose kernels had not finished).  Everything
        # of the call — staging buffers, kernels, collective, copies — runs on one side stream per device instead.
        side = _side_streams.get(dev.index)
        if side is None:
            side = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            return _sharded_device(X, dy, ref, orders, ls, Q, mine, per, world, group, normalize, kw)
    key = (dev.index, stream.cuda_stream)
    ctx = _stream_ctx.get(key)
    if ctx is None:
        ctx = _stream_ctx[key] = _lib.Context(dev.index, stream.cuda_stream)
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    n = X.shape[0]
    Q = np.asarray(Q, dtype=np.float64)
    n_q, n_ls = Q.shape[0], ls.shape[0]
    detf = kw.get("detf")
    orders32 = np.asarray(orders, dtype=np.int32)
    dy = np.asarray(dy, dtype=np.float64)
    n_mine = len(mine)
    ls_dim = ls.shape[1]
    sizes = (X.size, dy.size, n, max(n_mine, 1) * ls_dim, Q.size, n_q, (orders32.size + 1) // 2)
    pkey = key + sizes + (per, world, n_ls, id(group))
    plan = _grid_plans.get(pkey)
    if plan is None:
        if len(_grid_plans) > 8:
            _grid_plans.clear()
        plan = _grid_plans[pkey] = _GridPlan(dev, sizes, n_q, per, world, n_ls, n_mine)
    scalars = (bool(kw.get("q_x_dependent", False)), float(kw.get("constant", 1.0)), float(kw.get("noise", 0.0)),
               float(kw.get("nugget", 1e-10)), float(kw.get("center0", 0.0)), float(kw.get("disp0", 0.0)), float(kw.get("df0", 1.0)),
               float(kw.get("scale0", 1.0)), bool(kw.get("student", False)), detf is None, bool(normalize), Q.shape, orders32.size)
    if n_mine:
        h, o = plan.h_in_np, plan.offs
        h[o[0]:o[1]] = X.ravel()
        h[o[1]:o[2]] = dy.ravel()
        h[o[2]:o[3]] = np.asarray(ref, dtype=np.float64)          # broadcasts a scalar
        h[o[3]:o[3] + n_mine * ls_dim] = ls[mine].ravel()
        h[o[4]:o[5]] = Q.ravel()
        h[o[5]:o[6]] = 0.0 if detf is None else np.asarray(detf, dtype=np.float64)
        h[o[6]:o[7]].view(np.int32)[:orders32.size] = orders32
    t0 = _tick("pack", t0)

    def enqueue():
        if n_mine:
            plan.d_in.copy_(plan.h_in, non_blocking=True)
            v = plan.v
            ops.lml_grid_device(ctx, v[0].view(n, -1), v[1].view(n, -1), v[2], v[6].view(torch.int32)[:scalars[12]],
                                v[3][:n_mine * ls_dim].view(n_mine, -1), v[4].view(scalars[11]), None if scalars[9] else v[5],
                                plan.out, q_x_dependent=scalars[0], constant=scalars[1], noise=scalars[2], nugget=scalars[3],
                                center0=scalars[4], disp0=scalars[5], df0=scalars[6], scale0=scalars[7], student=scalars[8])
            if plan.out is not plan.send:
                plan.send[:, :n_mine] = plan.out[:, :n_mine]
        dist.all_gather_into_tensor(plan.recv, plan.send, group=group)            # the single collective of the path
        # rank r's local column j is length scale r + j * world: (world, n_q, per) -> (n_q, per, world), one strided copy
        plan.full.view(n_q, per, world).copy_(plan.recv.view(world, n_q, per).permute(1, 2, 0))
        plan.h_out[0].copy_(plan.full, non_blocking=True)
        if normalize:
            full_d = plan.full if per * world == n_ls else plan.full[:, :n_ls].contiguous()
            ops.grid_normalize_device(ctx, full_d, plan.post, plan.lse)
            plan.h_out[1, :, :n_ls].copy_(plan.post, non_blocking=True)
            plan.h_lse.copy_(plan.lse, non_blocking=True)

    g = plan.graphs.get(scalars)
    if isinstance(g, torch.cuda.CUDAGraph):
        g.replay()
    else:
        # first call: eager (workspaces are allocated, task lists uploaded); second: capture; failures keep the eager path
        if g == 1 and _USE_GRAPH and stream.cuda_stream != 0:
            try:
                stream.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    enqueue()
                plan.graphs[scalars] = graph
                graph.replay()
            except Exception as e:                          # noqa: BLE001 — any capture problem: stay on direct launches
                import warnings
                warnings.warn(f"gsum_b200: CUDA-graph capture of the sharded grid failed ({e!r}); using direct launches")
                plan.graphs[scalars] = -1
                torch.cuda.synchronize()
                enqueue()
        else:
            if g is None:
                if len(plan.graphs) > 4:
                    plan.graphs.clear()
                plan.graphs[scalars] = 1
            enqueue()
    t0 = _tick("enqueue", t0)
    stream.synchronize()
    t0 = _tick("synchronize", t0)
    full = plan.h_out[0].numpy()[:, :n_ls].copy()
    t0 = _tick("copy out", t0)
    if normalize:
        return full, plan.h_out[1].numpy()[:, :n_ls].copy(), float(plan.h_lse[0])
    return full


def shard_range(n, world_size, rank):
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one; the first ranks get the extras)."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _comm_device(group):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def sample_coverage_sharded(diagnostic, n_draws, intervals, seed=None, group=None, _evaluator=None):
    """Credible-interval coverage of `n_draws` fresh posterior draws with the draw axis sharded over the ranks.

    Each rank draws its slice on its GPU (`Diagnostic.sample_coverage(first_draw=lo, n_total=n_draws, counts=True)`) and
    the only exchange is one all-reduce (sum) of the int64 counts of (draw, point) pairs inside each interval.  Returns the
    (n_intervals,) mean coverage over all draws on every rank — identical to the 1-GPU value for the same seed."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_draws, world, rank)
    n_alpha = np.atleast_1d(intervals).shape[0]
    counts = np.zeros(n_alpha, dtype=np.int64)
    if hi > lo:
        evaluator = _evaluator or (lambda lo_, n_: diagnostic.sample_coverage(n_, intervals, seed=seed, first_draw=lo_,
                                                                               n_total=n_draws, counts=True, per_draw=False))
        counts = np.asarray(evaluator(lo, hi - lo), dtype=np.int64)
    t = torch.from_numpy(counts).to(_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)    # the single collective of the path
    n_points = diagnostic.mean.shape[0]
    return t.cpu().numpy().astype(np.float64) / (float(n_draws) * float(n_points))


def predict_sharded(process, X, return_std=False, group=None, **predict_kw):
    """`process.predict(X, return_std=...)` with the test points sharded over the ranks (contiguous blocks) and one
    all-gather of the [mean | std] blocks.  `process` is a fitted ConjugateGaussianProcess / ConjugateStudentProcess /
    TruncationGP / TruncationTP replica on every rank; extra keyword arguments (`order`, `kind`, `Xc`, `y`, ...) are passed
    through.  A full covariance (`return_cov`) couples all test points and is not sharded."""
    import torch
    import torch.distributed as dist

    if predict_kw.get("return_cov"):
        raise NotImplementedError("gsum_b200: predict_sharded shards test points; a full (M, M) covariance is not sharded")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    m = X.shape[0]
    per = -(-m // world)
    lo, hi = min(rank * per, m), min((rank + 1) * per, m)
    block = None
    if hi > lo:
        res = process.predict(X[lo:hi], return_std=return_std, **predict_kw)
        mean, std = res if return_std else (res, None)
        mean2 = mean[:, None] if mean.ndim == 1 else mean
        block = np.concatenate([mean2, std[:, None]], axis=1) if return_std else mean2
        shape1d, width = mean.ndim == 1, block.shape[1]
    else:
        shape1d, width = False, 0
    # every rank must agree on the block width even if it owns no points: rank 0 always owns some
    meta = torch.tensor([width, int(shape1d)], dtype=torch.int64, device=_comm_device(group))
    dist.broadcast(meta, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    width, shape1d = int(meta[0]), bool(int(meta[1]))
    send = np.zeros((per, width))
    if block is not None:
        send[:hi - lo] = block
    dev = _comm_device(group)
    recv = torch.empty((world * per, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, torch.from_numpy(send).to(dev), group=group)   # the single data collective
    full = recv.cpu().numpy()[:m]
    n_mean = width - 1 if return_std else width
    mean = full[:, 0] if shape1d else np.ascontiguousarray(full[:, :n_mean])
    return (mean, np.ascontiguousarray(full[:, n_mean])) if return_std else mean


# ---- the same sharding through the library's OWN communicator (C ABI: gsum_comm_init / gsum_grid_allgather) -------------------
# For hosts without torch.distributed: libgsum_b200.so resolves NCCL itself (dlopen) and owns the communicator; only the 128-byte
# id has to travel between the processes, by whatever channel the host has.  Here that channel is a torch process group (any
# backend) because the test harness has one; an MPI_Bcast or a socket does the same job.
def cabi_comm_init(ctx=None, group=None):
    """Create the library's NCCL communicator on `ctx` for the ranks of `group`; returns (nranks, rank)."""
    import ctypes as C
    import torch.distributed as dist
    from ._lib import default_context
    ctx = ctx or default_context()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ident = C.create_string_buffer(128)
    if rank == 0:
        ctx.check(ctx.lib.gsum_comm_unique_id(ctx.handle, ident), "gsum_comm_unique_id")
    box = [ident.raw]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident = C.create_string_buffer(box[0], 128)
    ctx.check(ctx.lib.gsum_comm_init(ctx.handle, world, rank, ident), "gsum_comm_init")
    return world, rank


def cabi_comm_destroy(ctx=None):
    from ._lib import default_context
    ctx = ctx or default_context()
    ctx.check(ctx.lib.gsum_comm_destroy(ctx.handle), "gsum_comm_destroy")


def lml_grid_sharded_cabi(X, dy, ref, orders, ls, Q, world, rank, normalize=False, ctx=None, **kw):
    """The sharded grid with host buffers and nothing but C-ABI calls: gsum_lml_grid on this rank's length scales, then
    gsum_grid_allgather (one ncclAllGather + the un-deal + optionally the normalisation, on the device)."""
    from ._lib import MEM_HOST, default_context
    ctx = ctx or default_context()
    ls = np.asarray(ls, dtype=np.float64).reshape(len(ls), -1)
    Q = np.asarray(Q, dtype=np.float64)
    n_ls, n_q = ls.shape[0], Q.shape[0]
    mine = shard_indices(n_ls, world, rank)
    per = -(-n_ls // world)
    block = np.full((n_q, per), -np.inf)
    if len(mine):
        block[:, :len(mine)] = ops.lml_grid(X, dy, ref, orders, ls[mine], Q, ctx=ctx, **kw)
    full = np.empty((n_q, n_ls))
    post = np.empty((n_q, n_ls)) if normalize else None
    lse = np.empty(1) if normalize else None
    p = lambda a: None if a is None else a.ctypes.data
    ctx.check(ctx.lib.gsum_grid_allgather(ctx.handle, p(block), n_q, per, n_ls, p(full), p(post), p(lse), MEM_HOST), "gsum_grid_allgather")
    return (full, post, float(lse[0])) if normalize else full
