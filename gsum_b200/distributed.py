"""Sharding of the (Q, l) likelihood grid over the GPUs of one node.

Cells are independent given (X, y) and all the cost is per length scale (one factorisation each), so the
length scales are dealt round-robin to the ranks of a ``torch.distributed`` process group, every rank
evaluates its (n_q, n_ls/P) block with the single-GPU path, and ONE all-gather of the FP64 blocks rebuilds
the grid on every rank; the max-shift normalisation then runs on the device.  There is no other exchange.
Each cell is computed by exactly one rank with the same kernels, so the gathered grid is bit-identical to
the single-GPU grid.

torch is used for the plumbing only (process group, device buffers for NCCL).
"""
from __future__ import annotations

import numpy as np

from . import ops

__all__ = ["shard_indices", "assemble_blocks", "lml_grid_sharded"]


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
    block = np.full((n_q, per), -np.inf)
    if len(mine):
        evaluator = kw.pop("_evaluator", ops.lml_grid)
        block[:, :len(mine)] = evaluator(X, dy, ref, orders, ls[mine], Q, **kw)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    send = torch.from_numpy(block).to(dev)
    recv = torch.empty((world * n_q, per), dtype=torch.float64, device=dev)   # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)     # the single collective of the path
    full = assemble_blocks(list(recv.view(world, n_q, per).cpu().numpy()), n_ls, world)
    if normalize:
        post, lse = ops.grid_normalize(full)
        return full, post, lse
    return full
