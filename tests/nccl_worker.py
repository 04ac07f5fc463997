"""Worker of tests/test_gpu_multirank.py: one process per GPU under torch.distributed.run (NCCL).  Every rank evaluates the
three shardings of the path on its own GPU — the (Q, l) grid dealt over length scales + ONE all-gather, the
posterior draws dealt over the draw axis + ONE all-reduce of the coverage counts, the test points of a predict dealt in
blocks + ONE all-gather — and compares each with the unsharded single-GPU evaluation of the same call, BIT FOR BIT.
Exit code 0 = every comparison held on this rank; the JSON line of rank 0 lists what was compared."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel
    import gsum_b200 as gb
    from gsum_b200 import distributed as gdist
    from oracle import gsum_oracle as o            # input generators only (partials)

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    report = {"world": world}
    try:
        # ---- (1) the grid: N = 300 (tile path) and N = 130 (small-N path), length-scale counts that do not divide evenly
        for n, n_ls, n_q in ((300, 2 * world + 3, 9), (130, 7, 5), (1024, 4 * world, 16)):
            rs = np.random.RandomState(n)
            X = np.linspace(0, 1, n)[:, None]
            coeffs = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, 4)
            orders = np.arange(4)
            y = o.partials(coeffs, 0.5, 1.0, orders)
            gp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0.1, disp=0.5, df=3, scale=1.2,
                                 optimizer=None).fit(X, y, orders=orders)
            ls_vals, q_vals = np.linspace(0.05, 0.4, n_ls), np.linspace(0.3, 0.7, n_q)
            want = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)                       # this GPU alone
            got = [gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=dist.group.WORLD) for _ in range(4)]
            if not all(np.array_equal(g_, want) for g_ in got):
                det = [(int(np.count_nonzero(g_ != want)), float(np.max(np.abs(g_ - want))), np.argwhere(g_ != want)[:4].tolist()) for g_ in got]
                raise AssertionError(f"sharded grid differs from the 1-GPU grid (N={n}, rank {rank}): per call (cells, max abs, where) {det}")
            # and every rank holds the same bytes
            t = torch.from_numpy(np.ascontiguousarray(got[-1])).cuda()
            ref = t.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(t, ref), "ranks disagree on the gathered grid"
            report[f"grid_N{n}"] = f"{n_q}x{n_ls} cells bit-identical to 1 GPU, 4 calls (graph replay from the 3rd)"
        gdist.release_graphs()
        # ---- (2) draws: coverage counts of 2000 posterior draws, draw axis sharded
        n = 256
        Xd = np.linspace(0, 1, n)[:, None]
        cov = 1.3 * (RBF(0.2)(Xd) + 1e-5 * np.eye(n))
        d = gb.Diagnostic(np.zeros(n), cov, random_state=1)
        intervals = np.linspace(0, 1, 21)
        want = d.sample_coverage(2000, intervals, seed=123, counts=True, per_draw=False)
        want = np.asarray(want, dtype=np.float64) / (2000.0 * n)
        got = gdist.sample_coverage_sharded(d, 2000, intervals, seed=123, group=dist.group.WORLD)
        assert np.array_equal(got, want), "sharded coverage differs from the 1-GPU coverage"
        report["draws"] = "coverage of 2000 draws x 21 levels identical to 1 GPU (integer counts, one all-reduce)"
        # ---- (3) predict: 1001 test points dealt in blocks
        rs = np.random.RandomState(7)
        n = 400
        g1 = np.linspace(0, 1, 20)
        X = o.cartesian(g1, g1)
        kern = RBF([0.05, 0.07], "fixed") + WhiteKernel(1e-6, "fixed")
        coeffs = np.linalg.cholesky(RBF([0.05, 0.07])(X) + 1e-8 * np.eye(n)) @ rs.randn(n, 5)
        orders = np.arange(5)
        y = o.partials(coeffs, 0.4, 1.0, orders)
        gp = gb.TruncationGP(kern, ratio=0.4, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
        Xt = rs.rand(1001, 2)
        for proc, kw in ((gp.coeffs_process, {}), (gp, dict(order=4, kind="both"))):
            wm, ws = proc.predict(Xt, return_std=True, **kw)
            gm, gs = gdist.predict_sharded(proc, Xt, return_std=True, group=dist.group.WORLD, **kw)
            # blocks of test points go through the same kernels, but the border tile a point lands in depends on the block
            # it is dealt into; the forward solve of a row does not depend on its neighbours, so the values are identical
            assert np.array_equal(gm, wm) and np.array_equal(gs, ws), "sharded predict differs from the 1-GPU predict"
        report["predict"] = "mean/std at 1001 points (coeffs_process and TruncationGP kind=both) identical to 1 GPU"
        # ---- (4) the same grid through the library's own communicator: gsum_comm_init + gsum_grid_allgather, C ABI only
        n, n_ls, n_q = 300, 2 * world + 3, 9
        rs = np.random.RandomState(n)
        X = np.linspace(0, 1, n)[:, None]
        dy = rs.randn(n, 4)
        ls_vals, q_vals = np.linspace(0.05, 0.4, n_ls), np.linspace(0.3, 0.7, n_q)
        from gsum_b200 import ops
        kw = dict(constant=1.0, noise=1e-6, nugget=1e-10, center0=0.1, disp0=0.5, df0=3.0, scale0=1.2)
        want = ops.lml_grid(X, dy, 1.0, np.arange(4), ls_vals[:, None], q_vals, **kw)
        wpost, wlse = ops.grid_normalize(want)
        w_, r_ = gdist.cabi_comm_init(group=dist.group.WORLD)
        assert (w_, r_) == (world, rank)
        for _ in range(2):
            got, post, lse = gdist.lml_grid_sharded_cabi(X, dy, 1.0, np.arange(4), ls_vals[:, None], q_vals, world, rank, normalize=True, **kw)
            assert np.array_equal(got, want) and np.array_equal(post, wpost) and lse == wlse, "C-ABI sharded grid differs from the 1-GPU grid"
        counts = np.arange(5, dtype=np.int64) * (rank + 1)
        from gsum_b200 import _lib as glib
        ctx = glib.default_context()
        ctx.check(ctx.lib.gsum_comm_allreduce_counts(ctx.handle, counts.ctypes.data, 5, 0), "gsum_comm_allreduce_counts")
        assert np.array_equal(counts, np.arange(5) * (world * (world + 1) // 2))
        gdist.cabi_comm_destroy()
        report["cabi"] = "gsum_comm_init + gsum_grid_allgather (+ normalisation) bit-identical to 1 GPU; int64 all-reduce exact"
        ok = 1
    except AssertionError as e:
        report["failed"] = str(e)
        ok = 0
    flag = torch.tensor([ok], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["all_ranks_ok"] = bool(flag.item())
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
