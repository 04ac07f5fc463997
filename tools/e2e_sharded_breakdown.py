"""Dev probe (run under torchrun): host-side time of the sharded grid API per section, per rank.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_sharded_breakdown.py
"""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(50, exit=True)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
from bench import make_inputs
import gsum_b200 as gb
from gsum_b200 import distributed as D

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
X, y, orders, ls_vals, q_vals = make_inputs(int(os.environ.get("GSUM_PROBE_NLS", "128")))
gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
for i in range(5):
    gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=dist.group.WORLD)
    print(f"rank{rank} warm call {i} done", flush=True)
D._PROF = {}
dist.barrier(); torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=dist.group.WORLD)
dt = (time.perf_counter() - t0) / n * 1e3
prof = {k: v / n * 1e3 for k, v in D._PROF.items()}
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: {dt:.3f} ms/call | in _sharded_device: " + "  ".join(f"{k} {v:.3f}" for k, v in prof.items()) +
              f" | facade outside it {dt - sum(prof.values()):.3f}", flush=True)
D.release_graphs()
dist.destroy_process_group()
