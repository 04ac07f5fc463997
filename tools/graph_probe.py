"""Dev probe: the sharded grid on a ONE-rank NCCL group, several calls (the third replays the captured CUDA graph)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
from bench import make_inputs
import gsum_b200 as gb
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29611")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
n_ls = int(sys.argv[1]) if len(sys.argv) > 1 else 16
X, y, orders, ls_vals, q_vals = make_inputs(n_ls)
gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
want = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)
for i in range(6):
    t0 = time.perf_counter()
    got = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals, group=dist.group.WORLD)
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.3f} ms, equal {np.array_equal(got, want)}", flush=True)
from gsum_b200 import distributed as D
D.release_graphs()
dist.destroy_process_group()
