"""Dev probe: where the end-to-end (numpy in / numpy out) time of one C4 grid goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
from bench import make_inputs, N_POINTS, N_Q
import gsum_b200 as gb
from gsum_b200 import _lib, ops
from gsum_b200.helpers import _order_differences
X, y, orders, ls_vals, q_vals = make_inputs(128)
dy = np.ascontiguousarray(_order_differences(y))
detf = N_POINTS * float(orders.sum()) * np.log(np.abs(q_vals))
gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
gp.fit(X, y, orders=orders)
def timeit(f, n=30):
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
kw = dict(constant=1.0, noise=1e-6, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0, scale0=1.0)
print("facade  gp.log_marginal_likelihood_grid : %.3f ms" % timeit(lambda: gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)))
print("ops.lml_grid (host buffers)              : %.3f ms" % timeit(lambda: ops.lml_grid(X, dy, 1.0, orders, ls_vals[:, None], q_vals, detf=detf, **kw)))
dev = torch.device("cuda", 0)
ctx = _lib.default_context(0)
t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
dX, ddy, dref, dord = t(X), t(dy), t(np.ones(N_POINTS)), t(orders.astype(np.int32), torch.int32)
dls, dQ, ddetf = t(ls_vals[:, None]), t(q_vals), t(detf)
ll = torch.empty((N_Q, 128), dtype=torch.float64, device=dev)
def dev_only():
    ops.lml_grid_device(ctx, dX, ddy, dref, dord, dls, dQ, ddetf, ll, **kw)
    ctx.synchronize()
print("ops.lml_grid_device + sync               : %.3f ms" % timeit(dev_only))
def dev_d2h():
    ops.lml_grid_device(ctx, dX, ddy, dref, dord, dls, dQ, ddetf, ll, **kw)
    ctx.synchronize()
    return ll.cpu()
print("ops.lml_grid_device + sync + D2H         : %.3f ms" % timeit(dev_d2h))
