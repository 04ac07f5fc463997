"""Dev probe: the C2 grid (N = 200, 64 x 64) on the small-N path vs the tile path: device time and agreement, and the oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
import gsum_b200 as gb
from gsum_b200 import _lib
from oracle import gsum_oracle as o
def run(n, nls, nq, label):
    rs = np.random.RandomState(1)
    X = np.linspace(0, 1, n)[:, None]
    coeffs = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, 6)
    orders = np.arange(6); y = o.partials(coeffs, 0.5, 1.0, orders)
    ls_vals, q_vals = np.linspace(0.02, 0.5, nls), np.linspace(0.3, 0.7, nq)
    kern = RBF(0.2) + WhiteKernel(1e-6, 'fixed')
    out = {}
    for mode in ("1", "0"):
        os.environ["GSUM_B200_SMALLN"] = mode
        ctx = _lib.Context(0)
        _lib._default = None
        import gsum_b200.ops as ops
        gp = gb.TruncationGP(kern, ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None).fit(X, y, orders=orders)
        f = lambda: ops.lml_grid(gp._grid_inputs(None, None, None)[0], gp._grid_inputs(None, None, None)[1], 1.0, orders, ls_vals[:, None], q_vals,
                                 detf=n * orders.sum() * np.log(q_vals), constant=1.0, noise=1e-6, nugget=1e-10, center0=0., disp0=0., df0=1., scale0=1., ctx=ctx)
        ll = f(); ctx.synchronize()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); ll = f(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
        out[mode] = ll
        print(f"{label} smalln={mode}: {min(ts) * 1e3:.3f} ms per grid (host buffers in/out)", flush=True)
        ctx.close()
    print(f"{label}: small vs tile path max rel diff {np.max(np.abs(out['1'] - out['0']) / np.abs(out['0'])):.2e}")
    ref = np.array([[o.truncation_lml(kern, [np.log(l)], X, y, orders, q * np.ones(n), np.ones(n), o.Priors(0, 0, 1, 1)) for l in ls_vals[::8]] for q in q_vals[::8]])
    print(f"{label}: small path vs oracle on {ref.size} cells: {np.max(np.abs(out['1'][::8, ::8] - ref) / np.abs(ref)):.2e}; tile path: {np.max(np.abs(out['0'][::8, ::8] - ref) / np.abs(ref)):.2e}")
run(200, 64, 64, "C2 N=200 64x64")
run(50, 16, 16, "C1-size N=50 16x16")
run(203, 9, 5, "N=203")

# device-resident timing (CUDA events) of the C2 grid on both paths
import torch
from gsum_b200 import ops
from gsum_b200.helpers import _order_differences
n = 200
rs = np.random.RandomState(1)
X = np.linspace(0, 1, n)[:, None]
coeffs = np.linalg.cholesky(RBF(0.2)(X) + 1e-6 * np.eye(n)) @ rs.randn(n, 6)
orders = np.arange(6); y = o.partials(coeffs, 0.5, 1.0, orders)
ls_vals, q_vals = np.linspace(0.02, 0.5, 64), np.linspace(0.3, 0.7, 64)
dev = torch.device("cuda", 0)
t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
for mode in ("1", "0"):
    os.environ["GSUM_B200_SMALLN"] = mode
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = _lib.Context(0, stream.cuda_stream)
        args = (t(X), t(_order_differences(y)), t(np.ones(n)), t(orders.astype(np.int32), torch.int32), t(ls_vals[:, None]), t(q_vals),
                t(n * orders.sum() * np.log(q_vals)))
        ll = torch.empty((64, 64), dtype=torch.float64, device=dev)
        kw = dict(constant=1.0, noise=1e-6, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0, scale0=1.0)
        for _ in range(3):
            ops.lml_grid_device(ctx, *args, ll, **kw)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ops.lml_grid_device(ctx, *args, ll, **kw); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"C2 device-resident smalln={mode}: {min(ts):.4f} ms per grid (CUDA events)", flush=True)
        ctx.close()
