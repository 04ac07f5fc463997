"""Dev tool: time the C4 grid (device-resident inputs) and its factorisation bracket.

    python tools/perf_chol.py [--nls 128] [--reps 10] [--stats] [nochain]      (nochain: GSUM_B200_CHAIN_MAX=0 beside the default)
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs, N_POINTS, N_Q
from gsum_b200 import _lib, ops
from gsum_b200.helpers import _order_differences

modes = ["default"] + (["nochain"] if "nochain" in sys.argv[1:] else [])
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
n_ls = int(sys.argv[sys.argv.index("--nls") + 1]) if "--nls" in sys.argv else 128
dev = torch.device("cuda", 0)
X, y, orders, ls_vals, q_vals = make_inputs(n_ls)
dy = np.ascontiguousarray(_order_differences(y))
detf = N_POINTS * float(orders.sum()) * np.log(np.abs(q_vals))
t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
dX, ddy, dref, dord = t(X), t(dy), t(np.ones(N_POINTS)), t(orders.astype(np.int32), torch.int32)
dls, dQ, ddetf = t(ls_vals[:, None]), t(q_vals), t(detf)
kw = dict(constant=1.0, noise=1e-6, nugget=1e-10, center0=0.0, disp0=0.0, df0=1.0, scale0=1.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
results = {}
for mode in modes:
    if mode == "nochain":
        os.environ["GSUM_B200_CHAIN_MAX"] = "0"
    else:
        os.environ.pop("GSUM_B200_CHAIN_MAX", None)
    if "--stats" in sys.argv:
        os.environ["GSUM_B200_DF_STATS"] = "1"
    else:
        os.environ.pop("GSUM_B200_DF_STATS", None)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = _lib.Context(0, stream.cuda_stream)
        ll = torch.empty((N_Q, n_ls), dtype=torch.float64, device=dev)
        for _ in range(3):
            ops.lml_grid_device(ctx, dX, ddy, dref, dord, dls, dQ, ddetf, ll, **kw)
        torch.cuda.synchronize()
        os.environ.pop("GSUM_B200_DF_STATS", None)
        ctx.profile(True)
        tot = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ops.lml_grid_device(ctx, dX, ddy, dref, dord, dls, dQ, ddetf, ll, **kw); e1.record(stream)
            e1.synchronize()
            tot.append(e0.elapsed_time(e1))
        ms, fl, nb = ctx.profile_read()
        results[mode] = ll.cpu().numpy().copy()
        import hashlib
        print("ll sha1", hashlib.sha1(results[mode].tobytes()).hexdigest()[:12], flush=True)
        print(f"{mode:12s} grid ms: min {min(tot):.3f} med {np.median(tot):.3f} | factor bracket {ms / nb:.3f} ms  "
              f"{fl / ms * 1e-9:.2f} TFLOP/s (algorithmic)  launches/step {ctx.launch_count // (reps + 3)}", flush=True)
        ctx.close()
names = list(results)
for other in names[1:]:
    a, b = results[names[0]], results[other]
    print(f"{names[0]} vs {other}: bit-identical {np.array_equal(a, b)}  max rel diff {np.max(np.abs(a - b) / np.abs(b)):.3e}")
