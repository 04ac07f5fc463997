"""Dev tool: C4 grid under two schedules; the cells where they differ most are arbitrated by x87 extended precision."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bench import make_inputs
from util import lml_extended_precision
from oracle import gsum_oracle as o
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
import gsum_b200 as gb
from gsum_b200 import _lib

modes = [a for a in sys.argv[1:] if not a.startswith("-")] or ["hetero", "pipeline"]
X, y, orders, ls_vals, q_vals = make_inputs(128)
res = {}
for mode in modes:
    os.environ["GSUM_B200_SCHEDULE"] = mode
    _lib._default_ctx.clear()                 # a fresh context picks the schedule up from the environment
    gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
    gp.fit(X, y, orders=orders)
    res[mode] = gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals)
a, b = res[modes[0]], res[modes[-1]]
rel = np.abs(a - b) / np.abs(b)
print("max rel diff", rel.max(), "median", np.median(rel))
idx = np.dstack(np.unravel_index(np.argsort(-rel.ravel())[:4], rel.shape))[0]
per_ls = rel.max(0)
print("per-ls max rel diff (every 8th):", np.array2string(per_ls[::8], precision=1))
kern = RBF(0.05) + WhiteKernel(1e-6, 'fixed')
n = len(X)
for (qa, lb) in list(map(tuple, idx)) + [(10, 40), (200, 90), (128, 127)]:
    q = q_vals[qa]
    c = o.coefficients(y, q, 1.0, orders)
    t0 = time.time()
    ex = lml_extended_precision(X, c, [ls_vals[lb]], 1e-6, 1e-10, 0.0, 0.0, 1.0, 1.0) - n * orders.sum() * np.log(q)
    ref = o.truncation_lml(kern, [np.log(ls_vals[lb])], X, y, orders, q * np.ones(n), np.ones(n), o.Priors(0, 0, 1, 1))
    print(f"cell q[{qa}] ls[{lb}]={ls_vals[lb]:.4f}: exact {ex:.10f} | " + " | ".join(f"{m} err {abs(res[m][qa, lb] - ex) / abs(ex):.2e}" for m in modes) +
          f" | reference err {abs(ref - ex) / abs(ex):.2e}  ({time.time() - t0:.1f}s)", flush=True)
