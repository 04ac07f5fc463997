"""Dev check: C4 grid under the hetero schedule, repeated, against the pipeline schedule; reports non-finite cells."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
import gsum_b200 as gb
from gsum_b200 import _lib
X, y, orders, ls_vals, q_vals = make_inputs(128)
def grid(mode, reps=1):
    os.environ["GSUM_B200_SCHEDULE"] = mode
    _lib._default_ctx.clear()
    gp = gb.TruncationGP(RBF(0.05) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
    gp.fit(X, y, orders=orders)
    return [gp.log_marginal_likelihood_grid(ls_vals, ratio_vals=q_vals) for _ in range(reps)]
mode = os.environ.get("HT_MODE", "hetero")
ref = grid("pipeline")[0]
print("pipeline finite:", np.isfinite(ref).all())
outs = grid(mode, int(sys.argv[1]) if len(sys.argv) > 1 else 6)
for r, o in enumerate(outs):
    bad = ~np.isfinite(o)
    d = np.abs(o - ref) / np.abs(ref)
    d[bad] = 0
    print(f"rep {r}: non-finite cells {bad.sum()} (ls columns {sorted(set(np.where(bad)[1]))[:10]}), max rel diff on the rest {d.max():.3e}, identical to rep 0: {np.array_equal(o, outs[0], equal_nan=True)}")
