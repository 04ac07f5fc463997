import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSUM_B200_SCHEDULE"] = "hetero"
from bench import make_inputs, N_POINTS
from gsum_b200 import ops
from gsum_b200.helpers import _order_differences
X, y, orders, ls_vals, q_vals = make_inputs(128)
dy = np.ascontiguousarray(_order_differences(y))
for rep in range(3):
    ll, logdet, status = ops.lml_grid(X, dy, np.ones(N_POINTS), orders, ls_vals[:, None], q_vals, noise=1e-6, nugget=1e-10, return_status=True)
    bad = np.where(~np.isfinite(ll).all(0))[0]
    print("rep", rep, "bad columns", len(bad), "status nonzero", np.count_nonzero(status), "logdet nonfinite", np.count_nonzero(~np.isfinite(logdet)))
    print("  status of bad:", status[bad][:12], "logdet of bad:", logdet[bad][:6], " ll sample", ll[0, bad[:4]] if len(bad) else None)
# plain batched cholesky of the same matrices
from sklearn.gaussian_process.kernels import RBF
Xs = X
A = np.stack([RBF(l)(Xs) + (1e-6 + 1e-10) * np.eye(N_POINTS) for l in ls_vals[::8]])
L, info, ld = ops.cholesky(A.copy(), return_info=True); print("info", info)
Lr = np.linalg.cholesky(A)
print("batched cholesky: finite", np.isfinite(L).all(), "max abs diff vs numpy", np.nanmax(np.abs(L - Lr)))
