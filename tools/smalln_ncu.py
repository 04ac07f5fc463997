"""Dev probe for ncu: a few C2-size grids (N = 200, 64 x 64) on the small-N path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsum_b200 import ops
n = 200
rs = np.random.RandomState(1)
X = np.linspace(0, 1, n)[:, None]
dy = rs.randn(n, 6)
ls_vals, q_vals = np.linspace(0.02, 0.5, 64), np.linspace(0.3, 0.7, 64)
for _ in range(3):
    ll = ops.lml_grid(X, dy, 1.0, np.arange(6), ls_vals[:, None], q_vals, constant=1.0, noise=1e-6, nugget=1e-10, center0=0., disp0=0., df0=1., scale0=1.)
print(ll[0, :3])
