import os, sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
import gsum_b200 as gb
from oracle import gsum_oracle as o
rs = np.random.RandomState(2)
g1 = np.linspace(0, 1, 50); X = o.cartesian(g1, g1); n = len(X)
kern = RBF([0.02, 0.03], 'fixed') + WhiteKernel(1e-6, 'fixed')
coeffs = np.linalg.cholesky(RBF([0.02, 0.03])(X) + 1e-8 * np.eye(n)) @ rs.randn(n, 6)
orders = np.arange(6); y = o.partials(coeffs, 0.4, 1.0, orders)
gp = gb.TruncationGP(kern, ratio=0.4, ref=1, center=0, disp=0, df=1, scale=1, optimizer=None)
for _ in range(3):
    t0 = time.perf_counter(); gp.fit(X, y, orders=orders); print("fit ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable(); gp.fit(X, y, orders=orders); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
