"""Who is closer to the exact value?  Device grid vs reference (golden) vs x87 extended precision, config C2 prior 3 (df0 = inf)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import lml_extended_precision, prior_kwargs
from oracle import gsum_oracle as o
import gsum_b200 as gb
from gsum_b200 import ops
from sklearn.gaussian_process.kernels import RBF, WhiteKernel
g = dict(np.load('tests/golden/c2_truncation_grid.npz'))
X, y, orders = g['X'], g['y'], g['orders']
for ip in [3, 0]:
    pk = prior_kwargs(g['priors'][ip])
    gp = gb.TruncationGP(RBF(0.2) + WhiteKernel(1e-6, 'fixed'), ratio=0.5, ref=1, optimizer=None, **pk).fit(X, y, orders=orders)
    ll = gp.log_marginal_likelihood_grid(g['ls_vals'], ratio_vals=g['q_vals'])
    want = g[f'g{ip}_ll']
    em, er = [], []
    for a in range(0, 8, 2):
        for b in range(8):
            q = g['q_vals'][a]
            c = o.coefficients(y, q, 1.0, orders)
            ex = lml_extended_precision(X, c, [g['ls_vals'][b]], 1e-6, 1e-10, pk['center'], pk['disp'], pk['df'], pk['scale']) - len(X) * orders.sum() * np.log(q)
            em.append(abs(ll[a, b] - ex) / abs(ex)); er.append(abs(want[a, b] - ex) / abs(ex))
    em, er = np.array(em), np.array(er)
    print('prior', ip, 'device: median %.2e max %.2e | reference: median %.2e max %.2e | ratio of medians %.2f' % (np.median(em), em.max(), np.median(er), er.max(), np.median(em) / np.median(er)))
    print('  per-ls device ', np.array2string(em.reshape(4, 8).max(0), precision=1))
    print('  per-ls ref    ', np.array2string(er.reshape(4, 8).max(0), precision=1))
# kernel-matrix entry accuracy in ulps vs extended precision
ld = np.longdouble
for l in [0.02, 0.2, 0.5]:
    Xs = X.astype(ld) / ld(l)
    Rex = np.exp(ld(-0.5) * (Xs - Xs.T) ** 2)
    Rd = ops.kernel_matrix(X, None, [l], 1.0, 0.0)
    Rs = RBF(l)(X)
    ulp = np.spacing(Rs)
    print('ls', l, 'max ulp err device %.2f sklearn %.2f ; rms device %.3f sklearn %.3f' % (
        np.max(np.abs(Rd.astype(ld) - Rex) / ulp), np.max(np.abs(Rs.astype(ld) - Rex) / ulp),
        np.sqrt(np.mean((np.abs(Rd.astype(ld) - Rex) / ulp).astype(float) ** 2)), np.sqrt(np.mean((np.abs(Rs.astype(ld) - Rex) / ulp).astype(float) ** 2))))
# factor accuracy: |L L^T - R| and forward error of L vs extended precision for l = 0.5
l = 0.5
R = RBF(l)(X) + 1e-6 * np.eye(len(X))
Lr = np.linalg.cholesky(R); Ld = ops.cholesky(R)
Rl = R.astype(ld); n = len(X); Lx = np.zeros_like(Rl)
for j in range(n):
    Lx[j, j] = np.sqrt(Rl[j, j] - (Lx[j, :j] ** 2).sum()); Lx[j + 1:, j] = (Rl[j + 1:, j] - Lx[j + 1:, :j] @ Lx[j, :j]) / Lx[j, j]
print('L forward err: device %.2e lapack %.2e' % (np.abs(Ld - Lx).max(), np.abs(Lr - Lx).max()), ' backward: device %.2e lapack %.2e' % (np.abs(Ld @ Ld.T - R).max(), np.abs(Lr @ Lr.T - R).max()))
