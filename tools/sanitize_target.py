"""Target of tools/sanitize.sh: one small call per kernel family of the factorisation path (many-matrices schedule, chain mode,
small-N kernel, solve-only border pass, pivoted Cholesky panel), checked against numpy so that a sanitizer-clean run is also a correct one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.gaussian_process.kernels import RBF
from gsum_b200 import ops

which = sys.argv[1:] or ["hetero", "chain", "smalln", "solve", "pstrf"]
rs = np.random.RandomState(0)


def mats(n, batch):
    X = np.sort(rs.rand(n))[:, None]
    return np.stack([RBF(0.05 * (b % 7 + 1))(X) + 1e-4 * np.eye(n) for b in range(batch)])


def check(name, got, want, tol=1e-9):
    err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    print(f"{name}: rel err {err:.2e}", flush=True)
    assert err < tol, name


if "hetero" in which:          # 44 matrices > chain threshold: GEMM CTAs + factor CTAs, TMA rings, flags
    A = mats(200, 44)
    L = ops.cholesky(A)
    check("hetero cholesky 44 x 200", L, np.linalg.cholesky(A))
if "chain" in which:           # 3 matrices: chain workers own the diagonal band
    A = mats(330, 3)
    L = ops.cholesky(A)
    check("chain cholesky 3 x 330", L, np.linalg.cholesky(A))
if "smalln" in which or "grid" in which:
    for n, n_ls in ((130, 5), (300, 5), (300, 44)):          # small-N kernel; chain mode + thin border; many-matrices + thin border
        X = np.linspace(0, 1, n)[:, None]
        dy = rs.randn(n, 4)
        ls_vals, q_vals = np.linspace(0.05, 0.4, n_ls), np.linspace(0.3, 0.7, 3)
        ll = ops.lml_grid(X, dy, 1.0, np.arange(4), ls_vals[:, None], q_vals, constant=1.0, noise=1e-4, nugget=1e-10, center0=0., disp0=0., df0=1., scale0=1.)
        assert np.isfinite(ll).all()
        print(f"grid N={n} n_ls={n_ls}: ok {ll[0, 0]:.6f}", flush=True)
if "solve" in which:
    A = mats(200, 1)[0]
    B = rs.randn(200, 70)
    Lc = np.linalg.cholesky(A)
    check("cho_solve 200 x 70", ops.cho_solve(Lc, B), np.linalg.solve(A, B), 1e-7)
if "pstrf" in which:
    A = mats(200, 1)[0]
    G, Lp, piv, rank, status = ops.pivoted_cholesky(A)
    check("pivoted cholesky 200", G @ G.T, A, 1e-12)
print("sanitize target: done", flush=True)
