"""numpy emulation of the one-sided Jacobi iteration of gsum_eigh (csrc/eig.cuh) under different pair orderings — the CPU
study behind GSUM_B200_EIGH_ORDER=modulus.  Counts sweeps and rotations per sweep for: the round-robin tournament (the
kernel's default), the modulus ordering (round s pairs positions i + j = s mod n), the modulus ordering on positions sorted
by decreasing column norm at the start of every sweep, each on A itself and on its pivoted-Cholesky factor, and (n <= 160)
the sequential row-cyclic order with de Rijk pivoting as the reference point.  Strict relative criterion sqrt(n) eps.

    python tools/jacobi_order_emulation.py N [length_scale] [noise]        (N = 512 takes ~2 minutes)

Results at N = 512, RBF(0.05) + 1e-4 I (session 5): A round-robin 24 sweeps, A modulus 24, A modulus + sort 17;
factor round-robin 18, factor modulus 16, factor modulus + sort 13.  The device run of the round-robin case takes the same
24 sweeps with the purely relative criterion, so the emulation is faithful."""
import numpy as np, sys, time
from sklearn.gaussian_process.kernels import RBF
from scipy.linalg import lapack

def pairs_rr(npad, r):
    m = npad - 1
    k = np.arange(1, npad // 2)
    p = np.concatenate([[m], (r + k) % m]); q = np.concatenate([[r], (r - k + m) % m])
    lo, hi = np.minimum(p, q), np.maximum(p, q)
    return lo, hi

def rot_params(a, b, g):
    zeta = (b - a) / (2 * g)
    t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
    t = np.where(zeta == 0, 1.0, t)
    c = 1 / np.sqrt(1 + t * t)
    return c, c * t

def jacobi_rr(G, tol, max_sweeps=60, derijk=False, sort_each=False):
    """rows of G are the columns being orthogonalised; vectorised round-robin"""
    n = G.shape[0]
    npad = n + (n & 1)
    counts = []
    for sweep in range(max_sweeps):
        if sort_each:
            G = G[np.argsort(-np.einsum('ij,ij->i', G, G))]
        rot = 0
        for r in range(npad - 1):
            p, q = pairs_rr(npad, r)
            ok = q < n
            p, q = p[ok], q[ok]
            x, y = G[p], G[q]
            a = np.einsum('ij,ij->i', x, x); b = np.einsum('ij,ij->i', y, y); g = np.einsum('ij,ij->i', x, y)
            do = np.abs(g) > tol * np.sqrt(a * b)
            if not do.any():
                continue
            rot += int(do.sum())
            gs = np.where(do, g, 1.0)
            c, s = rot_params(a, b, gs)
            c = np.where(do, c, 1.0); s = np.where(do, s, 0.0)
            xn = c[:, None] * x - s[:, None] * y
            yn = s[:, None] * x + c[:, None] * y
            if derijk:
                sw = do & (np.einsum('ij,ij->i', xn, xn) < np.einsum('ij,ij->i', yn, yn))
                xn2 = np.where(sw[:, None], yn, xn); yn = np.where(sw[:, None], xn, yn); xn = xn2
            G[p] = xn; G[q] = yn
        counts.append(rot)
        if rot == 0:
            break
    return G, counts

def jacobi_cyclic(G, tol, max_sweeps=60, derijk=False):
    n = G.shape[0]
    counts = []
    for sweep in range(max_sweeps):
        rot = 0
        for p in range(n - 1):
            if derijk:
                nr = np.einsum('ij,ij->i', G[p:], G[p:])
                j = p + int(np.argmax(nr))
                if j != p:
                    G[[p, j]] = G[[j, p]]
            for q in range(p + 1, n):
                x, y = G[p], G[q]
                a = x @ x; b = y @ y; g = x @ y
                if abs(g) > tol * np.sqrt(a * b):
                    rot += 1
                    c, s = rot_params(a, b, g)
                    G[p], G[q] = c * x - s * y, s * x + c * y
        counts.append(rot)
        if rot == 0:
            break
    return G, counts


import numpy as np, sys, time
from sklearn.gaussian_process.kernels import RBF
from scipy.linalg import lapack

def pairs_mod(n, s):
    i = np.arange(n)
    j = (s - i) % n
    m = i < j
    return i[m], j[m]

def jacobi_par(G, tol, pairs_fn, nrounds, max_sweeps=60, sort_each=False, tol_first=None):
    n = G.shape[0]
    counts = []
    for sweep in range(max_sweeps):
        if sort_each:
            G = G[np.argsort(-np.einsum('ij,ij->i', G, G))]
        rot = 0
        for r in range(nrounds):
            p, q = pairs_fn(r)
            x, y = G[p], G[q]
            a = np.einsum('ij,ij->i', x, x); b = np.einsum('ij,ij->i', y, y); g = np.einsum('ij,ij->i', x, y)
            do = np.abs(g) > tol * np.sqrt(a * b)
            if not do.any():
                continue
            rot += int(do.sum())
            gs = np.where(do, g, 1.0)
            c, s = rot_params(a, b, gs)
            c = np.where(do, c, 1.0); s = np.where(do, s, 0.0)
            G[p], G[q] = c[:, None] * x - s[:, None] * y, s[:, None] * x + c[:, None] * y
        counts.append(rot)
        if rot == 0:
            break
    return G, counts

n = int(sys.argv[1])
ls = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
X = np.linspace(0, 1, n)[:, None]
A = RBF(ls)(X) + noise * np.eye(n)
tol = np.sqrt(n) * 2.2e-16
wl = np.linalg.eigvalsh(A)
print("n", n, "ls", ls, "cond %.1e" % (wl[-1] / wl[0]), "eigs > 10*noise:", int((wl > 10 * noise).sum()))
c, piv, rank, info = lapack.dpstrf(A, lower=1)
Lp = np.tril(c); P = np.zeros((n, n)); P[piv - 1, np.arange(n)] = 1
F = P @ Lp
npad = n + (n & 1)
def rr(r):
    p, q = pairs_rr(npad, r); ok = q < n; return p[ok], q[ok]
for name, fn in [
    ("direct A, round-robin", lambda: jacobi_par(A.copy(), tol, rr, npad - 1)),
    ("direct A, modulus", lambda: jacobi_par(A.copy(), tol, lambda s: pairs_mod(n, s), n)),
    ("direct A, modulus + sort", lambda: jacobi_par(A.copy(), tol, lambda s: pairs_mod(n, s), n, sort_each=True)),
    ("pchol, round-robin", lambda: jacobi_par(F.T.copy(), tol, rr, npad - 1)),
    ("pchol, modulus", lambda: jacobi_par(F.T.copy(), tol, lambda s: pairs_mod(n, s), n)),
    ("pchol, modulus + sort", lambda: jacobi_par(F.T.copy(), tol, lambda s: pairs_mod(n, s), n, sort_each=True)),
]:
    t0 = time.time()
    G, counts = fn()
    w = np.sort(np.einsum('ij,ij->i', G, G) ** (0.5 if name.startswith("direct") else 1.0))
    print(f"{name:28s} sweeps {len(counts):3d}  err {np.max(np.abs(w - wl)) / wl[-1]:.1e}  rotations {counts}  ({time.time() - t0:.0f}s)", flush=True)
if n <= 160:
    for name, fn in [("pchol, row-cyclic + de Rijk", lambda: jacobi_cyclic(F.T.copy(), tol, derijk=True)), ("direct A, row-cyclic", lambda: jacobi_cyclic(A.copy(), tol))]:
        G, counts = fn()
        print(f"{name:28s} sweeps {len(counts):3d}  rotations {counts}", flush=True)
