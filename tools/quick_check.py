"""Developer smoke: raw C-ABI calls vs numpy on a GPU box (not part of the test-suite)."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from gsum_b200 import _lib
from oracle import gsum_oracle as o
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as CK
ctx = _lib.Context(0); lib = ctx.lib
def kernel_matrix(X1, X2, ls, const, noise):
    X1 = _lib.as_f64(X1); n1, d = X1.shape
    ls = np.atleast_1d(np.asarray(ls, float))
    if X2 is None:
        out = np.empty((n1, n1)); rc = lib.gsum_kernel_matrix(ctx.handle, X1.ctypes.data, n1, None, 0, d, ls.ctypes.data, len(ls), const, noise, out.ctypes.data, 0)
    else:
        X2 = _lib.as_f64(X2); out = np.empty((n1, len(X2))); rc = lib.gsum_kernel_matrix(ctx.handle, X1.ctypes.data, n1, X2.ctypes.data, len(X2), d, ls.ctypes.data, len(ls), const, noise, out.ctypes.data, 0)
    ctx.check(rc, 'kernel_matrix'); return out
rs = np.random.RandomState(0)
for n, d in [(50,1),(200,1),(130,2)]:
    X = rs.rand(n, d); ls = [0.2] if d==1 else [0.2,0.3]
    k = CK(1.5)*RBF(ls if d>1 else 0.2)+WhiteKernel(1e-6)
    K = kernel_matrix(X, None, ls, 1.5, 1e-6); Kr = k(X)
    print('K sym', n, d, np.abs(K-Kr).max(), np.array_equal(np.diag(K), np.diag(Kr)))
    X2 = rs.rand(77, d); K = kernel_matrix(X, X2, ls, 1.5, 1e-6); Kr = k(X, X2)
    print('K cross', np.abs(K-Kr).max())
# cholesky
for n, batch in [(50,3),(64,2),(200,4),(1024,2)]:
    X = np.linspace(0,1,n)[:,None]
    A = np.stack([RBF(0.05*(b+1))(X)+1e-6*np.eye(n) for b in range(batch)])
    L = A.copy(); info = np.zeros(batch, np.int32); logdet = np.zeros(batch)
    rc = lib.gsum_cholesky(ctx.handle, L.ctypes.data, n, batch, info.ctypes.data, logdet.ctypes.data, 0); ctx.check(rc,'chol')
    Lr = np.linalg.cholesky(A)
    print('chol', n, batch, 'info', info, 'max|dL|', np.abs(L-Lr).max(), 'recon', np.abs(L@np.swapaxes(L,1,2)-A).max(), 'logdet rel', np.abs(logdet - 2*np.log(np.diagonal(Lr,axis1=1,axis2=2)).sum(1)).max())
# non-PD
A = np.eye(100); A[70,70] = -1.0; L = A[None].copy(); info = np.zeros(1, np.int32)
lib.gsum_cholesky(ctx.handle, L.ctypes.data, 100, 1, info.ctypes.data, None, 0); print('nonPD info (expect 71):', info)
# lml grid
def grid(X, y, orders, ls_vals, q_vals, ref, pri, const=1.0, noise=1e-6, nugget=1e-10, student=0, qx=None):
    n, d = X.shape; n_c = y.shape[1]
    dy = np.ascontiguousarray(np.insert(np.diff(y, axis=-1), 0, y[:,0], axis=-1))
    refx = ref*np.ones(n); orders = np.asarray(orders, np.int32)
    ls = np.ascontiguousarray(np.asarray(ls_vals, float).reshape(len(ls_vals), -1))
    if qx is None:
        Q = np.ascontiguousarray(q_vals, float); detf = np.array([np.sum(n_c*np.log(np.abs(refx)) + orders.sum()*np.log(abs(q))) for q in q_vals]); xdep=0
    else:
        Q = np.ascontiguousarray(qx, float); detf = np.array([np.sum(n_c*np.log(np.abs(refx)) + orders.sum()*np.log(np.abs(qq))) for qq in qx]); xdep=1
    ll = np.empty((len(Q), len(ls))); logdet = np.empty(len(ls)); status = np.zeros(len(ls), np.int32)
    rc = lib.gsum_lml_grid(ctx.handle, X.ctypes.data, n, d, dy.ctypes.data, n_c, refx.ctypes.data, orders.ctypes.data, ls.ctypes.data, len(ls), ls.shape[1],
        Q.ctypes.data, len(Q), xdep, detf.ctypes.data, const, noise, nugget, pri['center'], pri['disp'], pri['df'], pri['scale'], student, ll.ctypes.data, logdet.ctypes.data, status.ctypes.data, 0)
    ctx.check(rc, 'lml_grid'); return ll, logdet, status
n = 200; X = np.linspace(0,1,n)[:,None]; orders = np.arange(6)
from scipy import stats
Kt = RBF(0.2)(X)+1e-6*np.eye(n)
coeffs = stats.multivariate_normal(np.zeros(n), Kt, allow_singular=True).rvs(6, random_state=1).T
y = o.partials(coeffs, 0.5, 1.0, orders)
ls_vals = np.linspace(0.02,0.5,8); q_vals = np.linspace(0.3,0.7,5)
kern = RBF(0.2)+WhiteKernel(1e-6,'fixed')
for pri in [dict(center=0,disp=0,df=1,scale=1), dict(center=0.3,disp=1,df=3,scale=0.7), dict(center=0.1,disp=0,df=np.inf,scale=1.3)]:
    for student in [0,1]:
        ll, logdet, status = grid(X, y, orders, ls_vals, q_vals, 1.0, pri, student=student)
        ref = o.lml_grid(kern, X, y, orders, ls_vals, q_vals, 1.0, o.Priors(**pri), student=bool(student))
        with np.errstate(invalid='ignore'):
            print('grid', pri, student, 'maxrel', np.nanmax(np.abs(ll-ref)/np.abs(ref)), 'nan match', np.array_equal(np.isnan(ll), np.isnan(ref)), status.max())
# x-dependent Q
qfun = lambda X, lam: (0.2+0.4*X[:,0])/lam
lams = [0.8, 1.0, 1.3]
qx = np.stack([qfun(X, l) for l in lams])
pri = dict(center=0.3,disp=1,df=3,scale=0.7)
ll, _, _ = grid(X, y, orders, ls_vals, None, 1.0, pri, qx=qx)
ref = np.array([[o.truncation_lml(kern, [np.log(l)], X, y, orders, qq, np.ones(n), o.Priors(**pri)) for l in ls_vals] for qq in qx])
print('grid xdep maxrel', np.abs(ll-ref).max()/np.abs(ref).max(), np.abs((ll-ref)/ref).max())
# timing C4
n = 1024; X = np.linspace(0,1,n)[:,None]
Kt = RBF(0.05)(X)+1e-6*np.eye(n)
coeffs = stats.multivariate_normal(np.zeros(n), Kt, allow_singular=True).rvs(6, random_state=3).T
y = o.partials(coeffs, 0.5, 1.0, orders)
ls_vals = np.geomspace(0.005,0.5,128); q_vals = np.linspace(0.2,0.8,256)
pri = dict(center=0,disp=0,df=1,scale=1)
for it in range(4):
    t0 = time.perf_counter(); ll, logdet, status = grid(X, y, orders, ls_vals, q_vals, 1.0, pri); t1 = time.perf_counter()
    print('C4 grid e2e ms', (t1-t0)*1e3, 'status fails', (status!=0).sum(), 'launches', ctx.launch_count)
idx = [(0,0),(255,127),(100,64),(17,90)]
for (a,b) in idx:
    r = o.truncation_lml(kern, [np.log(ls_vals[b])], X, y, orders, q_vals[a]*np.ones(n), np.ones(n), o.Priors(**pri))
    print('C4 cell', a, b, ll[a,b], r, abs(ll[a,b]-r)/abs(r))
