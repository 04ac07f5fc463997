"""Developer smoke #2: ops (cho_solve, fit, predict, diagnostics) vs the oracle on a GPU box."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from gsum_b200 import ops
from gsum_b200._lib import PREDICT_MEAN, PREDICT_VAR, PREDICT_COV
from oracle import gsum_oracle as o
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as CK
from scipy.linalg import cho_solve, solve_triangular
def rel(a, b): return np.abs(np.asarray(a)-np.asarray(b)).max()/max(1e-300, np.abs(b).max())
rs = np.random.RandomState(0)
# cho_solve
for n, k in [(50,3),(200,70),(333,5)]:
    X = np.sort(rs.rand(n))[:,None]; A = RBF(0.1)(X)+1e-4*np.eye(n); L = np.linalg.cholesky(A); B = rs.randn(n,k)
    print('cho_solve', n, k, rel(ops.cho_solve(L,B), cho_solve((L,True),B)), 'fwd', rel(ops.cho_solve(L,B,forward_only=True), solve_triangular(L,B,lower=True)))
# fit / predict
n = 150; X = np.linspace(0,1,n)[:,None]
kern = CK(1.5,'fixed')*RBF(0.2,'fixed')+WhiteKernel(1e-6,'fixed')
from scipy import stats
coeffs = stats.multivariate_normal(np.zeros(n), RBF(0.2)(X)+1e-6*np.eye(n), allow_singular=True).rvs(5, random_state=0).T
for pri in [dict(center=0,disp=0,df=1,scale=1), dict(center=0.3,disp=1,df=3,scale=0.7)]:
  for student in [False, True]:
    f = o.fit_conjugate(kern, X, coeffs, o.Priors(**pri), nugget=1e-10, student=student)
    h = ops.FitHandle(X, coeffs, 0.2, 1.5, 1e-6, 1e-10, pri['center'], pri['disp'], pri['df'], pri['scale'], student=student, want_L=True)
    print('fit', pri, student, [rel(getattr(h,k), np.squeeze(f[k])) for k in ['center','disp','df','scale','cov_factor','lml']], 'L', rel(h.L, f['corr_L']))
    Xn = np.linspace(0,1,77)[:,None]
    pf = o.predict_student if student else o.predict_conjugate
    m0 = np.full(n, float(h.center)); m1 = np.full(77, float(h.center))
    mean, var, cb = h.predict(Xn, PREDICT_VAR, mean_old=m0, mean_new=m1, basis_old=np.ones(n), basis_new=np.ones(77), want_cond_basis=True)
    mr, sr = pf(f, Xn, return_std=True)
    std = np.sqrt(var) + (np.sqrt(h.cov_factor*h.disp)*np.abs(cb) if student else 0)
    print('  predict mean', rel(mean, mr), 'std', rel(std, sr))
    mean, cov, cb = h.predict(Xn, PREDICT_COV, mean_old=m0, mean_new=m1, basis_old=np.ones(n), basis_new=np.ones(77), want_cond_basis=True, pred_noise=True)
    mr, cr = pf(f, Xn, return_cov=True, pred_noise=True)
    if student: cov = cov + h.cov_factor*h.disp*np.outer(cb,cb)
    print('  predict cov', rel(cov, cr))
    Xc = X[::3]; yc = coeffs[::3]
    mean, var, _ = h.predict(Xn, PREDICT_VAR, Xc=Xc, yc=yc, mean_old=np.full(len(Xc), h.center), mean_new=m1)
    mr, sr = o.predict_conjugate(f, Xn, return_std=True, Xc=Xc, y=yc)
    print('  predict Xc mean', rel(mean, mr), 'std', rel(np.sqrt(var), sr))
# truncation predict
orders = np.arange(6); Xs = X[::10]; 
y = o.partials(coeffs[:, :5].repeat(1,axis=1), 0.5, 1.0, np.arange(5))
qf = lambda X: (0.2+0.4*X[:,0]); rf = lambda X: 1.0+X[:,0]
ys = o.partials(coeffs[::10], qf(Xs), rf(Xs), np.arange(5))
cs = o.coefficients(ys, qf(Xs), rf(Xs), np.arange(5))[:,1:]
pri = dict(center=0.3,disp=1,df=3,scale=0.7)
k2 = RBF(0.2,'fixed')+WhiteKernel(1e-6,'fixed')
f = o.fit_conjugate(k2, Xs, cs, o.Priors(**pri))
h = ops.FitHandle(Xs, cs, 0.2, 1.0, 1e-6, 1e-10, pri['center'], pri['disp'], pri['df'], pri['scale'])
Xn = np.linspace(0,1,41)[:,None]; order = 3
for kind in ['interp']:
    mr, cr = o.predict_truncation(f, Xn, order, ys[:,order], qf, rf, return_cov=True, kind=kind, excluded=[0])
    gsold = o.geometric_sum(qf(Xs), 0, order, [0]); gsnew = o.geometric_sum(qf(Xn), 0, order, [0])
    m_old = rf(Xs)*gsold*h.center; m_new = rf(Xn)*gsnew*h.center
    mean, cov, _ = h.predict(Xn, PREDICT_COV, Xc=Xs, yc=ys[:,order], mean_old=m_old, mean_new=m_new, sc_old=rf(Xs), sc_new=rf(Xn), q_old=qf(Xs), q_new=qf(Xn), gs_start=0, gs_end=order, excluded=[0], truncation=True)
    print('trunc interp mean', rel(mean[:,0], mr), 'cov', rel(cov, cr))
    mean, var, _ = h.predict(Xn, PREDICT_VAR, Xc=Xs, yc=ys[:,order], mean_old=m_old, mean_new=m_new, sc_old=rf(Xs), sc_new=rf(Xn), q_old=qf(Xs), q_new=qf(Xn), gs_start=0, gs_end=order, excluded=[0], truncation=True)
    print('trunc interp var', rel(var, np.diag(cr)))
Kt = o.truncation_cov(f, Xn, None, qf, rf, start=order+1, end=np.inf, excluded=[0])
K = ops.process_cov(Xn, None, 0.2, 1.0, 1e-6, factor=h.cov_factor, sc1=rf(Xn), q1=qf(Xn), gs_start=order+1, gs_end=np.inf, excluded=[0])
print('process_cov sym', rel(K, Kt))
Kt = o.truncation_cov(f, Xn, Xs, qf, rf, start=0, end=order, excluded=[0])
K = ops.process_cov(Xn, Xs, 0.2, 1.0, 1e-6, factor=h.cov_factor, sc1=rf(Xn), sc2=rf(Xs), q1=qf(Xn), q2=qf(Xs), gs_start=0, gs_end=order, excluded=[0])
print('process_cov cross', rel(K, Kt))
# diagnostics
for Nd in [200, 700]:
    Xd = np.sort(rs.rand(Nd))[:,None]; amp = 1.0 + 0.5*rs.rand(Nd)
    cov = 1.3*np.outer(amp,amp)*(RBF(0.2)(Xd)+1e-5*np.eye(Nd)); mean = 0.2+0*Xd[:,0]
    ch = np.linalg.cholesky(cov); Z = rs.randn(Nd, 40); Y = o.draws_from_z(mean, ch, Z)
    E, md2 = ops.cholesky_errors(ch, mean, Y, True, True)
    print('chol errors', Nd, rel(E, o.cholesky_errors(Y.T, mean, ch).T), 'md2', rel(md2, o.md_squared(Y, mean, ch)))
    t0=time.time(); G, Lp, piv, rank, rc = ops.pivoted_cholesky(cov); t1=time.time()
    Gr, pr = o.pivoted_cholesky(cov, True)
    print('pchol', Nd, 'rc', rc, 'rank', rank, 'piv equal', np.array_equal(piv, pr), (piv!=pr).sum(), 'G', rel(G, Gr), 'recon', rel(G@G.T, cov), 'ms', (t1-t0)*1e3)
    Ep = ops.pc_errors(Lp, piv, mean, Y)
    print('pc errors', rel(Ep, o.pivoted_cholesky_errors(Y, mean, Gr)), 'sum sq vs md2', rel((Ep**2).sum(0), md2))
    D, _ = ops.draws(ch, mean, Z=Z); print('draws', rel(D, Y))
    iv = np.linspace(0,1,21)
    import scipy.stats as st
    sd = np.sqrt(np.diag(cov)); lower, upper = st.norm(loc=mean, scale=sd).interval(np.atleast_2d(iv).T)
    ci = ops.credible_interval(Y, lower, upper); cr = o.credible_interval(Y, mean, cov, iv)
    print('coverage equal', np.array_equal(ci, cr), np.abs(ci-cr).max())
    D2, cov2 = ops.draws(ch, mean, n_draws=2000, seed=7, lower=lower, upper=upper)
    print('philox draws: mean err', np.abs(D2.mean(1)-mean).max(), 'cov err', np.abs(np.cov(D2)-cov).max(), 'coverage vs alpha max dev', np.abs(cov2.mean(0)-iv).max(), 'fused==separate', np.array_equal(cov2, ops.credible_interval(D2, lower, upper)))
# low-rank pchol
A = rs.randn(100, 30); M = A@A.T
G, Lp, piv, rank, rc = ops.pivoted_cholesky(M)
from scipy.linalg.lapack import dpstrf
c,p,r,i = dpstrf(M, lower=True)
print('lowrank rc', rc, 'rank', rank, 'lapack rank', r, 'info', i, 'piv[:rank] equal', np.array_equal(piv[:rank], (p-1)[:r]))
post, lse = ops.grid_normalize(np.array([[-1000.0, -1001.0],[-999.0,-1005.0]]))
from scipy.special import logsumexp
print('normalize', post, lse, logsumexp([-1000,-1001,-999,-1005]))
# timing: N=4096 diagnostics
Nd = 4096; Xd = np.linspace(0,1,Nd)[:,None]; cov = 1.3*(RBF(0.2)(Xd)+1e-5*np.eye(Nd)); mean = np.zeros(Nd)
t0=time.time(); ch = ops.cholesky(cov); t1=time.time(); print('chol 4096 ms', (t1-t0)*1e3, rel(ch, np.linalg.cholesky(cov)))
t0=time.time(); G, Lp, piv, rank, rc = ops.pivoted_cholesky(cov); t1=time.time(); print('pchol 4096 ms', (t1-t0)*1e3, rc, rank, rel(G@G.T, cov))
sd = np.sqrt(np.diag(cov)); iv = np.linspace(0,1,101); lower, upper = st.norm(loc=mean, scale=sd).interval(np.atleast_2d(iv).T)
for nd in [10000, 100000]:
    t0=time.time(); _, cv = ops.draws(ch, mean, n_draws=nd, seed=1, lower=lower, upper=upper, want_draws=False); t1=time.time()
    print('draws+coverage', nd, 'ms', (t1-t0)*1e3, 'max dev', np.abs(cv.mean(0)-iv).max())
