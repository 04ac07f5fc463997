"""Golden vectors for the analytic likelihood gradient, from the REAL reference:
`ConjugateGaussianProcess.log_marginal_likelihood(theta, eval_gradient=True)` (gsum/models.py:912-1057) with all of the
kernel's hyperparameters free.  Run by hand in the build container; writes c1_gradient.npz next to this file."""
import os
import sys
import warnings

import numpy as np
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402
from make_golden import PRIORS, prior_array  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")


def main():
    from scipy import stats
    N = 50
    X = np.linspace(0, 1, N)[:, None]
    K = RBF(0.2)(X) + 1e-6 * np.eye(N)
    y = stats.multivariate_normal(np.zeros(N), K, allow_singular=True).rvs(5, random_state=0).T
    # theta = log [constant, length scale, noise level]  (sklearn order for C * RBF + White)
    thetas = np.log(np.array([[1.5, 0.2, 1e-4], [0.7, 0.1, 1e-3], [2.5, 0.35, 1e-5], [1.0, 0.05, 1e-2]]))
    out = dict(X=X, y=y, thetas=thetas, priors=np.stack([prior_array(p) for p in PRIORS]), nugget=np.array(1e-10))
    for ip, p in enumerate(PRIORS):
        gp = models.ConjugateGaussianProcess(C(1.5) * RBF(0.2) + WhiteKernel(1e-4), nugget=1e-10, optimizer=None, **p).fit(X, y)
        res = [gp.log_marginal_likelihood(theta=t, eval_gradient=True) for t in thetas]
        out[f"g{ip}_lml"] = np.array([r[0] for r in res])
        out[f"g{ip}_grad"] = np.array([r[1] for r in res])
    # NB: the reference's Student-t gradient cannot be frozen: ConjugateStudentProcess.log_marginal_likelihood calls
    # `kernel(X, eval_gradient)` (models.py:1200), passing the flag as Y, and crashes in sklearn.  The oracle restates the
    # formulas of models.py:1260-1271 and is pinned by central differences instead (tests/test_oracle.py).
    # only the length scale free (the notebooks' usual setting), and a 2-D anisotropic kernel
    gp = models.ConjugateGaussianProcess(RBF(0.2) + WhiteKernel(1e-4, 'fixed'), nugget=1e-10, optimizer=None, **PRIORS[1]).fit(X, y)
    th1 = np.log(np.array([[0.05], [0.2], [0.4]]))
    res = [gp.log_marginal_likelihood(theta=t, eval_gradient=True) for t in th1]
    out.update(ls_thetas=th1, ls_lml=np.array([r[0] for r in res]), ls_grad=np.array([r[1] for r in res]))
    g = np.linspace(0, 1, 8)
    X2 = helpers.cartesian(g, g)
    K2 = RBF([0.3, 0.15])(X2) + 1e-6 * np.eye(len(X2))
    y2 = stats.multivariate_normal(np.zeros(len(X2)), K2, allow_singular=True).rvs(3, random_state=5).T
    gp2 = models.ConjugateGaussianProcess(C(1.2) * RBF([0.3, 0.15]) + WhiteKernel(1e-4), nugget=1e-10, optimizer=None,
                                          **PRIORS[2]).fit(X2, y2)
    th2 = np.log(np.array([[1.2, 0.3, 0.15, 1e-4], [0.8, 0.2, 0.25, 1e-3]]))
    res = [gp2.log_marginal_likelihood(theta=t, eval_gradient=True) for t in th2]
    out.update(X2=X2, y2=y2, aniso_thetas=th2, aniso_lml=np.array([r[0] for r in res]), aniso_grad=np.array([r[1] for r in res]))
    path = os.path.join(HERE, "c1_gradient.npz")
    np.savez_compressed(path, **out)
    print("c1_gradient:", os.path.getsize(path), "bytes")
    print(out["g1_lml"], out["g1_grad"], out["aniso_grad"], sep="\n")


if __name__ == "__main__":
    main()
