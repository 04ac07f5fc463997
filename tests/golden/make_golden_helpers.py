"""Golden vectors for the free helpers either side of the path, from the REAL reference (gsum/helpers.py:202-368:
`stabilize`, `predictions`, `gaussian`, `rbf`, `hpd`, `hpd_pdf`, `median_pdf`, `kl_gauss`) and the covariance / sample
statistics of `make_gaussian_partial_sums` (gsum/datasets.py:8-72), imported by path with the import-only stubs of
_reference_loader.py.

    python tests/golden/make_golden_helpers.py        (build container only; writes helpers_datasets.npz)
"""
import os
import sys
import warnings

import numpy as np
import scipy.stats as st
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")
out = {}
rs = np.random.RandomState(11)

# ---- correlation functions -----------------------------------------------------------------------------------------------
out["corr_X1"] = X1 = np.sort(rs.rand(37, 1), axis=0) * 2.0
out["corr_X2"] = X2 = rs.rand(23, 3)
out["corr_Xp2"] = Xp2 = rs.rand(11, 3)
out["corr_ls"] = ls_vals = np.array([0.07, 0.4, 2.5])
for i, ls in enumerate(ls_vals):
    out[f"rbf_1d_{i}"], out[f"gauss_1d_{i}"] = helpers.rbf(X1, ls=ls), helpers.gaussian(X1, ls=ls)
    out[f"rbf_3d_{i}"], out[f"gauss_3d_{i}"] = helpers.rbf(X2, ls=ls), helpers.gaussian(X2, ls=ls)
    out[f"rbf_3d_cross_{i}"], out[f"gauss_3d_cross_{i}"] = helpers.rbf(X2, Xp2, ls=ls), helpers.gaussian(X2, Xp2, ls=ls)
Xdup = np.array([[0.0], [0.5], [0.5], [1.0]])
out["rbf_ls0_X"], out["rbf_ls0"] = Xdup, helpers.rbf(Xdup, ls=0)
out["stabilize_in"] = M = rs.randn(5, 5)
out["stabilize_out"] = helpers.stabilize(M)

# ---- kl_gauss -------------------------------------------------------------------------------------------------------------
n = 60
x = np.linspace(0, 1, n)[:, None]
cov1 = 1.7 * (RBF(0.3)(x) + 1e-2 * np.eye(n))
A = rs.randn(n, n) * 0.05
cov0 = 0.8 * (RBF(0.2)(x) + 1e-3 * np.eye(n)) + A @ A.T
mu0, mu1 = 0.3 * np.sin(3 * x[:, 0]), 0.1 * np.ones(n)
out["kl_mu0"], out["kl_cov0"], out["kl_mu1"], out["kl_cov1"] = mu0, cov0, mu1, cov1
out["kl_from_cov"] = np.array(helpers.kl_gauss(mu0, cov0, mu1, cov1=cov1))
chol1 = np.linalg.cholesky(cov1)
out["kl_chol1"] = chol1
out["kl_from_chol"] = np.array(helpers.kl_gauss(mu0, cov0, mu1, chol1=chol1))
out["kl_scalar"] = np.array(helpers.kl_gauss(0.2, 1.3, -0.4, cov1=0.9))
out["kl_scalar_mean1"] = np.array(helpers.kl_gauss(np.zeros(n), cov0, 0.25, chol1=chol1))       # broadcast prior mean
print("kl", out["kl_from_cov"], out["kl_from_chol"], out["kl_scalar"], out["kl_scalar_mean1"])

# ---- pdf summaries --------------------------------------------------------------------------------------------------------
out["pdf_x"] = xg = np.linspace(-4.0, 6.0, 501)
out["pdf_vals"] = pdf = 0.6 * st.norm.pdf(xg, 0.3, 0.8) + 0.4 * st.norm.pdf(xg, 2.2, 1.1)
out["pdf_alphas"] = alphas = np.array([0.5, 0.68, 0.95])
out["hpd_pdf"] = np.array([helpers.hpd_pdf(pdf, a, xg) for a in alphas])
out["median_pdf"] = np.array(helpers.median_pdf(pdf, xg))
out["hpd_norm"] = np.array([helpers.hpd(st.norm(0.3, 1.1), a) for a in alphas])
out["hpd_gamma"] = np.array([helpers.hpd(st.gamma, a, 2.0) for a in alphas])
out["hpd_t"] = np.array([helpers.hpd(st.t(4.5, loc=-1.0, scale=0.6), a) for a in alphas])
dist = st.norm(np.linspace(-1, 1, 7), np.linspace(0.5, 2.0, 7))
out["pred_mean"] = helpers.predictions(dist)
m, iv = helpers.predictions(dist, dob=[0.68, 0.95])
out["pred_intervals"] = iv
_, iv1 = helpers.predictions(dist, dob=0.5)
out["pred_interval_single"] = iv1
print("hpd_pdf", out["hpd_pdf"].tolist(), "median", out["median_pdf"], "pred", iv.shape, iv1.shape)

# ---- make_gaussian_partial_sums: the covariance it draws from, and the sample statistics of the reference's own draws -------
nX = 30
Xd = np.linspace(0, 1, nX)[:, None]
kern = ConstantKernel(1.5) * RBF(0.25) + WhiteKernel(1e-3)
ratio_fn, ref_fn = (lambda X: 0.3 + 0.2 * X[:, 0]), (lambda X: 2.0 - X[:, 0])
mean_fn = lambda X: 0.5 * np.ones(X.shape[0])
orders = np.array([0, 2, 3, 5])
out["ds_X"], out["ds_orders"] = Xd, orders
K = kern(Xd) + 1e-4 * np.eye(nX)
out["ds_K"] = K
n_draw = 4000
big = datasets.make_gaussian_partial_sums(Xd, orders=n_draw, kernel=kern, mean=mean_fn, ratio=1.0, ref=1.0, nugget=1e-4, random_state=5)
coeffs = helpers.coefficients(big, 1.0, 1.0, np.arange(n_draw))                       # ratio = ref = 1: plain differences
out["ds_ref_sample_mean"], out["ds_ref_sample_cov"] = coeffs.mean(axis=1), np.cov(coeffs)
y = datasets.make_gaussian_partial_sums(Xd, orders=orders, kernel=kern, mean=mean_fn, ratio=ratio_fn, ref=ref_fn, nugget=1e-4, random_state=5)
out["ds_y_shape"] = np.array(y.shape)
Xu, yu = datasets.make_gaussian_partial_sums_uniform(n_samples=12, n_features=2, orders=3, random_state=9)
out["ds_uniform_X"], out["ds_uniform_y_shape"] = Xu, np.array(yu.shape)
Xg, yg = datasets.make_gaussian_partial_sums_on_grid(n_samples=9, n_features=1, orders=4, random_state=9)
out["ds_grid_X"], out["ds_grid_y_shape"] = Xg, np.array(yg.shape)
print("datasets", y.shape, yu.shape, yg.shape, "sample cov err", np.max(np.abs(out["ds_ref_sample_cov"] - K)))

np.savez_compressed(os.path.join(HERE, "helpers_datasets.npz"), **out)
print("wrote helpers_datasets.npz:", len(out), "arrays")
