"""Golden vectors for the Student-t `Diagnostic` and for `Diagnostic.kl`, from the REAL reference
(gsum/diagnostics.py:38-68, 116-171).  statsmodels is not installed in this image, so its `MVT` is replaced by a stand-in
restating the published algorithm of `statsmodels.sandbox.distributions.multivariate.multivariate_t_rvs`
(`m + z / sqrt(chi2_df / df)`, z ~ N(0, sigma)) — here with an explicit RandomState and a Cholesky factor instead of the
global numpy state, and recording its (z, x) so the draws are reproducible from the saved arrays.  Everything the reference
computes itself (the t interval end points, coverage, error vectors, kl) comes from the reference's own code.
Run by hand in the build container; writes c5_student_diag.npz next to this file."""
import os
import sys
import warnings

import numpy as np
from sklearn.gaussian_process.kernels import RBF

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_loader import load_reference  # noqa: E402

helpers, models, datasets, diagnostics = load_reference()
warnings.filterwarnings("ignore")


class MVT:
    def __init__(self, mean, sigma, df):
        self.mean, self.sigma, self.df = np.asarray(mean), np.asarray(sigma), df
        self.random_state = None

    def rvs(self, size=1):
        rs = np.random.RandomState(self.random_state)
        d = len(self.mean)
        self.z = rs.standard_normal((d, size))
        self.x = rs.chisquare(self.df, size) / self.df
        zz = (np.linalg.cholesky(self.sigma) @ self.z).T
        return self.mean + zz / np.sqrt(self.x)[:, None]


def main():
    diagnostics.MVT = MVT
    rs = np.random.RandomState(5)
    N, df = 200, 7.0
    Xd = np.sort(rs.rand(N))[:, None]
    amp = 1.0 + 0.5 * rs.rand(N)
    cov = 1.3 * np.outer(amp, amp) * (RBF(0.2)(Xd) + 1e-5 * np.eye(N))
    mean = 0.2 + 0.1 * Xd[:, 0]
    d = diagnostics.Diagnostic(mean, cov, df=df, random_state=3)
    Y = d.samples(16)
    iv = np.linspace(0, 1, 51)
    lower, upper = d.udist.interval(np.atleast_2d(iv).T)
    # kl of the Gaussian diagnostic against a second (mean, cov)
    g = diagnostics.Diagnostic(mean, cov, random_state=1)
    cov0 = 0.9 * np.outer(amp, amp) * (RBF(0.25)(Xd) + 2e-5 * np.eye(N))
    mean0 = 0.25 + 0.05 * Xd[:, 0]
    out = dict(Xd=Xd, amp=amp, mean=mean, df=np.array(df), z=d.dist.z, x=d.dist.x, Y=Y, intervals=iv, lower=lower, upper=upper,
               coverage=d.credible_interval(Y, iv), md2=d.md_squared(Y), chol_errors=d.cholesky_errors(Y),
               pc_errors=d.pivoted_cholesky_errors(Y), ind_errors=d.individual_errors(Y),
               mean0=mean0, kl=np.array(g.kl(mean0, cov0)), kl_self=np.array(g.kl(mean, cov)))
    # cov / cov0 are rebuilt by the tests from (Xd, amp) with the same sklearn expression (bit-identical, keeps the file small)
    path = os.path.join(HERE, "c5_student_diag.npz")
    np.savez_compressed(path, **out)
    print(f"c5_student_diag: {os.path.getsize(path) / 1024:.1f} KiB; kl={out['kl']}, kl_self={out['kl_self']}")


if __name__ == "__main__":
    main()
