"""Timing of the device Jacobi eigensolver and the eig-route solves next to LAPACK on the host.
Usage: python tools/eig_probe.py [N ...]"""
import sys
import time

import numpy as np
from sklearn.gaussian_process.kernels import RBF

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from gsum_b200 import ops  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 2048]
    ops.ResidentEigen(np.eye(max(sizes)))        # allocate the workspaces outside the timings
    for n in sizes:
        X = np.linspace(0, 1, n)[:, None]
        A = RBF(0.05)(X) + 1e-4 * np.eye(n)
        t0 = time.perf_counter()
        res = ops.ResidentEigen(A)
        t1 = time.perf_counter()
        wl = np.linalg.eigvalsh(A)
        t2 = time.perf_counter()
        wl2, Vl = np.linalg.eigh(A)
        t3 = time.perf_counter()
        Y = np.random.RandomState(0).randn(n, 512)
        res.solve(Y)
        t4 = time.perf_counter()
        Xs = res.solve(Y)
        t5 = time.perf_counter()
        rounds = res.sweeps * (n - 1 + (n & 1))
        stream = 4.0 * n * n * 8 * rounds / (t1 - t0) / 1e9
        print(f"N={n}: device eigh {1e3 * (t1 - t0):.1f} ms ({res.sweeps} sweeps, {rounds} rounds, <= {stream:.0f} GB/s streamed), "
              f"LAPACK eigh {1e3 * (t3 - t2):.1f} ms; max|w - w_lapack| / w_max = {np.max(np.abs(res.w - wl)) / wl[-1]:.2e}; "
              f"solve of 512 rhs {1e3 * (t5 - t4):.2f} ms ({4.0 * n * n * 512 / (t5 - t4) / 1e12:.2f} TFLOP/s incl. copies), "
              f"residual {np.max(np.abs(A @ Xs - Y)) / np.max(np.abs(Y)):.1e} "
              f"(LAPACK's Q diag(1/eig) Q^T y: {np.max(np.abs(A @ (Vl @ ((Vl.T @ Y) / wl2[:, None])) - Y)) / np.max(np.abs(Y)):.1e})", flush=True)


if __name__ == "__main__":
    main()
