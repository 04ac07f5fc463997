import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSUM_B200_SCHEDULE"] = os.environ.get("HT_MODE", "hetero")
from gsum_b200 import ops
from sklearn.gaussian_process.kernels import RBF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X = np.linspace(0, 1, n)[:, None]
A = np.stack([RBF(l)(X) + 1e-6 * np.eye(n) for l in np.geomspace(0.01, 0.3, nb)])
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    L, info, ld = ops.cholesky(A.copy(), return_info=True)
    print("info", info, "max diff", np.nanmax(np.abs(L - np.linalg.cholesky(A))))
