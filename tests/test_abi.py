"""CPU: the C-ABI shared library loads, exports every symbol include/gsum_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from gsum_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gsum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsum_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ("gsum_ctx_create", "gsum_lml_grid", "gsum_cholesky", "gsum_cho_solve", "gsum_fit_create", "gsum_predict",
              "gsum_pivoted_cholesky", "gsum_pc_errors", "gsum_cholesky_errors", "gsum_draws", "gsum_credible_interval",
              "gsum_grid_normalize", "gsum_kernel_matrix", "gsum_process_cov"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/gsum_b200.h but not exported by libgsum_b200.so"


def test_python_binding_covers_header():
    assert sorted(_lib.exported_symbols()) == header_symbols()
    assert _lib.load_library().gsum_version() == 100


def test_no_cpu_fallback_without_device():
    """On a machine without CUDA the context cannot be created and every numerical entry point is unreachable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the no-device behaviour is exercised on the CPU builder only")
    with pytest.raises(_lib.GsumError, match="no CPU fallback"):
        _lib.Context(0)
    from gsum_b200 import ConjugateGaussianProcess
    from sklearn.gaussian_process.kernels import RBF
    X = np.linspace(0, 1, 8)[:, None]
    with pytest.raises(_lib.GsumError):
        ConjugateGaussianProcess(RBF(0.3, 'fixed')).fit(X, np.sin(X[:, 0]))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gsum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("gsum_oracle_free", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/gsum_b200.h compiles as C (gcc -std=c99 -pedantic), a C program links against libgsum_b200.so and runs: without a
    CUDA device `gsum_ctx_create` refuses (-2), with one the program evaluates a 3 x 3 kernel matrix through the C ABI."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not on PATH")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "native", "abi_smoke.c"), "-L", libdir, "-lgsum_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "version 100" in r.stdout, (r.stdout, r.stderr)
