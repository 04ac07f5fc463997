"""ctypes binding of the C ABI in ``include/gsum_b200.h`` (``libgsum_b200.so``, built in-tree by
``__graft_entry__.build()`` / ``python -m gsum_b200.build``).

There is no CPU path: if the shared library is missing, or no CUDA device is visible, the first call
raises.  numpy arrays go in and out as host pointers (``GSUM_MEM_HOST``); torch CUDA tensors as device
pointers (``GSUM_MEM_DEVICE``) on the context's stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSUM_B200_LIB") or os.path.join(_HERE, "libgsum_b200.so")    # override: A/B builds during development

MEM_HOST, MEM_DEVICE = 0, 1
PREDICT_MEAN, PREDICT_VAR, PREDICT_COV = 0, 1, 2

_c_double_p = C.POINTER(C.c_double)
_c_int32_p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in the order of include/gsum_b200.h
_SIGNATURES = {
    "gsum_version": (C.c_int, []),
    "gsum_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "gsum_ctx_destroy": (C.c_int, [_vp]),
    "gsum_ctx_synchronize": (C.c_int, [_vp]),
    "gsum_last_error": (C.c_char_p, [_vp]),
    "gsum_launch_count": (C.c_int64, [_vp]),
    "gsum_device_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "gsum_device_free": (C.c_int, [_vp, _vp]),
    "gsum_device_copy": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int32]),
    "gsum_ctx_profile": (C.c_int, [_vp, C.c_int]),
    "gsum_ctx_profile_read": (C.c_int, [_vp, _vp, _vp, _vp]),
    "gsum_kernel_matrix": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, C.c_double,
                                     C.c_double, _vp, C.c_int32]),
    "gsum_cholesky": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, _vp, C.c_int32]),
    "gsum_cho_solve": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int32, C.c_int32]),
    "gsum_lml_grid": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, C.c_int64, C.c_int32,
                                _vp, C.c_int64, C.c_int32, _vp, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_int32, _vp, _vp, _vp, C.c_int32]),
    "gsum_grid_normalize": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int32]),
    "gsum_comm_unique_id": (C.c_int, [_vp, _vp]),
    "gsum_comm_init": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "gsum_comm_destroy": (C.c_int, [_vp]),
    "gsum_grid_allgather": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int32]),
    "gsum_comm_allreduce_counts": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32]),
    "gsum_lml_grad_terms": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_double, C.c_double,
                                      C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int32]),
    "gsum_lml_grad_terms_eig": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_double, C.c_double,
                                          C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int32]),
    "gsum_fit_create": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_double,
                                  C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                  _vp, _vp, C.c_int32, C.POINTER(_vp)]),
    "gsum_fit_destroy": (C.c_int, [_vp]),
    "gsum_predict": (C.c_int, [_vp, _vp, _vp, C.c_int32]),
    "gsum_process_cov": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_double, C.c_double, _vp, C.c_int64, _vp, C.c_int64,
                                   _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, C.c_int32, C.c_double, C.c_double, _vp, C.c_int32]),
    "gsum_cholesky_errors": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, C.c_int32]),
    "gsum_quadratic_forms": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, C.c_int32]),
    "gsum_pointwise_fit": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, C.c_int32, _vp, _vp, C.c_double, C.c_double,
                                     _vp, _vp, _vp, C.c_int32]),
    "gsum_pointwise_loglike": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int64,
                                         C.c_double, C.c_double, _vp, _vp, C.c_int32]),
    "gsum_variogram_bins": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, C.c_int32]),
    "gsum_variogram_cov": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int32, C.c_int32, _vp,
                                     C.c_double, C.c_double, C.c_int32, _vp, C.c_int32]),
    "gsum_pivoted_cholesky": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int32]),
    "gsum_pc_errors": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, C.c_int32]),
    "gsum_draws": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_uint64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int32,
                             _vp, _vp, C.c_int32]),
    "gsum_credible_interval": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, _vp, C.c_int32, _vp, C.c_int32]),
    "gsum_eigh": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int32]),
    "gsum_eig_conditional": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp, _vp, C.c_int32]),
    "gsum_eig_solve": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int32, C.c_int32]),
}

_lib = None
_lock = threading.Lock()


class GsumError(RuntimeError):
    """Invalid argument / CUDA failure reported by the library (reference: ValueError)."""


def load_library():
    """Load libgsum_b200.so and attach signatures.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  gsum_b200 has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def _ptr(a):
    """Pointer + keep-alive for a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


def as_f64(a, copy=False):
    a = np.array(a, dtype=np.float64, order="C", copy=True) if copy else np.ascontiguousarray(a, dtype=np.float64)
    return a


class Context:
    """One device + one stream (``gsum_ctx``).  Not thread-safe; create one per thread / rank."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.gsum_ctx_create(int(device), _vp(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise GsumError(f"gsum_ctx_create(device={device}) failed with {rc}: no usable CUDA device "
                            "(gsum_b200 has no CPU fallback)")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gsum_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        """Negative -> GsumError; positive is returned to the caller (numerical status)."""
        if rc < 0:
            msg = self.lib.gsum_last_error(self.handle)
            raise GsumError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    def synchronize(self):
        self.check(self.lib.gsum_ctx_synchronize(self.handle), "gsum_ctx_synchronize")

    def profile(self, enable=True):
        self.check(self.lib.gsum_ctx_profile(self.handle, 1 if enable else 0), "gsum_ctx_profile")

    def profile_read(self):
        """(device ms, algorithmic flops, brackets) of the factorisations since the last read."""
        ms, fl, nb = C.c_double(0), C.c_double(0), C.c_int64(0)
        self.check(self.lib.gsum_ctx_profile_read(self.handle, C.addressof(ms), C.addressof(fl), C.addressof(nb)), "gsum_ctx_profile_read")
        return ms.value, fl.value, nb.value

    @property
    def launch_count(self):
        return int(self.lib.gsum_launch_count(self.handle))


class DeviceBuffer:
    """A caller-held buffer in HBM (``gsum_device_malloc``): factors that stay resident between calls.

    `shape` / `dtype` describe the array it holds; `.get()` copies it back to a new numpy array."""

    def __init__(self, ctx, shape, dtype=np.float64):
        self.ctx, self.shape, self.dtype = ctx, tuple(int(s) for s in shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _vp()
        ctx.check(ctx.lib.gsum_device_malloc(ctx.handle, self.nbytes, C.byref(p)), "gsum_device_malloc")
        self.ptr = p.value

    def put(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.shape != self.shape:
            raise ValueError(f"expected shape {self.shape}, got {a.shape}")
        self.ctx.check(self.ctx.lib.gsum_device_copy(self.ctx.handle, self.ptr, a.ctypes.data, self.nbytes, 0), "gsum_device_copy")
        return self

    def copy_from(self, other):
        self.ctx.check(self.ctx.lib.gsum_device_copy(self.ctx.handle, self.ptr, other.ptr, self.nbytes, 2), "gsum_device_copy")
        return self

    def get(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.ctx.check(self.ctx.lib.gsum_device_copy(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes, 1), "gsum_device_copy")
        return out

    def free(self):
        if getattr(self, "ptr", None) and self.ctx.handle is not None:
            self.ctx.lib.gsum_device_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Process-wide context per device (device defaults to $GSUM_B200_DEVICE, $LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("GSUM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _default_ctx.get(device)
    if ctx is None or ctx.handle is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
