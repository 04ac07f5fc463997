"""Flatten an sklearn kernel object into the descriptor the device kernels evaluate.

The reference takes arbitrary ``sklearn.gaussian_process.kernels`` objects (gsum/models.py:12,146-147,
599,708,953-960).  The device builder (csrc/cov.cuh) evaluates the family every gsum notebook and test
uses on this path::

    [ConstantKernel *] RBF(length_scale)  [* ConstantKernel]  [+ WhiteKernel(noise_level)]

i.e. ``c * exp(-0.5 |x/l - x'/l|^2) + noise * 1[x is x']``.  Anything else raises NotImplementedError —
there is no CPU fallback (BASELINE.json north_star).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Product, Sum, WhiteKernel

__all__ = ["KernelDesc", "flatten_kernel", "theta_layout"]


@dataclass
class KernelDesc:
    constant: float          # product of all ConstantKernel factors (1.0 if none)
    length_scale: np.ndarray  # (1,) isotropic or (d,) anisotropic
    noise: float             # WhiteKernel noise level (0.0 if none)

    def ls_for(self, d):
        ls = np.ascontiguousarray(self.length_scale, dtype=np.float64)
        if ls.shape[0] not in (1, d):
            raise ValueError(f"Anisotropic kernel must have the same number of dimensions as data ({ls.shape[0]}!={d})")
        return ls


def _flatten_product(k):
    """-> (constant, length_scale) for a product of ConstantKernels and exactly one RBF."""
    if type(k) is RBF:          # Matern / other subclasses are different kernels
        return 1.0, np.atleast_1d(np.asarray(k.length_scale, dtype=np.float64))
    if type(k) is ConstantKernel:
        return float(k.constant_value), None
    if isinstance(k, Product):
        c1, l1 = _flatten_product(k.k1)
        c2, l2 = _flatten_product(k.k2)
        if l1 is not None and l2 is not None:
            raise NotImplementedError("gsum_b200: a product of two RBF kernels is not supported on the device path")
        return c1 * c2, l1 if l1 is not None else l2
    raise NotImplementedError(f"gsum_b200: kernel {k!r} is not supported on the device path "
                              "(supported: [Constant *] RBF [+ WhiteKernel])")


def flatten_kernel(kernel) -> KernelDesc:
    noise = 0.0
    terms = []

    def split(k):
        nonlocal noise
        if isinstance(k, Sum):
            split(k.k1)
            split(k.k2)
        elif isinstance(k, WhiteKernel):
            noise += float(k.noise_level)
        else:
            terms.append(k)

    split(kernel)
    if len(terms) != 1:
        raise NotImplementedError(f"gsum_b200: kernel {kernel!r} is not supported on the device path "
                                  "(exactly one [Constant *] RBF term, plus optional WhiteKernel terms)")
    c, ls = _flatten_product(terms[0])
    if ls is None:
        raise NotImplementedError(f"gsum_b200: kernel {kernel!r} has no RBF factor")
    return KernelDesc(constant=c, length_scale=ls, noise=noise)


def theta_layout(kernel, d):
    """For every entry of ``kernel.theta`` (sklearn order: the tree is walked k1 before k2, fixed hyperparameters are
    skipped): ``(slot, weight)`` with slot 0 = constant, 1 + q = length scale q (q = 0 when isotropic), 1 + ls_dim =
    noise level — the order of the device's derivative matrices — and d/dtheta_i = weight * d/d(slot).  The weight is 1
    except for one of several WhiteKernel terms (its share of the total noise level)."""
    desc = flatten_kernel(kernel)
    ls_dim = desc.ls_for(d).shape[0]
    out = []

    def walk(k):
        if isinstance(k, (Sum, Product)):
            walk(k.k1)
            walk(k.k2)
            return
        for hp in k.hyperparameters:
            if hp.fixed:
                continue
            if type(k) is ConstantKernel:
                out.append((0, 1.0))
            elif type(k) is RBF:
                out.extend((1 + q, 1.0) for q in range(hp.n_elements))
            elif isinstance(k, WhiteKernel):
                out.append((1 + ls_dim, float(k.noise_level) / desc.noise if desc.noise > 0 else 0.0))
            else:
                raise NotImplementedError(f"gsum_b200: no analytic gradient for {k!r}")

    walk(kernel)
    return out
