"""gsum_b200 — B200-native conjugate-GP likelihood / prediction / diagnostics path of buqeye/gsum.

Drop-in for the reference's hot path: same class and function names (``ConjugateGaussianProcess``,
``ConjugateStudentProcess``, ``TruncationGP``, ``TruncationTP``, ``TruncationPointwise``, ``Diagnostic``, ``VariogramFourthRoot``, ``coefficients``,
``partials``, ``geometric_sum``, ``pivoted_cholesky``, ``cholesky_errors``, ``mahalanobis``,
``cartesian``); the arithmetic runs in hand-written sm_100a CUDA behind the C ABI of
``include/gsum_b200.h`` (``libgsum_b200.so``).  No CPU fallback: importing is cheap, the first numerical
call raises if the library is not built or no GPU is visible.
"""
from .helpers import (cartesian, cholesky_errors, coefficients, geometric_sum, mahalanobis, partials,
                      pivoted_cholesky)
from .models import (BaseConjugateProcess, ConjugateGaussianProcess, ConjugateStudentProcess, TruncationGP,
                     TruncationProcess, TruncationTP)
from .diagnostics import Diagnostic
from .pointwise import TruncationPointwise
from .variogram import VariogramFourthRoot

__version__ = "0.1.0"
__all__ = [
    "ConjugateGaussianProcess", "ConjugateStudentProcess", "TruncationGP", "TruncationTP", "TruncationProcess",
    "BaseConjugateProcess", "Diagnostic", "TruncationPointwise", "VariogramFourthRoot", "cartesian", "coefficients", "partials", "geometric_sum",
    "pivoted_cholesky", "cholesky_errors", "mahalanobis",
]
