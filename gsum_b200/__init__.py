"""gsum_b200 — B200-native conjugate-GP likelihood / prediction / diagnostics path of buqeye/gsum.

Drop-in for the reference's hot path: same class and function names (``ConjugateGaussianProcess``,
``ConjugateStudentProcess``, ``TruncationGP``, ``TruncationTP``, ``TruncationPointwise``, ``Diagnostic``, ``VariogramFourthRoot``, ``coefficients``,
``partials``, ``geometric_sum``, ``pivoted_cholesky``, ``cholesky_errors``, ``mahalanobis``,
``cartesian``, ``kl_gauss``, ``rbf``, ``gaussian``, the ``make_gaussian_partial_sums*`` generators); the arithmetic runs in hand-written sm_100a CUDA behind the C ABI of
``include/gsum_b200.h`` (``libgsum_b200.so``).  No CPU fallback: importing is cheap, the first numerical
call raises if the library is not built or no GPU is visible.
"""
from .helpers import (cartesian, cholesky_errors, coefficients, default_attributes, gaussian, geometric_sum, hpd, hpd_pdf,
                      kl_gauss, lazy_property, mahalanobis, median_pdf, partials, pivoted_cholesky, predictions, rbf, stabilize)
from .datasets import (generate_coefficients, make_gaussian_partial_sums, make_gaussian_partial_sums_on_grid,
                       make_gaussian_partial_sums_uniform, toy_data)
from .models import (BaseConjugateProcess, ConjugateGaussianProcess, ConjugateStudentProcess, TruncationGP,
                     TruncationProcess, TruncationTP)
from .diagnostics import Diagnostic
from .pointwise import TruncationPointwise
from .variogram import VariogramFourthRoot

__version__ = "0.1.0"
__all__ = [
    "ConjugateGaussianProcess", "ConjugateStudentProcess", "TruncationGP", "TruncationTP", "TruncationProcess",
    "BaseConjugateProcess", "Diagnostic", "TruncationPointwise", "VariogramFourthRoot", "cartesian", "coefficients", "partials", "geometric_sum",
    "pivoted_cholesky", "cholesky_errors", "mahalanobis", "stabilize", "rbf", "gaussian", "kl_gauss", "predictions", "hpd", "hpd_pdf",
    "median_pdf", "make_gaussian_partial_sums", "make_gaussian_partial_sums_uniform", "make_gaussian_partial_sums_on_grid",
    "toy_data", "generate_coefficients", "lazy_property", "default_attributes",
]
