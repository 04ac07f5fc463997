"""CPU: the Python host of the facade on a fake device — the SAME test bodies as the `-m gpu` parity tests
(tests/test_gpu_lml.py, tests/test_gpu_predict.py, tests/test_gpu_helpers.py), with `gsum_b200.ops` replaced by the numpy /
LAPACK stand-ins of tests/fake_device.py.  What this pins in the CPU suite: argument marshalling, coefficient / order /
excluded-order bookkeeping, the truncation scalings, the prior branches, the Student-t additions, caches and error mapping of
`models.py` against golden vectors of the real reference.  What it cannot pin — the kernels — is the `-m gpu` suite's job."""
import pytest

import fake_device
import test_gpu_diagnostics as D
import test_gpu_gradient as G
import test_gpu_helpers as H
import test_gpu_lml as L
import test_gpu_predict as P


@pytest.fixture
def ctx(monkeypatch):
    """Overrides the device-context fixture of conftest.py: no library, no device — the stand-ins instead."""
    fake_device.install(monkeypatch)
    return None


# the likelihood grid
test_kat_notebook_grid = L.test_kat_notebook_grid
test_c2_grid_all_priors = L.test_c2_grid_all_priors
test_c2_grid_x_dependent_ratio = L.test_c2_grid_x_dependent_ratio
test_c1_conjugate_lml = L.test_c1_conjugate_lml
test_grid_matches_oracle_random_inputs_2d = L.test_grid_matches_oracle_random_inputs_2d
test_edge_cases = L.test_edge_cases
# fit / predict / truncation predict (a6-a11)
test_c1_fit_posteriors_and_predict = P.test_c1_fit_posteriors_and_predict
test_interpolation_property = P.test_interpolation_property
test_interpolation_additive_constant_kernel_raises = P.test_interpolation_additive_constant_kernel_raises
test_prior_predict_and_cov_before_fit = P.test_prior_predict_and_cov_before_fit
test_c3_truncation_predict = P.test_c3_truncation_predict
test_c3_constrained_truncation_error = P.test_c3_constrained_truncation_error
test_sample_y_statistics = P.test_sample_y_statistics
# helpers and generators either side of the path
test_correlation_functions_against_reference_golden = H.test_correlation_functions_against_reference_golden
test_kl_gauss = H.test_kl_gauss
test_partial_sums_equal_the_factor_applied_to_the_same_normals = H.test_partial_sums_equal_the_factor_applied_to_the_same_normals
test_partial_sums_have_the_covariance_of_the_kernel = H.test_partial_sums_have_the_covariance_of_the_kernel
test_legacy_generators = H.test_legacy_generators
# analytic gradient and the L-BFGS fit (f1)
test_c1_gradient_all_priors = G.test_c1_gradient_all_priors
test_gradient_single_free_length_scale_and_aniso = G.test_gradient_single_free_length_scale_and_aniso
test_grad_terms_against_numpy = G.test_grad_terms_against_numpy
test_fit_with_free_hyperparameters = G.test_fit_with_free_hyperparameters
test_student_gradient = G.test_student_gradient
test_fit_with_optimizer_recovers_length_scale = P.test_fit_with_optimizer_recovers_length_scale
test_interpolation_property_free_theta = P.test_interpolation_property_free_theta
# diagnostics (a12-a14, f3)
test_c5_golden = D.test_c5_golden
test_pivoted_cholesky_known_answers = D.test_pivoted_cholesky_known_answers
test_draws_with_supplied_z_and_helpers = D.test_draws_with_supplied_z_and_helpers
test_c5_student_diag_golden_and_kl = D.test_c5_student_diag_golden_and_kl


def test_mahalanobis_inv_and_sqrt_mat(ctx):
    D.test_mahalanobis_inv_and_sqrt_mat()
