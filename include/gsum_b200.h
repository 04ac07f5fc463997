/* gsum_b200 — C ABI of the B200-native conjugate-GP likelihood / prediction / diagnostics path.
 *
 * The reference (buqeye/gsum) is pure Python and has no FFI of its own: its hot path calls
 * numpy.linalg / scipy.linalg / LAPACK / scikit-learn kernels directly.  Each entry point below replaces
 * one family of those call sites (cited per function, paths relative to the reference checkout) and is
 * what a ctypes binding inside gsum/models.py, gsum/helpers.py and gsum/diagnostics.py would bind (see
 * INTEGRATION.md for the stubs).
 *
 * Conventions
 *   - C linkage, plain C types only.  All matrices are FP64, row-major (numpy C order), contiguous.
 *   - The caller owns every buffer passed in or out.  `mem_kind` says where they live:
 *       GSUM_MEM_HOST    host pointers (numpy); the call copies in/out and returns after the result landed.
 *       GSUM_MEM_DEVICE  device pointers on the context's GPU (torch .data_ptr()); the call only enqueues
 *                        work on the context's stream and returns without synchronising.
 *   - Return value: 0 ok; < 0 invalid argument / CUDA failure (reference: ValueError); > 0 numerical
 *     failure (reference: numpy.linalg.LinAlgError), e.g. the 1-based index of the first non-positive
 *     pivot.  gsum_last_error() returns a message for the last non-zero return on that context.
 *   - A context is bound to one device and one stream and is not thread-safe; distinct contexts are
 *     independent.  Nothing here falls back to the CPU: without a CUDA device every call fails.
 */
#ifndef GSUM_B200_H
#define GSUM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSUM_MEM_HOST 0
#define GSUM_MEM_DEVICE 1
/* OR-ed into GSUM_MEM_HOST for the calls that take a factor (gsum_cho_solve, gsum_cholesky_errors, gsum_pc_errors,
 * gsum_draws): the factor argument (L / Lp and piv) is a DEVICE pointer — a factor the caller keeps resident in HBM
 * (gsum_device_malloc) — while every other buffer is a host pointer.  Saves re-uploading an N x N factor per call. */
#define GSUM_MEM_FACTOR_DEVICE 2

#define GSUM_PREDICT_MEAN 0
#define GSUM_PREDICT_VAR 1
#define GSUM_PREDICT_COV 2

typedef struct gsum_ctx gsum_ctx;
typedef struct gsum_fit gsum_fit;

/* ABI version (major*100 + minor). */
int gsum_version(void);

/* Context: device + stream (+ ctx-scoped workspaces).  `cuda_stream` may be NULL (a private stream is created). */
int gsum_ctx_create(int device, void *cuda_stream, gsum_ctx **out);
int gsum_ctx_destroy(gsum_ctx *ctx);
int gsum_ctx_synchronize(gsum_ctx *ctx);
const char *gsum_last_error(const gsum_ctx *ctx);
/* Number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
int64_t gsum_launch_count(const gsum_ctx *ctx);

/* Caller-held device buffers on the context's GPU — for factors that stay in HBM between calls (GSUM_MEM_DEVICE /
 * GSUM_MEM_FACTOR_DEVICE arguments) when the host side has no device allocator of its own.  gsum_device_copy is ordered
 * on the context's stream and returns when the copy is done; direction 0 = host->device, 1 = device->host, 2 = device->device.
 * The buffers come from the device's stream-ordered pool (cudaMallocAsync on the context's stream; the pool keeps what it is
 * given): use them with this context only.  gsum_device_free is stream-ordered as well: work already enqueued on the buffer
 * completes before the memory is reused, and the call does not synchronise. */
int gsum_device_malloc(gsum_ctx *ctx, size_t bytes, void **out);
int gsum_device_free(gsum_ctx *ctx, void *ptr);
int gsum_device_copy(gsum_ctx *ctx, void *dst, const void *src, size_t bytes, int32_t direction);

/* Optional profiling of the factorisation phase (K2+K3: every launch of the bordered Cholesky), used by bench.py for
 * the roofline: CUDA events are recorded on the context's stream around each factorisation while enabled.
 * gsum_ctx_profile_read synchronises, returns the summed device time (ms), the algorithmic flops of the bracketed work
 * (batch * (N^3/3 + N^2 * rhs)) and the number of brackets, and resets the counters. */
int gsum_ctx_profile(gsum_ctx *ctx, int enable);
int gsum_ctx_profile_read(gsum_ctx *ctx, double *ms_total, double *flops_total, int64_t *n_brackets);

/* K1 — kernel matrix  c * RBF_ls(X1, X2) [+ noise on the diagonal when X2 == NULL].
 * Replaces sklearn kernel __call__ at gsum/models.py:599,708,807,822-824,960 (RBF, ConstantKernel*RBF,
 * ... + WhiteKernel).  X1 (n1,d), X2 (n2,d) or NULL (symmetric k(X1), diagonal = c + noise exactly);
 * ls (ls_dim,) with ls_dim in {1, d}; out (n1, n2 or n1). */
int gsum_kernel_matrix(gsum_ctx *ctx, const double *X1, int64_t n1, const double *X2, int64_t n2, int32_t d,
                       const double *ls, int32_t ls_dim, double constant, double noise, double *out,
                       int32_t mem_kind);

/* K2 — batched lower Cholesky, in place: A (batch, n, n) symmetric in -> L in the lower triangle, zeros above.
 * Replaces numpy.linalg.cholesky at gsum/models.py:711,809,969,1211 and gsum/diagnostics.py:60.
 * info (batch,) int32: 0 or the 1-based column of the first non-positive pivot (LAPACK potrf); logdet (batch,)
 * = 2 * sum(log(diag L)) or NULL.  Returns 0 even if some matrices failed (inspect info). */
int gsum_cholesky(gsum_ctx *ctx, double *A, int64_t n, int64_t batch, int32_t *info, double *logdet,
                  int32_t mem_kind);

/* K3 — triangular solves with the lower factor L (n,n), B (n, nrhs) overwritten:
 *   forward_only != 0 :  B <- L^{-1} B        (scipy.linalg.solve_triangular(lower=True), gsum/helpers.py:505)
 *   forward_only == 0 :  B <- L^{-T} L^{-1} B (scipy.linalg.cho_solve, gsum/models.py:479) */
int gsum_cho_solve(gsum_ctx *ctx, const double *L, int64_t n, double *B, int64_t nrhs, int32_t forward_only,
                   int32_t mem_kind);

/* K1-K4 fused — the (Q, l) log-marginal-likelihood grid.
 * Replaces the nested loop of docs/notebooks/correlated_EFT_publication.ipynb cell 53 over
 * TruncationProcess.log_marginal_likelihood (gsum/models.py:1485-1507) ->
 * ConjugateGaussianProcess / ConjugateStudentProcess.log_marginal_likelihood (models.py:912-1057 / 1184-1273)
 * -> coefficients (helpers.py:71-101), compute_center/disp/df/scale_sq/cov_factor (models.py:169-503).
 *
 *   X (n,d); dy (n,n_c) = [y_0, diff(y)] of the retained orders (NOT yet divided by ref / Q**order);
 *   ref (n,); orders (n_c,) the integer powers; ls (n_ls, ls_dim);
 *   q_x_dependent == 0: Q (n_q,) scalar ratios — each length scale is factored once and its Gram is reused
 *                       for every Q;  != 0: Q (n_q, n) ratio evaluated at every x (one RHS block per Q);
 *   detf (n_q,) or NULL: the Jacobian sum_x [n_c log|ref| + sum(orders) log|Q|] subtracted per Q (models.py:1503-1506);
 *   kernel = constant * RBF(ls) + noise * I, diagonal += nugget (models.py:963);
 *   priors center0/disp0/df0/scale0 (df0 may be INFINITY); student != 0 selects the Student-t evidence.
 * Outputs: ll (n_q, n_ls) — `[ratio][ls]` as in the notebook; logdet (n_ls,) or NULL; status (n_ls,) or NULL
 * (non-zero where the Cholesky failed; those cells are -inf as in models.py:970-972). */
int gsum_lml_grid(gsum_ctx *ctx, const double *X, int64_t n, int32_t d, const double *dy, int32_t n_c,
                  const double *ref, const int32_t *orders, const double *ls, int64_t n_ls, int32_t ls_dim,
                  const double *Q, int64_t n_q, int32_t q_x_dependent, const double *detf, double constant,
                  double noise, double nugget, double center0, double disp0, double df0, double scale0,
                  int32_t student, double *ll, double *logdet, int32_t *status, int32_t mem_kind);

/* Gradient terms of the conjugate-GP likelihood — the device half of log_marginal_likelihood(theta, eval_gradient=True)
 * (gsum/models.py:957-1056: `kernel(X, eval_gradient=True)`, `cho_solve(L, eye(N))` at 1045, the einsum contractions of
 * compute_center 227-229, compute_scale_sq 451-454 and 1047-1048).  With RHS = [basis | y_1..y_nc] (n, r),
 * Z = R^{-1} RHS and dR_p = dR/dlog(theta_p), p = [constant, length scale(s), noise level] (P = ls_dim + 2):
 *   G (r, r) = RHS^T Z,   H (P, r, r) = Z^T dR_p Z,   tr (P,) = trace(R^{-1} dR_p),   logdet = log det R,
 *   info = 0 or the first non-positive pivot (then the other outputs are undefined; the reference returns -inf, 0).
 * R = constant * RBF_ls(X) + (noise + nugget) I as in gsum_lml_grid. */
int gsum_lml_grad_terms(gsum_ctx *ctx, const double *X, int64_t n, int32_t d, const double *RHS, int32_t r,
                        const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                        double *G, double *H, double *tr, double *logdet, int32_t *info, int32_t mem_kind);

/* On-device normalisation of a gathered grid (docs/notebooks/correlated_EFT_publication.ipynb cell 54):
 * post = exp(ll - max(ll)); lse = log(sum(exp(ll))).  ll, post: (count,); lse: 1 double or NULL. */
int gsum_grid_normalize(gsum_ctx *ctx, const double *ll, int64_t count, double *post, double *lse, int32_t mem_kind);

/* ---- the collective of the sharded grid (gsum_comm_init / _destroy) --------------------------------------------
 * One process per GPU, one context per process.  Rank 0 obtains an id (GSUM_COMM_ID_BYTES bytes, an ncclUniqueId) and hands it
 * to the other ranks by the host's own means (MPI, a socket, torch.distributed); every rank then calls gsum_comm_init.  NCCL is
 * resolved with dlopen("libnccl.so.2") at the first of these calls (GSUM_B200_NCCL_LIB overrides the name): libgsum_b200.so
 * does not link against it.  Replaces the Python list comprehension over `ls_vals` of
 * docs/notebooks/correlated_EFT_publication.ipynb cell 53 when the length scales are dealt over several GPUs.
 *
 * gsum_grid_allgather: this rank evaluated gsum_lml_grid on the length scales  rank, rank + nranks, rank + 2 nranks, ...  into
 * `block` (n_q, per), per = ceil(n_ls / nranks) (columns beyond the rank's share are ignored).  ONE ncclAllGather on the
 * context's stream, the deal undone on the device into `ll_full` (n_q, n_ls), and — if `post` or `lse` is given — the
 * normalisation of gsum_grid_normalize on the gathered grid (notebook cell 54), all without leaving the device.
 * gsum_comm_allreduce_counts: in-place sum over the ranks of int64 counts (coverage counts of sharded posterior draws). */
#define GSUM_COMM_ID_BYTES 128
int gsum_comm_unique_id(gsum_ctx *ctx, void *id_out);
int gsum_comm_init(gsum_ctx *ctx, int32_t nranks, int32_t rank, const void *nccl_unique_id);
int gsum_comm_destroy(gsum_ctx *ctx);
int gsum_grid_allgather(gsum_ctx *ctx, const double *block, int64_t n_q, int64_t per, int64_t n_ls, double *ll_full, double *post,
                        double *lse, int32_t mem_kind);
int gsum_comm_allreduce_counts(gsum_ctx *ctx, int64_t *counts, int64_t n, int32_t mem_kind);

/* fit with fixed kernel hyperparameters (gsum/models.py:671-738, optimizer=None / 'fixed' bounds):
 * factor R = c*RBF(X) + (noise + nugget) I once, keep X, y, L resident on the device, return the posterior
 * hyperparameters out7 = [center, disp, df, scale, cov_factor, log_marginal_likelihood, logdet R].
 * L_out (n,n) or NULL receives corr_L_.  Returns > 0 (LinAlgError) if R is not positive definite. */
int gsum_fit_create(gsum_ctx *ctx, const double *X, int64_t n, int32_t d, const double *y, int32_t n_c,
                    const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                    double center0, double disp0, double df0, double scale0, int32_t student, double *out7,
                    double *L_out, int32_t mem_kind, gsum_fit **fit);
int gsum_fit_destroy(gsum_fit *fit);

/* predict (gsum/models.py:753-845; Student-t addition 1128-1182; truncation scaling 1337-1354, 1389-1483).
 *
 * One GP conditional:  mean = mean_new + K_no K_oo^{-1} (yc - mean_old),   K_pred = K_nn - K_no K_oo^{-1} K_on.
 *   truncation == 0 (BaseConjugateProcess.predict, models.py:753-845): K = correlation R = c*RBF (+ noise I on
 *     k(X) diagonals), conditioning factor = chol(R_oo + nugget I) (the one kept by fit when Xc == NULL), and the
 *     returned (co)variance is cov_factor * R_pred (+ cov_factor * nugget on the diagonal if pred_noise).
 *   truncation != 0 (TruncationProcess.predict, models.py:1439-1452 / 1460-1473): every block is
 *     cov(X,X') = (sc[X] sc[X']) * gs(q[X] q[X']) * (cov_factor * c*RBF(X,X')) with NO white noise and NO nugget
 *     (models.py:1342-1348, helpers.geometric_sum), and K_oo is factored afresh (the reference uses an LU solve;
 *     K_oo is symmetric positive definite, so Cholesky gives the same result to rounding).
 * Hyperparameters (length scale, c, noise, nugget, cov_factor) always come from `fit`.
 * All pointers follow `mem_kind`; optional ones may be NULL. */
typedef struct gsum_predict_args {
    const double *Xnew; int64_t m;          /* (m, d) query points */
    const double *Xc; int64_t n_cond;       /* (n_cond, d) conditioning points; NULL = the X given to fit */
    const double *yc; int32_t n_y;          /* (n_cond, n_y) conditioning values; NULL = the y given to fit */
    const double *mean_old;                 /* (n_cond,) prior mean at the conditioning points; NULL = 0 */
    const double *mean_new;                 /* (m,) prior mean at the query points; NULL = 0 */
    const double *basis_old;                /* (n_cond,) and */
    const double *basis_new;                /* (m,): if cond_basis_out != NULL it receives basis_new - K_no K_oo^{-1} basis_old (models.py:1171, 1549) */
    const double *sc_old, *sc_new;          /* truncation: ref(X) at old / new points; NULL = 1 */
    const double *q_old, *q_new;            /* truncation: ratio(X) at old / new points; NULL = no geometric sum */
    double gs_start, gs_end;                /* geometric sum range; gs_end may be INFINITY */
    const int32_t *excluded; int32_t n_excluded;   /* host pointer, at most 8 entries */
    double kernel_add;                      /* truncation: added to the kernel value before cov_factor (disp_ for the
                                               Student-t process: cov = var * (corr + B V B^T), models.py:1124-1125); else 0 */
    int32_t truncation;
    int32_t want;                           /* GSUM_PREDICT_MEAN / _VAR / _COV */
    int32_t pred_noise;
    double *mean_out;                       /* (m, n_y) */
    double *var_out;                        /* _VAR: (m,) diagonal of K_pred (variances; the facade takes sqrt); _COV: (m, m) */
    double *cond_basis_out;                 /* (m,) or NULL */
} gsum_predict_args;
int gsum_predict(gsum_ctx *ctx, gsum_fit *fit, const gsum_predict_args *args, int32_t mem_kind);

/* cov(X, Xp) of the (truncation) process without conditioning (gsum/models.py:562-599, 1342-1348): out (n1, n2).
 * X2 == NULL evaluates k(X1) (WhiteKernel contributes on the diagonal, as sklearn does when Y is None).
 * Kernel = constant * RBF(ls) + noise I.  sc1/sc2/q1/q2 as in gsum_predict_args (NULL = no scaling); `factor`
 * multiplies (kernel + kernel_add) first: out = (sc sc') * gs(q q') * (factor * (k + kernel_add)). */
int gsum_process_cov(gsum_ctx *ctx, int32_t d, const double *ls, int32_t ls_dim, double constant, double noise,
                     const double *X1, int64_t n1, const double *X2, int64_t n2, const double *sc1, const double *sc2, const double *q1, const double *q2, double gs_start,
                     double gs_end, const int32_t *excluded, int32_t n_excluded, double factor, double kernel_add,
                     double *out, int32_t mem_kind);

/* Diagnostics.
 * gsum_cholesky_errors: E = L^{-1}(Y - mean)  (gsum/helpers.py:504-505, diagnostics.py:100-101); Y, E (n, n_curves);
 *   md2 (n_curves,) or NULL receives the squared Mahalanobis distances (helpers.py:512-517, diagnostics.py:112-114). */
int gsum_cholesky_errors(gsum_ctx *ctx, const double *L, int64_t n, const double *mean, const double *Y,
                         int64_t n_curves, double *E, double *md2, int32_t mem_kind);

/* gsum_quadratic_forms: q_c = (y_c - mean)^T A (y_c - mean) for every column c of Y (n, n_curves); A (n, n) row-major,
 *   not required to be symmetric.  Replaces np.diag((y - mean) @ inv @ (y - mean).T) of mahalanobis(inv=...)
 *   (gsum/helpers.py:521-522) without forming the (n_curves, n_curves) product. */
int gsum_quadratic_forms(gsum_ctx *ctx, const double *A, int64_t n, const double *mean, const double *Y, int64_t n_curves,
                         double *q, int32_t mem_kind);

/* gsum_pivoted_cholesky: LAPACK dpstrf(lower) semantics (gsum/helpers.py:185-199): M (n,n) symmetric PSD ->
 *   Lp (n,n) lower factor of P^T M P, piv (n,) 0-based, rank; G_out (n,n) or NULL = Lp[p_inv] (M = G G^T).
 *   Returns 1 (LinAlgError 'M is not positive-semidefinite') when rank < n, like helpers.py:189-190. */
int gsum_pivoted_cholesky(gsum_ctx *ctx, const double *M, int64_t n, double *Lp, int32_t *piv, int32_t *rank,
                          double *G_out, int32_t mem_kind);

/* gsum_pc_errors: solve(G, Y - mean) with G = Lp[p_inv] (gsum/diagnostics.py:103-104) by permute + forward substitution. */
int gsum_pc_errors(gsum_ctx *ctx, const double *Lp, const int32_t *piv, int64_t n, const double *mean,
                   const double *Y, int64_t n_curves, double *E, int32_t mem_kind);

/* gsum_draws: draws = mean + L z (gsum/diagnostics.py:82, models.py:872 up to the sampler's stream).
 *   Z (n, n_draws) caller-supplied standard normals, or NULL to generate them on the device (Philox, `seed`;
 *   draw j of this call is GLOBAL draw first_draw + j of the seed's stream, so the draw axis shards over GPUs,
 *   SURVEY.md 8e, with every rank reproducing its slice of the unsharded call).
 *   draw_scale (n_draws,) or NULL multiplies draw j's normals: multivariate-t draws are Gaussian draws times
 *   sqrt(df / chi2_df) (statsmodels MVT.rvs behind gsum/diagnostics.py:51-55).
 *   draws_out (n, n_draws) or NULL.  Credible-interval coverage fused into the same pass
 *   (gsum/diagnostics.py:148-171): lower, upper (n_alpha, n) interval bounds per point (or NULL to skip);
 *   coverage_out (n_draws, n_alpha) or NULL = mean over points of 1[lower < y < upper];
 *   count_out (n_alpha,) int64 or NULL = number of (draw, point) pairs inside each interval (exact integers: the
 *   quantity a sharded run all-reduces). */
int gsum_draws(gsum_ctx *ctx, const double *L, int64_t n, const double *mean, const double *Z, int64_t n_draws,
               uint64_t seed, int64_t first_draw, const double *draw_scale, double *draws_out, const double *lower,
               const double *upper, int32_t n_alpha, double *coverage_out, int64_t *count_out, int32_t mem_kind);

/* ---- TruncationPointwise (gsum/models.py:1573-1836) and VariogramFourthRoot (gsum/helpers.py:525-730) ----
 * gsum_pointwise_fit: per point x of y (n, n_o): coefficients of the kept orders (mask[k] = 1; helpers.py:71-101), the
 *   posterior scale sqrt((df0 scale0^2 + sum c^2) / (df0 + n_m)) (models.py:1630-1634) and the truncation-error scale
 *   ref * sqrt(geometric_sum(ratio^2, order+1, inf, excluded)) * scale per kept order (models.py:1683-1685).
 *   ratio, ref (n,); coeffs, trunc_scale (n, n_m); scale (n,); n_m = number of kept orders. */
int gsum_pointwise_fit(gsum_ctx *ctx, const double *y, int64_t n, int32_t n_o, const int32_t *orders, const int32_t *mask,
                       const int32_t *excluded, int32_t n_ex, const double *ratio, const double *ref, double df0, double scale0,
                       double *coeffs, double *scale, double *trunc_scale, int32_t mem_kind);
/* gsum_pointwise_loglike: the two sums over points of TruncationPointwise.log_likelihood (models.py:1796-1803) for n_r
 *   ratio sets at once: S1[r] = sum_x log(df scale_x^2 / 2), S2[r] = sum_b (log|ref_b| + sum(kept orders) log(ratio_b)),
 *   b over the numpy broadcast of ref (n_ref = 1 or n) and ratios[r] (n_rat = 1 or n). */
int gsum_pointwise_loglike(gsum_ctx *ctx, const double *y, int64_t n, int32_t n_o, const int32_t *orders, const int32_t *mask,
                           const double *ratios, int64_t n_r, int64_t n_rat, const double *ref, int64_t n_ref, double df0,
                           double scale0, double *S1, double *S2, int32_t mem_kind);
/* gsum_variogram_bins: distances and sqrt|z_i - z_j| of all pairs j < i (np.tril_indices order), their distance bins
 *   (np.digitize with `bounds`), the (n, n) grid of bins and the per-bin count / sum of distances / sum of sqrt|dz| per curve
 *   (helpers.py:547-600).  z (ncurves, n); hij, bin_idx (n(n-1)/2,), dij (n(n-1)/2, ncurves); counts (nbnd+1,) int64. */
int gsum_variogram_bins(gsum_ctx *ctx, const double *X, int64_t n, int32_t d, const double *z, int32_t ncurves,
                        const double *bounds, int32_t nbnd, int32_t *bin_grid, double *hij, int32_t *bin_idx, double *dij,
                        int64_t *counts, double *hsum, double *dsum, int32_t mem_kind);
/* gsum_variogram_cov: VariogramFourthRoot.cov(bin1, bin2) (helpers.py:640-696): mean over (pair in bin 1) x (pair in bin 2) of
 *   corr_ijkl sqrt(var_ij var_kl) with the fourth-root correlation (1 - rho^2) 2F1(3/4, 3/4; 1/2; rho^2) - 1 (helpers.py:625-638).
 *   (i1, j1), (i2, j2): the pairs of the two bins; gamma_tilde (nbins, ncurves), ncurves <= 8; tab: 3 * 56 + 2 series
 *   coefficients of the hypergeometric function (see csrc/pointwise.cuh); same_is_one: a pair with itself has correlation 1
 *   (cov_ijkl, helpers.py:655-657) or the formula's value (corr_ijkl); out (ncurves,). */
int gsum_variogram_cov(gsum_ctx *ctx, const int32_t *i1, const int32_t *j1, int64_t nb1, const int32_t *i2, const int32_t *j2,
                       int64_t nb2, const int32_t *bin_grid, int64_t n, const double *gamma_tilde, int32_t nbins, int32_t ncurves,
                       const double *tab, double var_factor, double corr_factor, int32_t same_is_one, double *out, int32_t mem_kind);

/* gsum_credible_interval: coverage of given curves Y (n, n_curves) (gsum/diagnostics.py:148-171). */
int gsum_credible_interval(gsum_ctx *ctx, const double *Y, int64_t n, int64_t n_curves, const double *lower,
                           const double *upper, int32_t n_alpha, double *coverage_out, int32_t mem_kind);

/* decomposition='eig' route.
 * gsum_eigh: symmetric eigendecomposition A = V diag(w) V^T — scipy `eigh(R)` at gsum/models.py:714, 811, 974, 1166,
 *   1216 and numpy `eigh(cov)` at gsum/diagnostics.py:63.  A (n,n) symmetric (only read); w (n,) ascending;
 *   V (n,n) row-major with the eigenvectors in its COLUMNS, each returned with its largest-magnitude component
 *   positive (LAPACK fixes no sign; compare up to sign).  One-sided Jacobi on the device; sweeps_out (host int32, or
 *   NULL) receives the number of sweeps.  Returns 1 when the iteration did not converge within its sweep limit and 2
 *   for an indefinite A holding a pair of eigenvalues +lambda, -lambda (the one-sided iteration cannot separate those;
 *   the covariance / correlation matrices of the cited call sites are positive semi-definite up to rounding). */
int gsum_eigh(gsum_ctx *ctx, const double *A, int64_t n, double *w, double *V, int32_t *sweeps_out, int32_t mem_kind);

/* gsum_lml_grad_terms_eig: gsum_lml_grad_terms on the 'eig' route — the same outputs with R^-1 taken from the device
 *   eigendecomposition (Q diag(1/eig) Q^T, gsum/models.py:480-484, 1043) and logdet R = sum log eig (models.py:1019);
 *   info is always 0 (the reference's 'eig' branch has no failure exit, models.py:973-974). */
int gsum_lml_grad_terms_eig(gsum_ctx *ctx, const double *X, int64_t n, int32_t d, const double *RHS, int32_t r,
                            const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                            double *G, double *H, double *tr, double *logdet, int32_t *info, int32_t mem_kind);

/* gsum_eig_solve: with (w, V) from gsum_eigh and Y (n, nrhs), mean (n,) or NULL (subtracted from every column of Y):
 *   mode 0: X = V diag(1/w) V^T (Y - mean)            `solve_sqrt(..., decomposition='eig')`, gsum/models.py:480-484
 *   mode 1: X = diag(|w|^-1/2) V^T (Y - mean)         `eigen_errors`, gsum/diagnostics.py:63-68, 106-107 (row k of X
 *           belongs to w[k]: the caller reverses the rows for the reference's largest-first order)
 *   X (n, nrhs).  With GSUM_MEM_FACTOR_DEVICE, w and V are device pointers and Y / mean / X host pointers. */
int gsum_eig_solve(gsum_ctx *ctx, const double *w, const double *V, int64_t n, const double *Y, int64_t nrhs,
                   const double *mean, double *X, int32_t mode, int32_t mem_kind);

/* gsum_eig_conditional: the conditioning products of `predict` on the 'eig' route (gsum/models.py:826-836, 1174):
 *   with R^-1 = V diag(1/w) V^T, R_on (n, m) the cross-correlation and D (n, k) the columns to condition on
 *   (y - mean_old and, for the Student-t process, the old basis):
 *     lin_out (m, k)  = R_no R^-1 D                        (or NULL)
 *     var_out (m,)    = diag(R_no R^-1 R_on)               (or NULL)
 *     cov_out (m, m)  = R_no R^-1 R_on                     (or NULL)
 *   computed as U = V^T R_on once, then U^T diag(1/w) [V^T D | U].  GSUM_MEM_FACTOR_DEVICE as for gsum_eig_solve. */
int gsum_eig_conditional(gsum_ctx *ctx, const double *w, const double *V, int64_t n, const double *R_on, int64_t m,
                         const double *D, int64_t k, double *lin_out, double *var_out, double *cov_out, int32_t mem_kind);

#ifdef __cplusplus
}
#endif
#endif /* GSUM_B200_H */
