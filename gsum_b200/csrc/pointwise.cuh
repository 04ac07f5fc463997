// the pointwise truncation model and the fourth-root variogram.
//
//   TruncationPointwise (gsum/models.py:1573-1836): per point, coefficients of the partial sums, the scaled-inverse-chi^2
//   posterior (df, scale) and the Student-t truncation-error scale; its log-likelihood over a set of expansion parameters.
//   Elementwise / per-point reductions: HBM-bound, one pass over y per ratio set.
//
//   VariogramFourthRoot (gsum/helpers.py:525-730): pairwise distances and sqrt|z_i - z_j| binned by distance, and the
//   covariance of two bins — a sum over all (pair in bin 1) x (pair in bin 2) of a correlation that needs
//   2F1(3/4, 3/4; 1/2; rho^2): O(N^4) index-pair work, FP64-vector bound.
#pragma once
#include "common.cuh"

#define PW_MAXO 32                      // expansion orders per fit
#define VG_MAXC 8                       // curves per variogram pass
#define VG_NT 56                        // terms of the two hypergeometric series

struct PointwiseArgs {
    const double *y;        // (n, n_o) partial sums
    const int *orders;      // (n_o)
    const int *mask;        // (n_o) 1 = order takes part (not excluded)
    const int *excluded;    // (n_ex) excluded orders
    int n_ex;
    int64_t n;
    int n_o;
    double df0, scale0;
};

// c_k = (y_k - y_{k-1}) / (ref * ratio^order_k)                                       (gsum/helpers.py:96-100)
__device__ __forceinline__ double pw_coeff(const double *yrow, int k, double ref, double ratio, int order) {
    const double dy = k == 0 ? yrow[0] : yrow[k] - yrow[k - 1];
    return dy / (ref * pow(ratio, (double)order));
}

// fit (models.py:1651-1690): coeffs (n, n_m), scale (n), trunc_scale (n, n_m) = ref * sqrt(geometric_sum(ratio^2, k+1, inf,
// excluded)) * scale for every kept order k (helpers.py:149-182)
__global__ void pointwise_fit_kernel(PointwiseArgs P, const double *__restrict__ ratio, const double *__restrict__ ref,
                                     double *__restrict__ coeffs, double *__restrict__ scale, double *__restrict__ trunc_scale, int n_m) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.n) return;
    const double *yrow = P.y + x * P.n_o;
    const double q = ratio[x], r = ref[x];
    double csq = 0.0;
    int m = 0;
    for (int k = 0; k < P.n_o; k++) {
        if (!P.mask[k]) continue;
        const double c = pw_coeff(yrow, k, r, q, P.orders[k]);
        coeffs[x * n_m + m] = c;
        csq += c * c;
        m++;
    }
    const double df = P.df0 + (double)n_m;
    const double sc = sqrt((P.df0 * P.scale0 * P.scale0 + csq) / df);
    scale[x] = sc;
    const double x2 = q * q;
    m = 0;
    for (int k = 0; k < P.n_o; k++) {
        if (!P.mask[k]) continue;
        const int start = P.orders[k] + 1;
        double s = (pow(x2, (double)start) - pow(x2, (double)INFINITY)) / (1.0 - x2);
        for (int e = 0; e < P.n_ex; e++)
            if (P.excluded[e] >= start) s -= pow(x2, (double)P.excluded[e]);
        trunc_scale[x * n_m + m] = r * sqrt(s) * sc;
        m++;
    }
}

// log_likelihood (models.py:1762-1804) for n_r ratio sets at once: the two sums over points
//   S1[r] = sum_x log(df * scale_x^2 / 2),   S2[r] = sum_b (log|ref_b| + sum(orders kept) * log(ratio_b)),
// b running over the broadcast of ref (n_ref in {1, n}) and ratio (n_rat in {1, n}) exactly as numpy broadcasts them there
// (a scalar ratio with a scalar ref gives ONE Jacobian term, not n: the reference's own convention).
__global__ void __launch_bounds__(256) pointwise_loglike_kernel(PointwiseArgs P, const double *__restrict__ ratios, int n_rat,
                                                                 const double *__restrict__ ref, int n_ref, int n_m, double so,
                                                                 double *__restrict__ S1, double *__restrict__ S2) {
    __shared__ double red[32];
    const int r = blockIdx.x;
    const double *rat = ratios + (int64_t)r * n_rat;
    const double df = P.df0 + (double)n_m;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t x = threadIdx.x; x < P.n; x += blockDim.x) {
        const double q = rat[n_rat == 1 ? 0 : x], rf = ref[n_ref == 1 ? 0 : x];
        const double *yrow = P.y + x * P.n_o;
        double csq = 0.0;
        for (int k = 0; k < P.n_o; k++) {
            if (!P.mask[k]) continue;
            const double c = pw_coeff(yrow, k, rf, q, P.orders[k]);
            csq += c * c;
        }
        const double sc2 = (P.df0 * P.scale0 * P.scale0 + csq) / df;
        s1 += log(df * sc2 / 2.0);
    }
    const int64_t nb = (n_rat > n_ref ? n_rat : n_ref);
    for (int64_t b = threadIdx.x; b < nb; b += blockDim.x)
        s2 += log(fabs(ref[n_ref == 1 ? 0 : b])) + so * log(rat[n_rat == 1 ? 0 : b]);
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { S1[r] = s1; S2[r] = s2; }
}

// ---- variogram --------------------------------------------------------------------------------------------------------
// np.digitize(h, bounds) with increasing bounds: the number of bounds <= h
__device__ __forceinline__ int vg_digitize(double h, const double *bounds, int nbnd) {
    int lo = 0, hi = nbnd;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (bounds[mid] <= h) lo = mid + 1; else hi = mid; }
    return lo;
}
// bin_grid (n, n): the bin of ||X_i - X_j|| for every (i, j)                       (helpers.py:549-551)
__global__ void variogram_grid_kernel(const double *__restrict__ X, int64_t n, int d, const double *__restrict__ bounds, int nbnd,
                                      int *__restrict__ bin_grid) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    const int64_t i = e / n, j = e % n;
    double s = 0.0;
    for (int a = 0; a < d; a++) { const double v = X[i * d + a] - X[j * d + a]; s += v * v; }
    bin_grid[e] = vg_digitize(sqrt(s), bounds, nbnd);
}
// the strict lower triangle in np.tril_indices(n, -1) order: pair p <-> (i, j), j < i   (helpers.py:574-576)
__device__ __forceinline__ void vg_pair(int64_t p, int64_t &i, int64_t &j) {
    i = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
    while (i * (i - 1) / 2 > p) i--;
    while ((i + 1) * i / 2 <= p) i++;
    j = p - i * (i - 1) / 2;
}
// per pair: distance, bin, sqrt|z_i - z_j| per curve                                (helpers.py:567-572)
__global__ void variogram_pairs_kernel(const double *__restrict__ X, const double *__restrict__ z, int64_t n, int d, int ncurves,
                                       const double *__restrict__ bounds, int nbnd, double *__restrict__ hij, int *__restrict__ bin_idx,
                                       double *__restrict__ dij) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n * (n - 1) / 2) return;
    int64_t i, j;
    vg_pair(p, i, j);
    double s = 0.0;
    for (int a = 0; a < d; a++) { const double v = X[i * d + a] - X[j * d + a]; s += v * v; }
    const double h = sqrt(s);
    hij[p] = h;
    bin_idx[p] = vg_digitize(h, bounds, nbnd);
    for (int c = 0; c < ncurves; c++) dij[p * ncurves + c] = sqrt(fabs(z[(int64_t)c * n + i] - z[(int64_t)c * n + j]));
}
// per bin: count, sum of distances, sum of sqrt|dz| per curve — one CTA per bin, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) variogram_binsum_kernel(const double *__restrict__ hij, const int *__restrict__ bin_idx,
                                                                const double *__restrict__ dij, int64_t npairs, int ncurves,
                                                                long long *__restrict__ counts, double *__restrict__ hsum,
                                                                double *__restrict__ dsum) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    double cnt = 0.0, hs = 0.0;
    for (int64_t p = threadIdx.x; p < npairs; p += blockDim.x)
        if (bin_idx[p] == b) { cnt += 1.0; hs += hij[p]; }
    cnt = block_sum(cnt, red);
    hs = block_sum(hs, red);
    if (threadIdx.x == 0) { counts[b] = (long long)(cnt + 0.5); hsum[b] = hs; }
    for (int c = 0; c < ncurves; c++) {
        double ds = 0.0;
        for (int64_t p = threadIdx.x; p < npairs; p += blockDim.x)
            if (bin_idx[p] == b) ds += dij[p * ncurves + c];
        ds = block_sum(ds, red);
        if (threadIdx.x == 0) dsum[b * ncurves + c] = ds;
    }
}

// F(z) = (1 - z) 2F1(3/4, 3/4; 1/2; z) = 2F1(-1/4, -1/4; 1/2; z) on [0, 1)  (Euler's transformation), evaluated by its Maclaurin
// series for z <= 1/2 and, above, by the logarithmic expansion in w = 1 - z of the case c = a + b + 1 (Abramowitz & Stegun
// 15.3.11):  F = K0 + K1 * sum_n A_n w^(n+1) (ln w + B_n).  tab = [a_n | A_n | B_n | K0, K1], VG_NT terms each (host-made from
// Gamma / digamma); both series agree with scipy.special.hyp2f1 to 3e-15 on [0, 1 - 1e-12].
__device__ __forceinline__ double vg_hyp(double z, const double *__restrict__ tab) {
    if (z <= 0.5) {
        double s = 0.0;
#pragma unroll 8
        for (int n = VG_NT - 1; n >= 0; n--) s = fma(s, z, tab[n]);
        return s;
    }
    const double w = 1.0 - z, lw = log(w);
    double s = 0.0;
#pragma unroll 8
    for (int n = VG_NT - 1; n >= 0; n--) s = fma(s, w, tab[VG_NT + n] * (lw + tab[2 * VG_NT + n]));
    return tab[3 * VG_NT] + tab[3 * VG_NT + 1] * w * s;
}
struct VarioCovArgs {
    const int *i1, *j1;     // pairs of bin 1 (nb1)
    const int *i2, *j2;     // pairs of bin 2 (nb2)
    int64_t nb1, nb2;
    const int *bin_grid;    // (n, n)
    int64_t n;
    const double *gamma_tilde;   // (Nb, ncurves)
    int ncurves;
    const double *tab;
    double var_factor, corr_factor;
    int same_is_one;        // 1: (i, j) == (k, l) has correlation 1 (cov_ijkl, helpers.py:655-657); 0: the formula's value (corr_ijkl)
};
// cov(bin1, bin2) * nb1 * nb2 (helpers.py:640-696): sum over all (ij) x (kl) of corr_ijkl * sqrt(var_ij var_kl).  Each CTA takes
// a tile of 16 pairs of bin 1 x all pairs of bin 2 (strided over its threads) and writes one partial per curve.
__global__ void __launch_bounds__(256) variogram_cov_kernel(VarioCovArgs P, double *__restrict__ partial) {
    __shared__ double red[32];
    const int64_t p1_0 = (int64_t)blockIdx.x * 16;
    double acc[VG_MAXC];
#pragma unroll
    for (int c = 0; c < VG_MAXC; c++) acc[c] = 0.0;
    const int64_t p1_n = (P.nb1 - p1_0) < 16 ? (P.nb1 - p1_0) : 16;
    for (int64_t e = threadIdx.x; e < p1_n * P.nb2; e += blockDim.x) {
        const int64_t a = p1_0 + e / P.nb2, b = e % P.nb2;
        const int i = P.i1[a], j = P.j1[a], k = P.i2[b], l = P.j2[b];
        const int64_t n = P.n;
        const int b_jk = P.bin_grid[j * n + k], b_il = P.bin_grid[i * n + l], b_ik = P.bin_grid[i * n + k], b_jl = P.bin_grid[j * n + l];
        const int b_ij = P.bin_grid[i * n + j], b_kl = P.bin_grid[k * n + l];
        const bool same = P.same_is_one && (i == k) && (j == l);
#pragma unroll
        for (int c = 0; c < VG_MAXC; c++) {
            if (c >= P.ncurves) break;
            const double *g = P.gamma_tilde + c;
            const int nc = P.ncurves;
            const double g_ij = g[b_ij * nc], g_kl = g[b_kl * nc];
            double corr = 1.0;
            if (!same) {
                const double rho = (g[b_jk * nc] + g[b_il * nc] - g[b_ik * nc] - g[b_jl * nc]) / (2.0 * sqrt(g_ij * g_kl));
                if (rho >= 1.0) corr = 1.0;
                else if (rho <= -1.0) corr = -1.0;
                else if (rho != rho) corr = rho;                    // NaN stays NaN, as in numpy
                else corr = P.corr_factor * (vg_hyp(rho * rho, P.tab) - 1.0);
            }
            const double v_ij = P.var_factor * sqrt(g_ij), v_kl = P.var_factor * sqrt(g_kl);
            acc[c] += corr * sqrt(v_ij * v_kl);
        }
    }
#pragma unroll
    for (int c = 0; c < VG_MAXC; c++) {
        if (c >= P.ncurves) break;
        const double s = block_sum(acc[c], red);
        if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * P.ncurves + c] = s;
    }
}
__global__ void __launch_bounds__(256) variogram_cov_reduce_kernel(const double *__restrict__ partial, int64_t nblocks, int ncurves,
                                                                    double denom, double *__restrict__ out) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[b * ncurves + c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[c] = s / denom;
}
