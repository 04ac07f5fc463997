// K3/K5/K6 helpers around the bordered solver: staging right-hand sides as (transposed) border rows,
// reading solved rows back, and the memory-bound row reductions that turn  Vt = R_no L^{-T}  and
// w = L^{-1}(y - m)  into posterior means / variances / Mahalanobis distances.
#pragma once
#include "common.cuh"

// dst rows (cols x dst_ld, zero padded to rows_pad x dst_ld)  <-  sign * (src (rows_src x cols) - sub[row])^T
// i.e. border row c holds column c of the (n x m) source; `perm` optionally gathers source rows (pivot order);
// `flip` reverses the row index (backward solves run as forward solves on the flipped system).
__global__ void __launch_bounds__(256) transpose_in_kernel(const double *__restrict__ src, int64_t n, int64_t m,
                                                           const double *__restrict__ sub, const int32_t *__restrict__ perm,
                                                           int flip, double sign, double *__restrict__ dst, int64_t dst_ld,
                                                           int64_t rows_pad) {
    __shared__ double tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;   // x: source row (point), c: source col (rhs)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t x = x0 + r, cc = c0 + tx;
        double v = 0.0;
        if (x < n && cc < m) {
            int64_t xs = flip ? (n - 1 - x) : x;
            if (perm) xs = perm[xs];
            v = src[xs * m + cc];
            if (sub) v -= sub[xs];
            v *= sign;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t cc = c0 + r, x = x0 + tx;
        if (cc < rows_pad && x < dst_ld) dst[cc * dst_ld + x] = tile[tx][r];
    }
}

// dst (n x m)  <-  scale * rows(src)^T [+ add[row]]    (inverse of the above; `flip` un-reverses the point index)
__global__ void __launch_bounds__(256) transpose_out_kernel(const double *__restrict__ src, int64_t src_ld, int64_t n, int64_t m,
                                                            int flip, double scale, const double *__restrict__ add,
                                                            double *__restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t x0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t cc = c0 + r, x = x0 + tx;
        tile[r][tx] = (cc < m && x < n) ? src[cc * src_ld + x] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t x = x0 + r, cc = c0 + tx;
        if (x < n && cc < m) {
            const int64_t xd = flip ? (n - 1 - x) : x;
            double v = scale * tile[tx][r];
            if (add) v += add[xd];
            dst[xd * m + cc] = v;
        }
    }
}

// out[j][c] = base[j] + sign * sum_x Vt[j][x] * Wy[c][x]      (posterior mean  m_new + R_no R^-1 (y - m_old), gsum/models.py:831-832)
// One warp per row j; ny small.  Fixed summation order.
__global__ void __launch_bounds__(256) rows_dot_kernel(const double *__restrict__ Vt, int64_t ld, int64_t m, int64_t n,
                                                       const double *__restrict__ Wy, int ny, const double *__restrict__ base,
                                                       double sign, double *__restrict__ out, int64_t out_ld) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= m) return;
    const double *v = Vt + j * ld;
    for (int c = 0; c < ny; c++) {
        const double *w = Wy + (int64_t)c * ld;
        double s = 0.0;
        for (int64_t x = lane; x < n; x += 32) s += v[x] * w[x];
        s = warp_sum(s);
        if (lane == 0) out[j * out_ld + c] = (base ? base[j] : 0.0) + sign * s;
    }
}

// out[j] = sum_x Vt[j][x]^2      (squared Mahalanobis distance, gsum/helpers.py:512-517; and diag(V^T V) for return_std)
__global__ void __launch_bounds__(256) rows_sqnorm_kernel(const double *__restrict__ Vt, int64_t ld, int64_t m, int64_t n,
                                                          double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= m) return;
    const double *v = Vt + j * ld;
    double s = 0.0;
    for (int64_t x = lane; x < n; x += 32) s += v[x] * v[x];
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

// ---- truncation-process covariance scaling (gsum/models.py:1342-1348, helpers.py:149-182) ---------------------
// gs(x) = (x^start - x^(end+1)) / (1 - x) - sum_{n in excluded, start<=n<=end} x^n      (end may be +inf: x^inf -> 0 for |x|<1)
struct GeoSum {
    int enabled;
    double start, end;
    int n_excl;
    int excl[8];
};
__device__ __forceinline__ double geo_sum(const GeoSum &g, double x) {
    double s = (pow(x, g.start) - pow(x, g.end + 1.0)) / (1.0 - x);
    for (int e = 0; e < g.n_excl; e++) {
        const double ne = (double)g.excl[e];
        if (ne >= g.start && ne <= g.end) s -= pow(x, ne);
    }
    return s;
}

// In place on a (rows x cols) block with leading dimension ld:
//   K[r][c] <- ((sc_r[r] * sc_c[c]) * gs(q_r[r] * q_c[c])) * (factor * (K[r][c] + kadd))      — same association as the reference
// (kadd = disp for the Student-t process, whose cov is var * (corr + B V B^T), gsum/models.py:1124-1125)
// rows beyond `rows` / cols beyond `cols` are left untouched.  If `unit_pad_from` >= 0 the block is square and the
// diagonal entries with index >= unit_pad_from are set to 1 (identity padding of a factor block).
__global__ void __launch_bounds__(256) scale_cov_kernel(double *K, int64_t ld, int64_t rows, int64_t cols,
                                                        const double *__restrict__ sc_r, const double *__restrict__ sc_c,
                                                        const double *__restrict__ q_r, const double *__restrict__ q_c, GeoSum g,
                                                        double factor, double kadd) {
    const int64_t r = blockIdx.x;
    if (r >= rows) return;
    const double sr = sc_r ? sc_r[r] : 1.0;
    const double qr = q_r ? q_r[r] : 0.0;
    for (int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.y * blockDim.x) {
        double ref_mat = sr * (sc_c ? sc_c[c] : 1.0);
        double ratio_sum = (g.enabled && q_r) ? geo_sum(g, qr * q_c[c]) : 1.0;
        K[r * ld + c] = (ref_mat * ratio_sum) * (factor * (K[r * ld + c] + kadd));
    }
}

// q_c = d_c^T A d_c, d_c = y_c - mean (mahalanobis(inv=...), gsum/helpers.py:521-522).  One warp per row i of A: the lanes
// stride over the columns (coalesced row reads), each accumulating (A d)_i for QF_C curves at a time; the centred curves of
// the chunk sit in shared memory.  part[(i, c)] = d_ic (A d_c)_i is written per row and summed in a fixed order by
// quadform_reduce_kernel: deterministic, no floating-point atomics.  HBM traffic: A once per QF_C curves.
#define QF_C 8
#define QF_WARPS 8
__global__ void __launch_bounds__(32 * QF_WARPS) quadform_rows_kernel(const double *__restrict__ A, int64_t n, const double *__restrict__ mean,
                                                                        const double *__restrict__ Y, int64_t n_curves, int64_t c0,
                                                                        double *__restrict__ part) {
    extern __shared__ double qf_d[];                 // (n, QF_C) centred curves of this chunk
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nc = (int)((n_curves - c0) < QF_C ? (n_curves - c0) : QF_C);
    for (int64_t e = tid; e < n * QF_C; e += blockDim.x) {
        const int64_t j = e / QF_C; const int c = (int)(e % QF_C);
        qf_d[e] = c < nc ? Y[j * n_curves + c0 + c] - mean[j] : 0.0;
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * QF_WARPS + w; i < n; i += (int64_t)gridDim.x * QF_WARPS) {
        double acc[QF_C];
#pragma unroll
        for (int c = 0; c < QF_C; c++) acc[c] = 0.0;
        const double *row = A + i * n;
        for (int64_t j = lane; j < n; j += 32) {
            const double a = row[j];
#pragma unroll
            for (int c = 0; c < QF_C; c++) acc[c] = fma(a, qf_d[j * QF_C + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < QF_C; c++) {
            const double v = warp_sum(acc[c]);
            if (lane == 0 && c < nc) part[i * QF_C + c] = qf_d[i * QF_C + c] * v;
        }
    }
}
__global__ void quadform_reduce_kernel(const double *__restrict__ part, int64_t n, int64_t n_curves, int64_t c0, double *__restrict__ q) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    if (c0 + c >= n_curves) return;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += part[i * QF_C + c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) q[c0 + c] = s;
}
