// K1: batched RBF (+ ConstantKernel scale, + WhiteKernel / nugget diagonal) covariance builder.
//
// Mirrors the arithmetic order of scikit-learn's RBF.__call__ (sklearn/gaussian_process/kernels.py,
// RBF: `pdist(X / length_scale, 'sqeuclidean')` -> `exp(-0.5 * d)` -> diagonal forced to exactly 1;
// with Y: `cdist(X / l, Y / l)`), of Product/Sum (constant * rbf + white) and of the reference's
// nugget (`R[diag] += nugget`, gsum/models.py:963; `corr_ + nugget * I`, models.py:711):
//     off-diagonal:  c * exp(-0.5 * sum_d (x_id/l_d - x_jd/l_d)^2)
//     diagonal    :  (c * 1 + noise) + nugget
// The coordinates are divided by the length scale first (true division, as numpy does) and the squared
// distance is accumulated without FMA contraction.
#pragma once
#include "common.cuh"

#define COV_MAXD 8              // input dimensions supported by the device kernels

// XS[b][n][d] = X[n][d] / ls[b][d or 0]
__global__ void scale_coords_kernel(const double *__restrict__ X, const double *__restrict__ ls, double *__restrict__ XS,
                                    int64_t n, int d, int ls_dim, int64_t batch) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = batch * n * d;
    if (idx >= total) return;
    int dd = (int)(idx % d);
    int64_t b = idx / (n * d);
    int64_t nd = idx % (n * d);
    double l = ls[b * ls_dim + (ls_dim == 1 ? 0 : dd)];
    XS[idx] = X[nd] / l;
}

__device__ __forceinline__ double rbf_sqdist(const double *xi, const double *xj, int d) {
    double s = 0.0;
    for (int q = 0; q < d; q++) {
        const double df = __dsub_rn(xi[q], xj[q]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

struct CovArgs {
    const double *XS;     // (batch, n, d) scaled coordinates
    int64_t n; int d;
    double constant, noise, nugget;
    double *A;            // bordered batch base
    int64_t ld, bstride;
    int T;                // tile rows/cols of the factor part (n padded to 64)
    int tiles_per_cta;    // lower tiles handled by one CTA (cov_tiles_per_cta)
};

// One 64x64 tile with every row and column inside the matrix.  Thread = two adjacent columns (kept in registers) x eight
// rows (row coordinates are warp-wide broadcasts from shared memory); same arithmetic order as the general loop.
template <int D>
__device__ __forceinline__ void cov_tile_inside(const double *xr, const double *xc, double constant, double dval, bool diag_tile,
                                                double *At, int64_t ld, int tid) {
    const int c = (tid & 31) * 2, rbase = tid >> 5;
    double c0[D], c1[D];
#pragma unroll
    for (int q = 0; q < D; q++) { c0[q] = xc[c * COV_MAXD + q]; c1[q] = xc[(c + 1) * COV_MAXD + q]; }
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int r = rbase + 8 * it;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) {
            const double x = xr[r * COV_MAXD + q];
            const double d0 = __dsub_rn(x, c0[q]), d1 = __dsub_rn(x, c1[q]);
            s0 = q == 0 ? __dmul_rn(d0, d0) : __dadd_rn(s0, __dmul_rn(d0, d0));       // 0 + d^2 == d^2
            s1 = q == 0 ? __dmul_rn(d1, d1) : __dadd_rn(s1, __dmul_rn(d1, d1));
        }
        double v0 = constant * rbf_exp_neg(-0.5 * s0);
        double v1 = constant * rbf_exp_neg(-0.5 * s1);
        if (diag_tile) { if (r == c) v0 = dval; if (r == c + 1) v1 = dval; }
        *reinterpret_cast<double2 *>(At + (int64_t)r * ld + c) = make_double2(v0, v1);
    }
}

// Symmetric build: lower tiles (i >= k) of each matrix in the batch; identity in the padding.
// One CTA (256 threads) per 64x64 tile; each thread produces 8 adjacent pairs -> 16-byte coalesced stores.
// Tiles per CTA (measured on C4: 1 -> 0.30 ms with the general loop, 4 -> 0.185 ms, 8 -> 0.195, 34 -> 0.24)
// With few matrices (a strong-scaled shard) four tiles per CTA leave the launch a single under-filled wave of long CTAs
// (16 length scales: 27 us); the count adapts so that the grid keeps at least ~2 waves of 8 resident CTAs per SM.
#ifndef COV_TILES_PER_CTA
#define COV_TILES_PER_CTA 4
#endif
static inline int cov_tiles_per_cta(int64_t tiles_per_matrix, int64_t batch, int sm_count) {
    int t = COV_TILES_PER_CTA;
    while (t > 1 && tiles_per_matrix * batch / t < (int64_t)2 * 8 * sm_count) t >>= 1;
    return t;
}
__global__ void __launch_bounds__(256) cov_sym_kernel(CovArgs P) {
    __shared__ double xr[GSUM_TILE * COV_MAXD], xc[GSUM_TILE * COV_MAXD];
    const int64_t b = blockIdx.y;
    const double *XS = P.XS + b * P.n * P.d;
    const int tid = threadIdx.x;
    double *A = P.A + b * P.bstride;
    const double dval = __dadd_rn(__dadd_rn(P.constant, P.noise), P.nugget);
    const int ntri = P.T * (P.T + 1) / 2;
    for (int tix = blockIdx.x * P.tiles_per_cta; tix < ntri && tix < (blockIdx.x + 1) * P.tiles_per_cta; tix++) {
        // decode lower-triangular tile index
        int i = (int)((sqrt(8.0 * tix + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= tix) i++;
        while (i * (i + 1) / 2 > tix) i--;
        const int k = tix - i * (i + 1) / 2;
        __syncthreads();                                   // the previous tile's coordinates are no longer in use
        for (int e = tid; e < GSUM_TILE * P.d; e += 256) {
            int r = e / P.d, q = e % P.d;
            int64_t gr = (int64_t)i * GSUM_TILE + r, gc = (int64_t)k * GSUM_TILE + r;
            xr[r * COV_MAXD + q] = gr < P.n ? XS[gr * P.d + q] : 0.0;
            xc[r * COV_MAXD + q] = gc < P.n ? XS[gc * P.d + q] : 0.0;
        }
        __syncthreads();
        // Tiles that lie entirely inside the matrix (all of them when N is a multiple of 64) take a path without bounds
        // checks whose inner loop is the arithmetic only.
        if ((int64_t)(i + 1) * GSUM_TILE <= P.n && P.d <= 3) {
            double *At = A + (int64_t)i * GSUM_TILE * P.ld + (int64_t)k * GSUM_TILE;
            if (P.d == 1) cov_tile_inside<1>(xr, xc, P.constant, dval, i == k, At, P.ld, tid);
            else if (P.d == 2) cov_tile_inside<2>(xr, xc, P.constant, dval, i == k, At, P.ld, tid);
            else cov_tile_inside<3>(xr, xc, P.constant, dval, i == k, At, P.ld, tid);
            continue;
        }
#pragma unroll 4
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += 256) {
            const int r = e >> 5, c = (e & 31) * 2;
            const int64_t gr = (int64_t)i * GSUM_TILE + r, gc = (int64_t)k * GSUM_TILE + c;
            double v[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t cc = gc + u;
                const double arg = -0.5 * rbf_sqdist(xr + r * COV_MAXD, xc + (c + u) * COV_MAXD, P.d);
                const double ex = P.constant * rbf_exp_neg(arg);
                if (gr >= P.n || cc >= P.n) v[u] = (gr == cc) ? 1.0 : 0.0;
                else if (gr == cc) v[u] = dval;
                else v[u] = ex;
            }
            *reinterpret_cast<double2 *>(A + gr * P.ld + gc) = make_double2(v[0], v[1]);
        }
    }
}

// General (cross-)covariance  K[r][c] = scale_r[r] * scale_c[c] * g(r,c) * c0 * exp(-0.5 |x1_r/l - x2_c/l|^2)  [+ diag terms]
// written to an arbitrary row-major destination.  `sym_diag`: treat r == c as the diagonal of k(X) (exactly 1, plus
// noise + nugget) — the `Y is None` branch; otherwise the WhiteKernel contributes nothing (sklearn WhiteKernel
// returns zeros when Y is given — the quirk gsum/models.py:583,824,1118 documents).
struct CrossArgs {
    const double *XS1; int64_t n1;     // scaled coords of rows
    const double *XS2; int64_t n2;     // scaled coords of cols
    int d;
    double constant, diag_add;          // diag_add = noise (+ nugget) applied when sym_diag and r == c
    int sym_diag;
    double *out; int64_t ldo;           // destination (rows_out x ldo); entries beyond (n1, n2) up to the padded extents are zero-filled
    int64_t rows_out, cols_out;
};
__global__ void __launch_bounds__(256) cov_cross_kernel(CrossArgs P) {
    __shared__ double xr[GSUM_TILE * COV_MAXD], xc[GSUM_TILE * COV_MAXD];
    const int64_t r0 = (int64_t)blockIdx.y * GSUM_TILE, c0 = (int64_t)blockIdx.x * GSUM_TILE;
    const int tid = threadIdx.x;
    for (int e = tid; e < GSUM_TILE * P.d; e += 256) {
        int r = e / P.d, q = e % P.d;
        xr[r * COV_MAXD + q] = (r0 + r) < P.n1 ? P.XS1[(r0 + r) * P.d + q] : 0.0;
        xc[r * COV_MAXD + q] = (c0 + r) < P.n2 ? P.XS2[(c0 + r) * P.d + q] : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += 256) {
        const int r = e >> 5, c = (e & 31) * 2;
        const int64_t gr = r0 + r;
        if (gr >= P.rows_out) continue;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int64_t gc = c0 + c + u;
            if (gc >= P.cols_out) continue;
            double v;
            if (gr >= P.n1 || gc >= P.n2) v = 0.0;
            else if (P.sym_diag && gr == gc) v = __dadd_rn(P.constant, P.diag_add);
            else v = P.constant * rbf_exp_neg(-0.5 * rbf_sqdist(xr + r * COV_MAXD, xc + (c + u) * COV_MAXD, P.d));
            P.out[gr * P.ldo + gc] = v;
        }
    }
}

// Border rows: dst[b][T*64 + rr][c] = src[rr][c] for rr < r, c < n; zero elsewhere (rows up to Rpad, cols up to ld).
__global__ void border_fill_kernel(double *A, int64_t ld, int64_t bstride, int T, int Rpad, const double *__restrict__ src,
                                   int r, int64_t n, int64_t src_ld) {
    const int64_t b = blockIdx.z;
    const int rr = blockIdx.x;             // rows on x: the border may hold > 65535 rows
    double *dst = A + b * bstride + ((int64_t)T * GSUM_TILE + rr) * ld;
    for (int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; c < ld; c += (int64_t)gridDim.y * blockDim.x)
        dst[c] = (rr < r && c < n) ? src[(int64_t)rr * src_ld + c] : 0.0;
}
