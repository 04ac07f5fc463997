// K4: fused normal-inverse-chi^2 reduction + Gaussian / Student-t marginal log-likelihood.
//
// After the bordered factorisation the border rows hold Wt = [B | C]^T L^{-T}.  `gram_rows_kernel`
// reduces them to the small Gram  G = W^T W  (everything the conjugate updates need: tr G, 1^T G 1, h, b),
// and `lml_cell_kernel` evaluates one (Q, l) cell in O(n_c^2) from it — the closed forms of
// gsum/models.py:169-503 (compute_center/disp/df/scale_sq/cov_factor), :1007-1039 (Gaussian),
// :1241-1258 (Student-t) and :1503-1506 (truncation Jacobian), collapsed as in SURVEY.md Appendix B.
#pragma once
#include "common.cuh"

#define LML_MAXR 16   // basis row + up to 15 coefficient curves per (l, Q) cell

// RHS staging for x-dependent Q (fused `coefficients`, gsum/helpers.py:98-100):
//   row 0              : basis (ones)
//   row 1 + q*n_c + m  : dy[x][m] / (ref[x] * Qx[q][x] ** orders[m])
// For the separable (scalar Q) path n_q_rows == 1 and Qx == nullptr: rows are dy[x][m] / ref[x].
__global__ void stage_rhs_kernel(double *__restrict__ dst, int64_t dst_ld, const double *__restrict__ dy,
                                 const double *__restrict__ ref, const double *__restrict__ Qx,
                                 const int32_t *__restrict__ orders, int64_t n, int n_c, int64_t n_q_rows) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (x >= n) return;
    if (row == 0) { dst[x] = 1.0; return; }
    const int64_t q = (row - 1) / n_c;
    const int m = (int)((row - 1) % n_c);
    double den = ref[x];
    if (Qx) den = den * pow(Qx[q * n + x], (double)orders[m]);
    dst[row * dst_ld + x] = dy[x * n_c + m] / den;
}

// Gram of the rows {0} U {1 + q*n_c .. 1 + (q+1)*n_c} of Wt for every (matrix b, block q):
//   G[b][q] is (R x R), R = n_c + 1, row-major, index 0 = basis row.
// One CTA per (q, b); fixed reduction order -> bit-reproducible regardless of how the grid is sharded.
template <int R>
__global__ void __launch_bounds__(256) gram_rows_kernel(const double *__restrict__ A, int64_t ld, int64_t bstride, int T,
                                                        int64_t n, double *__restrict__ G, int64_t n_q_rows) {
    __shared__ double red[R * (R + 1) / 2][8];
    const int64_t b = blockIdx.y, q = blockIdx.x;
    const double *W = A + b * bstride + (int64_t)T * GSUM_TILE * ld;
    const double *rows[R];
    rows[0] = W;
#pragma unroll
    for (int a = 1; a < R; a++) rows[a] = W + (1 + q * (R - 1) + (a - 1)) * ld;
    double acc[R * (R + 1) / 2];
#pragma unroll
    for (int p = 0; p < R * (R + 1) / 2; p++) acc[p] = 0.0;
    for (int64_t x = threadIdx.x; x < n; x += 256) {
        double v[R];
#pragma unroll
        for (int a = 0; a < R; a++) v[a] = rows[a][x];
        int p = 0;
#pragma unroll
        for (int a = 0; a < R; a++)
#pragma unroll
            for (int c = a; c < R; c++) acc[p++] += v[a] * v[c];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < R * (R + 1) / 2; p++) {
        double s = warp_sum(acc[p]);
        if (lane == 0) red[p][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < R * (R + 1) / 2) {
        double s = 0.0;
        for (int i = 0; i < 8; i++) s += red[threadIdx.x][i];
        // unpack p -> (a, c)
        int p = threadIdx.x, a = 0;
        while (p >= R - a) { p -= R - a; a++; }
        const int c = a + p;
        double *g = G + (b * n_q_rows + q) * R * R;
        g[a * R + c] = s;
        g[c * R + a] = s;
    }
}

// Generic fallback for R > 8: one pair per loop iteration (rows stay L1/L2 resident).
__global__ void __launch_bounds__(256) gram_rows_generic_kernel(const double *__restrict__ A, int64_t ld, int64_t bstride,
                                                                int T, int64_t n, double *__restrict__ G, int64_t n_q_rows, int R) {
    __shared__ double red[32];
    const int64_t b = blockIdx.y, q = blockIdx.x;
    const double *W = A + b * bstride + (int64_t)T * GSUM_TILE * ld;
    double *g = G + (b * n_q_rows + q) * R * R;
    for (int a = 0; a < R; a++)
        for (int c = a; c < R; c++) {
            const double *ra = a == 0 ? W : W + (1 + q * (R - 1) + (a - 1)) * ld;
            const double *rc = c == 0 ? W : W + (1 + q * (R - 1) + (c - 1)) * ld;
            double s = 0.0;
            for (int64_t x = threadIdx.x; x < n; x += 256) s += ra[x] * rc[x];
            s = block_sum(s, red);
            if (threadIdx.x == 0) { g[a * R + c] = s; g[c * R + a] = s; }
        }
}

struct LmlCellArgs {
    const double *G;          // (n_l, n_g, R, R): n_g = 1 (separable) or n_q (x-dependent Q)
    const double *logdet_part;  // (n_l, T)
    const int *info;          // (n_l)
    int T, R;
    int64_t n, n_l, n_q;
    int separable;
    const double *Q;          // (n_q) scalar ratios (separable only)
    const int32_t *orders;    // (n_c)
    const double *detf;       // (n_q) Jacobian term, may be null
    double center0, disp0, df0, scale0;
    int student;
    double *ll;               // (n_q, n_l)
    double *logdet_out;       // (n_l) or null
    // optional posterior outputs for a single cell (fit): [center, disp, df, scale_sq, cov_factor]
    double *post;
};

__device__ __forceinline__ double lml_lognorm(double df, double scale_sq, double disp) {
    // gsum/models.py:1241-1247
    double v = lgamma(0.5 * df) - 0.5 * df * log(0.5 * df * scale_sq);
    if (disp != 0.0) v += 0.5 * log(2.0 * M_PI * fabs(disp));
    return v;
}

// Cells of one chunk of length scales: local index l in [0, P.n_l) lands in column l0 + l of the (n_q, n_l_total) grid.
__global__ void lml_cell_chunk_kernel(LmlCellArgs P, int64_t l0, int64_t n_l_total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.n_l * P.n_q) return;
    const int64_t l = idx % P.n_l, q = idx / P.n_l;
    const int R = P.R, nc = R - 1;
    double logdet = 0.0;
    for (int k = 0; k < P.T; k++) logdet += P.logdet_part[l * P.T + k];
    if (P.logdet_out && q == 0) P.logdet_out[l] = P.info[l] ? nan("") : logdet;
    if (P.info[l] != 0) {                       // Cholesky failed -> -inf (gsum/models.py:970-972, 1212-1214)
        P.ll[q * n_l_total + l0 + l] = -INFINITY;
        return;
    }
    const double *G = P.G + (P.separable ? l : (l * P.n_q + q)) * R * R;
    // 1 / Q^order (gsum/helpers.py:98: `ratio ** orders`): integer orders, so the power is a short product (square and multiply,
    // a few ulp from pow(): far inside the 1e-10 tolerance) instead of six library pow() calls — the kernel is one thread per cell
    // and latency-bound (16.9 -> ~6 us per launch)
    double sc[LML_MAXR];
    sc[0] = 1.0;
    const double Qq = P.separable ? P.Q[q] : 1.0;
#pragma unroll 1
    for (int m = 0; m < nc; m++) {
        double v = 1.0;
        if (P.separable) {
            int e = P.orders[m];
            const bool neg = e < 0;
            e = neg ? -e : e;
            double b = Qq;
            v = 1.0;
            while (e) { if (e & 1) v *= b; b *= b; e >>= 1; }
            v = neg ? v : 1.0 / v;
        }
        sc[m + 1] = v;
    }
    // sufficient statistics: tr G_C, 1^T G_C 1, h^T 1, b
    double trG = 0.0, s11 = 0.0, hs = 0.0;
    for (int a = 1; a < R; a++) {
        trG += G[a * R + a] * sc[a] * sc[a];
        hs += G[a * R] * sc[a];
        double rowsum = 0.0;
        for (int c = 1; c < R; c++) rowsum += G[a * R + c] * sc[c];
        s11 += rowsum * sc[a];
    }
    const double bb = G[0];
    const double ncd = (double)nc, N = (double)P.n;
    const double yRy = s11 / (ncd * ncd);        // ybar^T R^-1 ybar
    const double BRy = hs / ncd;                 // B^T R^-1 ybar
    const double eta0 = P.center0, V0 = P.disp0, df0 = P.df0, tau0sq = P.scale0 * P.scale0;
    const double df = df0 + N * ncd;                                              // models.py:302
    double V = 0.0, eta = eta0;
    if (V0 != 0.0) {                                                              // models.py:269-270, 219-220
        V = 1.0 / (1.0 / V0 + ncd * bb);
        eta = V * (eta0 / V0 + ncd * BRy);
    }
    const double quad = trG - ncd * yRy;                                          // models.py:430-433
    const double aRa = yRy - 2.0 * eta0 * BRy + eta0 * eta0 * bb;
    const double BRa = BRy - bb * eta0;
    const double quad2 = ncd * (aRa - ncd * BRa * BRa * V);                       // models.py:435-445
    const double tausq = isinf(df0) ? tau0sq : (df0 * tau0sq + quad + quad2) / df;  // models.py:419-422, 447-448
    const double var = isinf(df) ? tausq : df * tausq / (df - 2.0);               // models.py:500-503
    double ll;
    if (!P.student) {                                                             // models.py:1007-1039
        const double Seta = quad + ncd * (yRy - 2.0 * eta * BRy + eta * eta * bb);
        ll = -0.5 * Seta / var - 0.5 * ncd * (N * log(var) + logdet) - 0.5 * ncd * N * log(2.0 * M_PI);
    } else {                                                                      // models.py:1241-1258
        ll = lml_lognorm(df, tausq, V) - lml_lognorm(df0, tau0sq, V0) - 0.5 * ncd * (N * log(2.0 * M_PI) + logdet);
    }
    if (P.detf) ll -= P.detf[q];                                                  // models.py:1503-1506
    P.ll[q * n_l_total + l0 + l] = ll;
    if (P.post) {
        P.post[0] = eta; P.post[1] = V; P.post[2] = df; P.post[3] = tausq; P.post[4] = var;
    }
}
