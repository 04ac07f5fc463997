// Small-N path of the (Q, l) likelihood grid: ONE CTA per length scale, the whole bordered matrix in shared memory.
//
// For the tutorial sizes (config C1: N = 50, C2: N = 200) a 64x64-tile schedule is all overhead: a 200 x 200 matrix is ten
// tile tasks of a 148-CTA cooperative launch, R makes a round trip through HBM, and the Gram needs a launch of its own.
// Here the CTA of a length scale
//   1. generates R(l) = c RBF_l(X) + (noise + nugget) I straight into shared memory (K1 fused: same operation order as
//      cov.cuh — X / l first, squared distance without FMA contraction, exp(-d / 2), diagonal exactly c + noise, + nugget),
//      as the lower triangle of 8x8 blocks (N = 200: 325 blocks = 163 KiB), with one more block row holding the right-hand
//      sides TRANSPOSED (basis, the n_c coefficient curves dy_n / ref: r <= 8 rows);
//   2. factors it right-looking by 8-column panels: warp 0 factors the 8x8 diagonal block in registers (every lane holds
//      the block: a chain rsqrt -> mul -> fma per column, results exchanged by warp-wide broadcast loads), one thread per row
//      below substitutes its eight entries, and all warps apply the rank-8 trailing update as FP64 tensor-core
//      DMMA.8x8x4 pairs block by block;
//   3. gets the forward solves for free: the border block row is swept by the same recurrences (it ends as
//      W^T = RHS^T L^{-T}), and its own diagonal block accumulates -W^T W — the Gram of the right-hand sides, which is all
//      the conjugate closed forms need (lml.cuh) — so neither the factor nor W ever leaves the SM.
// Outputs per length scale: the (r x r) Gram, the log-determinant, the status (LAPACK potrf convention).
#pragma once
#include "common.cuh"
#include "cov.cuh"

#define SN_MAXN 208                     // 26 block rows of 8
#define SN_MAXR 8                       // right-hand-side rows (basis + coefficient curves)
#define SN_THREADS 512
#define SN_MAXD 3                       // input dimensions of this path

struct SmallNArgs {
    const double *X;        // (n, d)
    const double *ls;       // (n_ls, ls_dim)
    const double *dy;       // (n, n_c) order-by-order corrections
    const double *ref;      // (n)
    int64_t n; int d, ls_dim, n_c;
    double constant, noise, nugget;
    double *G;              // (n_ls, r, r), r = n_c + 1
    double *logdet;         // (n_ls)
    int *info;              // (n_ls)
};

__host__ __device__ inline size_t smalln_smem_bytes(int64_t n) {
    const int nbk = (int)((n + 7) / 8);
    return sizeof(double) * ((size_t)(nbk + 1) * (nbk + 2) / 2 * 64 + (size_t)nbk * 8 * SN_MAXD + 64);
}

__global__ void __launch_bounds__(SN_THREADS, 1) smalln_lml_kernel(SmallNArgs P) {
    extern __shared__ __align__(16) double sn_smem[];
    __shared__ double red[32];
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int n = (int)P.n, nbk = (n + 7) / 8, d = P.d, r_rhs = P.n_c + 1;
    const int64_t l = blockIdx.x;
    double *S = sn_smem;                                          // block (bi, bj <= bi) at (bi (bi+1) / 2 + bj) * 64, row-major 8x8
    double *xs = S + (size_t)(nbk + 1) * (nbk + 2) / 2 * 64;      // scaled coordinates (nbk * 8, SN_MAXD)
    double *rsd = xs + (size_t)nbk * 8 * SN_MAXD;                 // 8 reciprocal pivots of the current panel
#define SN_BLK(bi, bj) (S + ((size_t)(bi) * ((bi) + 1) / 2 + (bj)) * 64)
    if (tid == 0) s_fail = 0;
    for (int e = tid; e < nbk * 8 * d; e += SN_THREADS) {
        const int i = e / d, q = e % d;
        xs[i * SN_MAXD + q] = i < n ? P.X[(int64_t)i * d + q] / P.ls[l * P.ls_dim + (P.ls_dim == 1 ? 0 : q)] : 0.0;
    }
    __syncthreads();
    // ---- 1. R(l) and the right-hand sides ------------------------------------------------------------------------------
    const double dval = __dadd_rn(__dadd_rn(P.constant, P.noise), P.nugget);
    const int nblocks = (nbk + 1) * (nbk + 2) / 2;
    for (int blk = w; blk < nblocks; blk += SN_THREADS / 32) {        // a warp per 8x8 block: the index is decoded once per block
        int bi = (int)((sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
        while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
        while (bi * (bi + 1) / 2 > blk) bi--;
        const int bj = blk - bi * (bi + 1) / 2;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = lane + 32 * h, r = e >> 3, c = e & 7;
            const int gi = bi * 8 + r, gj = bj * 8 + c;
            double v;
            if (bi < nbk) {
                if (gi >= n || gj >= n) v = gi == gj ? 1.0 : 0.0;
                else if (gi == gj) v = dval;
                else {
                    const double a = -0.5 * rbf_sqdist(xs + gi * SN_MAXD, xs + gj * SN_MAXD, d);
                    v = P.constant * rbf_exp_neg(a);
                }
            } else if (bj < nbk && r < r_rhs && gj < n) {
                v = r == 0 ? 1.0 : P.dy[(int64_t)gj * P.n_c + (r - 1)] / P.ref[gj];     // stage_rhs_kernel, separable path
            } else v = 0.0;
            S[(size_t)blk * 64 + e] = v;
        }
    }
    __syncthreads();
    // ---- 2. / 3. right-looking factorisation over 8-column panels; the border block row rides along -----------------------
    // Per panel p:  S (one thread per row below the diagonal block substitutes its eight entries) | B (rank-8 DMMA update of
    // every block (bi, bj), p < bj <= bi <= nbk) with one panel of look-ahead: warp 0 updates the next diagonal block first
    // and factors it (F) while warps 1-7 update the rest.
    auto factor_block = [&](int p) {
        // F: the 8x8 diagonal block in registers, every lane redundantly (potrf_lean_factor_block's arithmetic)
        double *blk = SN_BLK(p, p);
        double a[8][8];
#pragma unroll
        for (int m = 0; m < 8; m++)
#pragma unroll
            for (int q = 0; q <= m; q++) a[m][q] = blk[m * 8 + q];
        int fail = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double dd = a[j][j];
            if (!(dd > 0.0) && fail == 0) fail = p * 8 + j + 1;
            const double rs = rsqrt(dd);
            a[j][j] = dd * rs;
            if (lane == j) rsd[j] = rs;
#pragma unroll
            for (int m = j + 1; m < 8; m++) a[m][j] *= rs;
#pragma unroll
            for (int m = j + 1; m < 8; m++)
#pragma unroll
                for (int q = j + 1; q <= m; q++) a[m][q] = fma(-a[m][j], a[q][j], a[m][q]);
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 8; m++)
            if (lane == m) {
#pragma unroll
                for (int q = 0; q < 8; q++) blk[m * 8 + q] = q <= m ? a[m][q] : 0.0;
            }
        if (lane == 0 && fail && s_fail == 0) s_fail = fail;
    };
    auto update_block = [&](int p, int bi, int bj, double a0, double a1) {
        const double *Lb = SN_BLK(bj, p) + g * 8 + t;
        double *Cb = SN_BLK(bi, bj) + g * 8 + 2 * t;
        double2 cv = *reinterpret_cast<const double2 *>(Cb);
        dmma884(cv.x, cv.y, a0, Lb[0]);
        dmma884(cv.x, cv.y, a1, Lb[4]);
        *reinterpret_cast<double2 *>(Cb) = cv;
    };
    if (w == 0) factor_block(0);
    for (int p = 0; p < nbk; p++) {
        __syncthreads();
        {
            const int row = (p + 1) * 8 + tid;
            if (row < (nbk + 1) * 8) {
                double *x = SN_BLK(row >> 3, p) + (row & 7) * 8;
                const double *Lpp = SN_BLK(p, p);
                double v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = x[c];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double sacc = v[c];
#pragma unroll
                    for (int m = 0; m < c; m++) sacc = fma(-v[m], Lpp[c * 8 + m], sacc);
                    v[c] = sacc * rsd[c];
                }
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = v[c];
            }
        }
        __syncthreads();
        if (w == 0) {
            if (p + 1 < nbk) {
                const double *La = SN_BLK(p + 1, p) + g * 8 + t;
                update_block(p, p + 1, p + 1, -La[0], -La[4]);
                __syncwarp();
                factor_block(p + 1);
            } else {
                const double *La = SN_BLK(nbk, p) + g * 8 + t;          // last panel: only the Gram block is left
                update_block(p, nbk, nbk, -La[0], -La[4]);
            }
        } else {
            // block rows p+1 .. nbk dealt over warps 1-7; the A fragment of a row is loaded once, its blocks are independent
            for (int bi = p + 1 + (w - 1); bi <= nbk; bi += SN_THREADS / 32 - 1) {
                const double *La = SN_BLK(bi, p) + g * 8 + t;
                const double a0 = -La[0], a1 = -La[4];
                const int bj0 = (bi == p + 1) ? p + 2 : p + 1;           // (p+1, p+1) is warp 0's
                const int bj1 = (p + 1 == nbk) ? bi - 1 : bi;             // last panel: (nbk, nbk) is warp 0's
#pragma unroll 4
                for (int bj = bj0; bj <= bj1; bj++) update_block(p, bi, bj, a0, a1);
            }
        }
    }
    __syncthreads();
    // ---- outputs ----------------------------------------------------------------------------------------------------------
    const int fail = s_fail;
    double ld = 0.0;
    for (int j = tid; j < n; j += SN_THREADS) ld += log(SN_BLK(j >> 3, j >> 3)[(j & 7) * 9]);
    ld = block_sum(ld, red);
    if (tid == 0) {
        P.logdet[l] = fail ? nan("") : 2.0 * ld;                  // 2 sum log L_jj (gsum/models.py:1015,1250)
        P.info[l] = fail;
    }
    if (tid < r_rhs * r_rhs) {
        const int a = tid / r_rhs, b = tid % r_rhs;
        P.G[(l * r_rhs + a) * r_rhs + b] = fail ? nan("") : -SN_BLK(nbk, nbk)[a * 8 + b];
    }
#undef SN_BLK
}
