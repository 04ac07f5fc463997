// Gradient terms of the conjugate-GP likelihood (gsum/models.py:957-1056, eval_gradient=True, decomposition='cholesky').
//
// With  RHS = [B | y_1 .. y_nc]  (n x r),  Z = R^{-1} RHS  and the kernel-matrix derivatives  dR_p = dR / d log(theta_p),
// every quantity of the reference's gradient is a linear combination of
//     G   = RHS^T Z                    (r x r)        (quadratic forms  u^T R^{-1} v)
//     H_p = Z^T dR_p Z                 (r x r)        (the einsum('ji,jkp,ki->p', ...) terms of compute_scale_sq /
//                                                     compute_center and alpha^T dK alpha of models.py:1041-1056)
//     t_p = tr(R^{-1} dR_p)
// The device produces G, H_p, t_p and logdet R; the O(r^2) algebra that follows lives in the Python host next to the
// reference formulas it mirrors (gsum_b200/models.py::_lml_gradient).
//
// Hyper-parameters (log space, sklearn's):  p = 0: ConstantKernel value c;  p = 1 .. ls_dim: RBF length scale(s);
// p = ls_dim + 1: WhiteKernel noise level.   R = c * rbf + (noise + nugget) I  =>
//     dR_0 = c * rbf (diagonal c),   dR_{1+q} = c * rbf_ij * (x_iq/l_q - x_jq/l_q)^2  (summed over q when isotropic),
//     dR_last = noise * I.
#pragma once
#include "common.cuh"
#include "cov.cuh"

#define GRAD_MAXR 16

// Y_p[i][b] = sum_j dR_p(i,j) Z[j][b];  trow_p[i] = sum_j Rinv[j][i] dR_p(i,j).   grid (ceil(n/128), P, ceil(n/128)), block 128:
// blockIdx.z is a chunk of 128 columns j — every CTA writes the PARTIAL sums of its chunk (Y[jc][p][i][b], trow[jc][p][i])
// and grad_reduce_kernel adds the chunks in a fixed order (deterministic, and (n/128)^2 P CTAs instead of (n/128) P).
__global__ void __launch_bounds__(128) grad_rows_kernel(const double *__restrict__ XS, int64_t n, int d, int ls_dim, double constant,
                                                        double noise, const double *__restrict__ Zall, int64_t ldz, int r,
                                                        double *__restrict__ Y, double *__restrict__ trow) {
    extern __shared__ double sh[];                 // one chunk of 128 columns j: XS (128 x d) and Z (128 x r)
    double *xs_j = sh, *z_j = sh + 128 * COV_MAXD;
    const int p = blockIdx.y, P = ls_dim + 2;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const bool live = i < n;
    double xi[COV_MAXD];
    for (int q = 0; q < d; q++) xi[q] = live ? XS[i * d + q] : 0.0;
    double acc[GRAD_MAXR];
#pragma unroll
    for (int b = 0; b < GRAD_MAXR; b++) acc[b] = 0.0;
    double tacc = 0.0;
    const double *Rinv = Zall + r;                 // columns r .. r + n - 1 of [Z | R^{-1}]
    const int64_t jc = blockIdx.z, nP = gridDim.y;
    {
        const int64_t j0 = jc * 128;
        for (int e = threadIdx.x; e < 128 * d; e += 128) {
            const int64_t j = j0 + e / d;
            xs_j[(e / d) * COV_MAXD + e % d] = j < n ? XS[j * d + e % d] : 0.0;
        }
        for (int e = threadIdx.x; e < 128 * r; e += 128) {
            const int64_t j = j0 + e / r;
            z_j[(e / r) * GRAD_MAXR + e % r] = j < n ? Zall[j * ldz + e % r] : 0.0;
        }
        __syncthreads();
        const int jn = live ? (int)((n - j0 < 128) ? (n - j0) : 128) : 0;
        for (int jj = 0; jj < jn; jj++) {
            const int64_t j = j0 + jj;
            double w;
            if (p == P - 1) w = (i == j) ? noise : 0.0;
            else {
                double s = 0.0, sq = 0.0;
                for (int q = 0; q < d; q++) {
                    const double df = __dsub_rn(xi[q], xs_j[jj * COV_MAXD + q]);
                    const double d2 = __dmul_rn(df, df);
                    s = __dadd_rn(s, d2);
                    if (ls_dim == 1 || q == p - 1) sq += d2;
                }
                const double kv = (i == j) ? constant : constant * rbf_exp_neg(-0.5 * s);
                w = (p == 0) ? kv : kv * sq;
            }
            if (w != 0.0) {
                tacc = fma(Rinv[j * ldz + i], w, tacc);          // R^{-1} is symmetric: read row j, column i (coalesced over i)
#pragma unroll
                for (int b = 0; b < GRAD_MAXR; b++) if (b < r) acc[b] = fma(w, z_j[jj * GRAD_MAXR + b], acc[b]);
            }
        }
    }
    if (live) {
        for (int b = 0; b < r; b++) Y[(((int64_t)jc * nP + p) * n + i) * GRAD_MAXR + b] = acc[b];
        trow[((int64_t)jc * nP + p) * n + i] = tacc;
    }
}
// H_p[a][b] = sum_i Z[i][a] Y_p[i][b],  t_p = sum_i trow_p[i];  G[a][b] = sum_i RHS[i][a] Z[i][b]  (blockIdx.x = P handles G).
// One block per (p, a, b); Y_p[i][b] and trow_p[i] are the sums of their `nchunk` partials, added in chunk order: fixed
// summation order throughout.
__global__ void __launch_bounds__(256) grad_reduce_kernel(const double *__restrict__ Zall, int64_t ldz, const double *__restrict__ RHS,
                                                          const double *__restrict__ Y, const double *__restrict__ trow, int64_t n, int r,
                                                          int P, int nchunk, double *__restrict__ H, double *__restrict__ tr,
                                                          double *__restrict__ G) {
    __shared__ double red[32];
    const int p = blockIdx.x, ab = blockIdx.y, a = ab / r, b = ab % r;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        if (p < P) {
            double y = 0.0;
            for (int jc = 0; jc < nchunk; jc++) y += Y[(((int64_t)jc * P + p) * n + i) * GRAD_MAXR + b];
            s += Zall[i * ldz + a] * y;
        } else s += RHS[i * r + a] * Zall[i * ldz + b];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) { if (p < P) H[((int64_t)p * r + a) * r + b] = s; else G[a * r + b] = s; }
    if (p < P && ab == 0) {
        double t = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += 256) {
            double y = 0.0;
            for (int jc = 0; jc < nchunk; jc++) y += trow[((int64_t)jc * P + p) * n + i];
            t += y;
        }
        t = block_sum(t, red);
        if (threadIdx.x == 0) tr[p] = t;
    }
}
// B = [RHS | I]  (n x (r + n), row major)
__global__ void grad_stage_kernel(const double *__restrict__ RHS, int64_t n, int r, double *__restrict__ B) {
    const int64_t ld = r + n, i = blockIdx.y;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ld; c += (int64_t)gridDim.x * blockDim.x)
        B[i * ld + c] = c < r ? RHS[i * r + c] : (c - r == i ? 1.0 : 0.0);
}
