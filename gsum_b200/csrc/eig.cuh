// decomposition='eig' route: a symmetric FP64 eigensolver and the solves built on it.
//
//   reference call sites: scipy `eigh(R)` at gsum/models.py:714, 811, 974, 1166, 1216; numpy `eigh(cov)` at
//   gsum/diagnostics.py:63; `solve_sqrt(..., 'eig')` = Q diag(1/eig) Q^T y at models.py:480-484;
//   `eigen_errors` = solve(Q diag(sqrt(eig)), y - mean) at diagnostics.py:106-107.
//
// Eigensolver: one-sided (Hestenes) Jacobi on G = A V.  G and V are kept TRANSPOSED (row j = column j), so a plane
// rotation of columns (p, q) touches four contiguous rows; A is symmetric, so G starts as a plain copy of A.  One launch
// per round of the round-robin tournament: n/2 disjoint pairs, one CTA per pair, which
//   (1) reads rows g_p, g_q once for alpha = |g_p|^2, beta = |g_q|^2, gamma = g_p . g_q,
//   (2) skips the pair when |gamma| <= tol sqrt(alpha beta)  (tol = sqrt(n) eps, the LAPACK dgesvj criterion) or
//       |gamma| <= 0.1 eps |A|_F min(|g_p|, |g_q|)  (below the eps |A| accuracy any eigensolver delivers),
//   (3) otherwise rotates rows p, q of G and of V^T and counts the rotation.
// A sweep is n - 1 rounds; the iteration stops after a sweep without rotations.  On exit the rows of G are orthogonal:
// |lambda_j| = |g_j|, sign from v_j . g_j, eigenvector j = row j of V^T.  Accuracy is that of LAPACK's eigh: eigenvalues
// to eps |A| absolute, residuals |A v - lambda v| <= O(eps |A|).  (Scope: the covariance / correlation matrices of
// the reference's call sites, i.e. positive semi-definite up to rounding.  An indefinite matrix with a pair of eigenvalues
// +lambda, -lambda has a repeated SINGULAR value whose vectors the one-sided iteration cannot separate; gsum_eigh detects
// that through the Rayleigh quotients and reports it instead of returning wrong vectors.)
// Measured dead ends (profiles/r01_notes.md, session 5): de Rijk column ordering inside the rotation (more sweeps, not
// fewer, with the round-robin tournament); iterating on the Cholesky / pivoted-Cholesky factor of A (same 17-22 sweeps on
// RBF + noise matrices: the ~n-member cluster of eigenvalues at the noise level converges linearly either way, the
// rotation count falling by ~0.75 per sweep); a block variant (8 + 8 columns per CTA, 16 x 16 Gram diagonalised in shared
// memory, 1/8 of the rounds) — as many sweeps, and each round still streams all of G and V^T: 153 vs 99 ms at N = 1024.
// HBM/L2-bound: each round streams G and V^T once (4 n^2 x 8 B read+write when every pair rotates).
#pragma once
#include "common.cuh"

#define JAC_THREADS 256
#define JAC_MAX_SWEEPS 120

// Round-robin pairing (circle method) of np indices (np even): round r in [0, np-1), slot k in [0, np/2).
__device__ __forceinline__ void jacobi_pair(int np, int r, int k, int &p, int &q) {
    const int m = np - 1;
    if (k == 0) { p = m; q = r; }
    else { p = (r + k) % m; q = (r - k + m) % m; }
    if (p > q) { int t = p; p = q; q = t; }
}

// Pair of one CTA in one round.  order == nullptr: round-robin tournament over np indices (np / 2 CTAs, np - 1 rounds).
// order != nullptr: MODULUS ordering on SORTED positions — round s pairs the positions (i, j) with i + j = s (mod n), i < j
// (n CTAs, one per position i, half of them idle; n rounds), and order[] maps a position to its row, rows ranked by
// decreasing norm at the start of the sweep.  Emulated in numpy on RBF + 1e-4 I at N = 512 (profiles/r01_notes.md): strict
// relative criterion, iteration on the pivoted-Cholesky factor: round-robin 18 sweeps, modulus 16, modulus + sorted
// positions 13; iteration on A: 24 / 24 / 17.
__device__ __forceinline__ bool jacobi_select(int n, int np, int round, const int32_t *__restrict__ order, int &p, int &q) {
    if (order) {
        const int i = blockIdx.x, j = (round - i + n) % n;
        if (i >= j) return false;
        p = order[i]; q = order[j];
        return true;
    }
    jacobi_pair(np, round, blockIdx.x, p, q);
    return q < n;                                         // the padding index of an odd n sits this round out
}

// G <- A (n x n, leading dimension ld), Vt <- I
__global__ void __launch_bounds__(256) jacobi_init_kernel(const double *__restrict__ A, double *__restrict__ G, double *__restrict__ Vt,
                                                          int n, int64_t ld) {
    const int64_t r = blockIdx.x;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        G[r * ld + c] = A[r * (int64_t)n + c];
        Vt[r * ld + c] = (c == r) ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(JAC_THREADS) jacobi_round_kernel(double *__restrict__ G, double *__restrict__ Vt, int n, int64_t ld, int np,
                                                                   int round, double tol, double tol_abs, double tol_gamma,
                                                                   unsigned int *__restrict__ rotations, const int32_t *__restrict__ order) {
    __shared__ double red[3][JAC_THREADS / 32];
    int p, q;
    if (!jacobi_select(n, np, round, order, p, q)) return;
    double *gp = G + p * ld, *gq = G + q * ld;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        const double x = gp[i], y = gq[i];
        a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
    }
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[0][w] = a; red[1][w] = b; red[2][w] = g; }
    __syncthreads();
    a = b = g = 0.0;
#pragma unroll
    for (int i = 0; i < JAC_THREADS / 32; i++) { a += red[0][i]; b += red[1][i]; g += red[2][i]; }
    // rotate when the pair is non-orthogonal relatively (|gamma| > tol |g_p| |g_q|, the dgesvj criterion) and, if an
    // absolute tolerance is set, also on the scale of eps |A| (|gamma| > tol_abs min(|g_p|, |g_q|): ignoring gamma moves a
    // singular value by gamma / (2 sigma) at most).  Measured on RBF + noise matrices (N = 1024): tol_abs = eps |A|_F saves a
    // quarter of the sweeps (24 -> 18) but triples the residual of R^-1 y; 0.1 eps |A|_F (the default) takes 19 sweeps
    // with the residual of the purely relative criterion (3e-10 vs 6e-10).
    // The negated form also catches a zero column and NaN.
    if (!(fabs(g) > tol * sqrt(a) * sqrt(b)) || !(fabs(g) > tol_abs * sqrt(fmin(a, b))) || !(fabs(g) > tol_gamma)) return;
    if (threadIdx.x == 0) atomicAdd(rotations, 1u);
    const double zeta = (b - a) / (2.0 * g);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
    if (!Vt) {                                            // factor mode: the eigenvectors are the normalised rows of G
        for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
            const double x = gp[i], y = gq[i];
            gp[i] = cs * x - sn * y; gq[i] = sn * x + cs * y;
        }
        return;
    }
    double *vp = Vt + p * ld, *vq = Vt + q * ld;
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        const double x = gp[i], y = gq[i];
        gp[i] = cs * x - sn * y; gq[i] = sn * x + cs * y;
        const double u = vp[i], v = vq[i];
        vp[i] = cs * u - sn * v; vq[i] = sn * u + cs * v;
    }
}

// Register-resident round for n <= 256 * EPT (used for n <= 1024, see gsum_eigh): every thread loads its EPT elements of the four rows up front (all loads
// in flight at once, one memory round trip instead of two dependent passes), the CTA reduces alpha / beta / gamma, and the
// rotation is applied from registers.  Rows of pairs that do not rotate are only read.  The V^T rows are fetched together
// with the G rows — they are needed unless the pair is skipped, and the early fetch hides their latency behind the
// reduction (skipped pairs pay 2x the read, which the L2 absorbs: late sweeps are launch-bound, not bandwidth-bound).
template <int EPT>
__global__ void __launch_bounds__(JAC_THREADS) jacobi_round_reg_kernel(double *__restrict__ G, double *__restrict__ Vt, int n, int64_t ld, int np,
                                                                       int round, double tol, double tol_abs, double tol_gamma,
                                                                       unsigned int *__restrict__ rotations, const int32_t *__restrict__ order) {
    __shared__ double red[3][JAC_THREADS / 32];
    int p, q;
    if (!jacobi_select(n, np, round, order, p, q)) return;
    double *gp = G + p * ld, *gq = G + q * ld;
    double x[EPT], y[EPT], u[EPT], v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        const int i = threadIdx.x + e * JAC_THREADS;
        x[e] = i < n ? gp[i] : 0.0;
        y[e] = i < n ? gq[i] : 0.0;
    }
    double *vp = nullptr, *vq = nullptr;
    if (Vt) {
        vp = Vt + p * ld; vq = Vt + q * ld;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int i = threadIdx.x + e * JAC_THREADS;
            u[e] = i < n ? vp[i] : 0.0;
            v[e] = i < n ? vq[i] : 0.0;
        }
    }
    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; e++) { a = fma(x[e], x[e], a); b = fma(y[e], y[e], b); g = fma(x[e], y[e], g); }
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[0][w] = a; red[1][w] = b; red[2][w] = g; }
    __syncthreads();
    a = b = g = 0.0;
#pragma unroll
    for (int i = 0; i < JAC_THREADS / 32; i++) { a += red[0][i]; b += red[1][i]; g += red[2][i]; }
    if (!(fabs(g) > tol * sqrt(a) * sqrt(b)) || !(fabs(g) > tol_abs * sqrt(fmin(a, b))) || !(fabs(g) > tol_gamma)) return;
    if (threadIdx.x == 0) atomicAdd(rotations, 1u);
    const double zeta = (b - a) / (2.0 * g);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        const int i = threadIdx.x + e * JAC_THREADS;
        if (i < n) {
            gp[i] = cs * x[e] - sn * y[e]; gq[i] = sn * x[e] + cs * y[e];
            if (Vt) { vp[i] = cs * u[e] - sn * v[e]; vq[i] = sn * u[e] + cs * v[e]; }
        }
    }
}

// w[j] = sign(v_j . g_j) |g_j|;  rq[j] = v_j . g_j (the Rayleigh quotient: equals w[j] when v_j is an eigenvector);  flip[j] = -1 when the largest-magnitude component of v_j is negative (the eigenvector is
// returned with that component positive), else +1.  One CTA per row.
__global__ void __launch_bounds__(JAC_THREADS) jacobi_finish_kernel(const double *__restrict__ G, const double *__restrict__ Vt, int n, int64_t ld,
                                                                    double *__restrict__ w, double *__restrict__ flip, double *__restrict__ rq) {
    __shared__ double red[4][JAC_THREADS / 32];
    const int64_t j = blockIdx.x;
    const double *g = G + j * ld, *v = Vt + j * ld;
    double nn = 0.0, dot = 0.0, big = -1.0, bigv = 0.0;
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        const double x = g[i], y = v[i];
        nn = fma(x, x, nn); dot = fma(x, y, dot);
        if (fabs(y) > big) { big = fabs(y); bigv = y; }
    }
    nn = warp_sum(nn); dot = warp_sum(dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, big, o), ov = __shfl_xor_sync(0xffffffffu, bigv, o);
        if (ob > big) { big = ob; bigv = ov; }
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { red[0][wp] = nn; red[1][wp] = dot; red[2][wp] = big; red[3][wp] = bigv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        nn = dot = 0.0; big = -1.0; bigv = 0.0;
        for (int i = 0; i < JAC_THREADS / 32; i++) {
            nn += red[0][i]; dot += red[1][i];
            if (red[2][i] > big) { big = red[2][i]; bigv = red[3][i]; }
        }
        w[j] = copysign(sqrt(nn), dot);
        rq[j] = dot;
        flip[j] = bigv < 0.0 ? -1.0 : 1.0;
    }
}

// V[i][k] = flip[perm[k]] * Vt[perm[k]][i]      (eigenvectors as COLUMNS of a row-major (n, n) array, LAPACK order)
__global__ void __launch_bounds__(256) jacobi_gather_kernel(const double *__restrict__ Vt, int64_t ld, const int32_t *__restrict__ perm,
                                                            const double *__restrict__ flip, int n, double *__restrict__ V) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, i = i0 + tx;
        double v = 0.0;
        if (k < n && i < n) { const int src = perm[k]; v = flip[src] * Vt[src * ld + i]; }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, k = k0 + tx;
        if (i < n && k < n) V[(int64_t)i * n + k] = tile[tx][r];
    }
}

// ---- C (M x N) = op(A) (M x K) . diag(1/bdiv) (B (K x N) - bsub[k]) with an optional row scaling --------------------------------
// Row-major operands; op(A) = A (lda >= K) or A^T (A stored K x M, lda >= M).  64 x 64 tile per CTA, four warps of
// 32 x 32, FP64 tensor-core MMA (DMMA 8x8x4) on fragments read from padded shared memory (stride 36: conflict-free).
// row_mode: 0 none, 1 C[m][:] /= rs[m], 2 C[m][:] /= sqrt(|rs[m]|).
// History: the single-buffered version ran at 16-19 TFLOP/s (0.46-0.54 of the DGEMM peak); a register-staged prefetch of the
// next K slab (16 + 16 doubles per thread) pushed it to 255 registers with spills and HALVED the rate (M = N = 4096,
// K = 1024: 1.92 -> 4.34 ms); the shipped kernel double-buffers through cp.async instead.
struct EigGemmArgs {
    const double *A; int64_t lda; int transA;
    const double *B; int64_t ldb;
    double *C; int64_t ldc;
    int64_t M, N, K;
    const double *bsub;                 // B[k][:] - bsub[k]
    const double *bdiv;                 // (B[k][:] - bsub[k]) / bdiv[k]
    const double *rs; int row_mode;
};

#define EG_LDN 68                                   // row stride of the [k][64] tiles: 68 % 16 == 4 -> conflict-free fragments
#define EG_A_ELEMS (64 * GSUM_LDH)                  // A tile: [64 m][36] (op(A) = A) or [32 k][68] (op(A) = A^T); 2304 >= 2176
#define EG_B_ELEMS (GSUM_KH * EG_LDN)
#define EG_STAGE (EG_A_ELEMS + EG_B_ELEMS + 2 * GSUM_KH)
#define EG_SMEM_BYTES (2 * EG_STAGE * sizeof(double))

// 8-byte cp.async with zero fill (src_bytes = 0 copies nothing and writes zeros)
__device__ __forceinline__ void cp_async8_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem_src), "r"(src_bytes));
}

// Two-stage cp.async pipeline: K slab k + 1 streams into shared memory while the DMMA loop runs on slab k.  Tiles keep the
// global layout's contiguous direction ([m][k] for A, [k][m] for A^T, [k][n] for B), so every copy is a plain 8-byte
// cp.async; the per-k transforms of B (subtract bsub[k], divide by bdiv[k]) are applied to the fragments from two small
// per-slab vectors.  (The single-buffered first version ran the DMMA pipe at 45-54 %, profiles/r01_ncu_eig.txt.)
__global__ void __launch_bounds__(128) eig_gemm_kernel(EigGemmArgs P) {
    extern __shared__ __align__(16) double eg_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int stage, int64_t k0) {
        double *As = eg_smem + (size_t)stage * EG_STAGE, *Bs = As + EG_A_ELEMS, *sub = Bs + EG_B_ELEMS, *inv = sub + GSUM_KH;
#pragma unroll 4
        for (int e = tid; e < 64 * GSUM_KH; e += 128) {
            if (P.transA) {
                const int m = e & 63, k = e >> 6;
                const int64_t gm = m0 + m, gk = k0 + k;
                const bool ok = gm < P.M && gk < P.K;
                cp_async8_zfill(As + k * EG_LDN + m, ok ? P.A + gk * P.lda + gm : P.A, ok ? 8 : 0);
            } else {
                const int k = e & (GSUM_KH - 1), m = e / GSUM_KH;
                const int64_t gm = m0 + m, gk = k0 + k;
                const bool ok = gm < P.M && gk < P.K;
                cp_async8_zfill(As + m * GSUM_LDH + k, ok ? P.A + gm * P.lda + gk : P.A, ok ? 8 : 0);
            }
            const int n = e & 63, k = e >> 6;
            const int64_t gn = n0 + n, gk = k0 + k;
            const bool ok = gn < P.N && gk < P.K;
            cp_async8_zfill(Bs + k * EG_LDN + n, ok ? P.B + gk * P.ldb + gn : P.B, ok ? 8 : 0);
        }
        if (tid < GSUM_KH) {
            const int64_t gk = k0 + tid;
            const bool ok = gk < P.K;
            sub[tid] = (ok && P.bsub) ? P.bsub[gk] : 0.0;
            inv[tid] = (ok && P.bdiv) ? 1.0 / P.bdiv[gk] : 1.0;
        }
        cp_async_commit();
    };

    const int64_t nk = (P.K + GSUM_KH - 1) / GSUM_KH;
    issue(0, 0);
    for (int64_t it = 0; it < nk; it++) {
        if (it + 1 < nk) { issue((int)((it + 1) & 1), (it + 1) * GSUM_KH); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double *As = eg_smem + (size_t)(it & 1) * EG_STAGE, *Bs = As + EG_A_ELEMS, *sub = Bs + EG_B_ELEMS, *inv = sub + GSUM_KH;
#pragma unroll
        for (int kk = 0; kk < GSUM_KH; kk += 4) {
            double a[4], b[4];
            if (P.transA) {
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[(kk + t) * EG_LDN + wm + 8 * i + g];
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[(wm + 8 * i + g) * GSUM_LDH + kk + t];
            }
            const double sb = sub[kk + t], iv = inv[kk + t];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = (Bs[(kk + t) * EG_LDN + wn + 8 * j + g] - sb) * iv;
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t gm = m0 + wm + 8 * i + g;
        if (gm >= P.M) continue;
        double s = 1.0;
        if (P.row_mode == 1) s = 1.0 / P.rs[gm];
        else if (P.row_mode == 2) s = 1.0 / sqrt(fabs(P.rs[gm]));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t gn = n0 + wn + 8 * j + 2 * t;
            if (gn < P.N) P.C[gm * P.ldc + gn] = s * acc[i][j][0];
            if (gn + 1 < P.N) P.C[gm * P.ldc + gn + 1] = s * acc[i][j][1];
        }
    }
}

// out[j] = sum_k U[k][j]^2 / w[k]      (diag of U^T diag(1/w) U; U is (n x m) row-major: coalesced over j)
// 32 columns per CTA, 32 slices of the k range per column, fixed-order reduction through shared memory (deterministic).
__global__ void __launch_bounds__(1024) eig_colquad_kernel(const double *__restrict__ U, int64_t n, int64_t m, const double *__restrict__ w,
                                                           double *__restrict__ out) {
    __shared__ double part[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * 32 + tx;
    double s = 0.0;
    if (j < m)
        for (int64_t k = ty; k < n; k += 32) { const double u = U[k * m + j]; s = fma(u * (1.0 / w[k]), u, s); }
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < m) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 32; r++) t += part[r][tx];
        out[j] = t;
    }
}

// ---- factor mode ------------------------------------------------------------------------------------------------------
// A numerically positive definite: A = F F^T with F = P L_p from the pivoted-Cholesky kernel (gsum_pivoted_cholesky's
// G_out).  The iteration runs on the columns of F (rows of G = F^T): it diagonalises F^T F = L_p^T L_p, which diagonal
// pivoting makes strongly graded and diagonally dominant, and needs no V — with F V = U Sigma, A = U Sigma^2 U^T, so
// lambda_j = |g_j|^2 and eigenvector j = g_j / |g_j|.  Here gamma is already on the scale of the eigenvalues: pairs with
// |gamma| <= tol_gamma = c eps |A|_F are skipped, and one Newton-Schulz step restores orthogonality afterwards.
// Measured (RBF(0.05) + 1e-4 I; profiles/r01_eig_probe.txt): with c = 0.01, 12-14 sweeps instead of 19-20 and no V^T to
// rotate — N = 1024 / 2048 / 4096 in 55-67 / 197 / 1810 ms against 89 / 385 / 3070 ms for the default mode and
// 61 / 290 / 2330 ms for LAPACK on the host — but the skipped in-cluster couplings add up over the ~n^2 pairs of the
// noise-level cluster: the residual of R^-1 y is 2e-9 / 7e-9 / 1.4e-8 where both the default mode and LAPACK's own
// Q diag(1/eig) Q^T y give 3e-10 / 4e-10 / 9e-10 (c = 0.1: 2e-8 ... 1e-7 at the same sweep count; c <= 0.001: 21-28
// sweeps).  Parity comes first, so this mode is OPT-IN (GSUM_B200_EIGH_FACTOR=1).
__global__ void __launch_bounds__(256) jacobi_init_factor_kernel(const double *__restrict__ F, int64_t ldf, double *__restrict__ G, int n, int64_t ld) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.y * 32, i0 = blockIdx.x * 32;           // G[j][i] = F[i][j]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        tile[r][tx] = (i < n && j < n) ? F[(int64_t)i * ldf + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, i = i0 + tx;
        if (j < n && i < n) G[(int64_t)j * ld + i] = tile[tx][r];
    }
}

// factor mode: w[j] = |g_j|^2, flip[j] = +-1 / |g_j| (largest-magnitude component of the eigenvector positive)
__global__ void __launch_bounds__(JAC_THREADS) jacobi_finish_factor_kernel(const double *__restrict__ G, int n, int64_t ld,
                                                                           double *__restrict__ w, double *__restrict__ flip) {
    __shared__ double red[3][JAC_THREADS / 32];
    const int64_t j = blockIdx.x;
    const double *g = G + j * ld;
    double nn = 0.0, big = -1.0, bigv = 0.0;
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        const double x = g[i];
        nn = fma(x, x, nn);
        if (fabs(x) > big) { big = fabs(x); bigv = x; }
    }
    nn = warp_sum(nn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, big, o), ov = __shfl_xor_sync(0xffffffffu, bigv, o);
        if (ob > big) { big = ob; bigv = ov; }
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { red[0][wp] = nn; red[1][wp] = big; red[2][wp] = bigv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        nn = 0.0; big = -1.0; bigv = 0.0;
        for (int i = 0; i < JAC_THREADS / 32; i++) {
            nn += red[0][i];
            if (red[1][i] > big) { big = red[1][i]; bigv = red[2][i]; }
        }
        w[j] = nn;
        flip[j] = (bigv < 0.0 ? -1.0 : 1.0) / sqrt(nn);
    }
}

// T <- 1.5 I - 0.5 S  (in place): the Newton-Schulz step U <- U (3 I - U^T U) / 2 that restores the orthogonality of the
// factor-mode eigenvectors from the 1e-10 the eigenvalue-scale criterion leaves to 1e-20 (quadratic), at two GEMMs.
__global__ void __launch_bounds__(256) newton_schulz_T_kernel(double *__restrict__ S, int64_t n) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) S[i * n + j] = (i == j ? 1.5 : 0.0) - 0.5 * S[i * n + j];
}
