// K7/K8: pivoted Cholesky (LAPACK dpstrf semantics), batched posterior draws  m + L z  and fused
// credible-interval coverage counting.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "chol.cuh"

// ------------------------------------------------------------------------------------------------
// Pivoted Cholesky  (gsum/helpers.py:185-199 -> LAPACK dpstrf, lower, NB = 64)
//
// Storage: the FULL symmetric matrix Af (np x np, both triangles kept up to date) and a separate factor
// buffer Lb (np x np, zero initialised), both indexed by ORIGINAL ("physical") row.  A symmetric
// permutation then moves no data at all: LAPACK's row/column interchanges become a swap of two entries
// of piv[] (logical position -> physical row).  Column j of the factor for physical row p is
//     Lb[p][j] = (Af[pj][p] - sum_{m=k}^{j-1} Lb[p][m] Lb[pj][m]) / sqrt(ajj)       (dgemv + dscal of dpstrf)
// with pj the pivot's physical row — a coalesced read of one matrix row — and after every block of 64
// columns the trailing matrix gets the rank-64 DMMA update  Af -= Lb[:,k:k+64] Lb[:,k:k+64]^T  (dsyrk).
// The running diagonal follows dpstrf's arithmetic order exactly: `work` is reset at each block start and
// accumulates the squares of the current block's columns only; pivot = first maximum (in logical order) of
// diag - work; stop when that maximum <= n * eps * (largest initial diagonal entry).
// Output: G = Lb (rows in original order: M = G G^T, the array gsum's pivoted_cholesky returns) and
// Lp[i] = Lb[piv[i]] (LAPACK's lower factor of P^T M P).
// ------------------------------------------------------------------------------------------------
#define PSTRF_ROWS 128               // physical rows per CTA of the panel kernel (one row per thread)
#define PSTRF_NB 64

struct PstrfState {          // device-resident scalars
    int rank;                // columns completed
    int info;                // 0 ok, 1 = stopped early (matrix not positive definite to working precision)
    double dstop;
};
struct PstrfSlot { double v; int i; int pad; };       // one CTA's pivot candidate: value and LOGICAL position

__global__ void pstrf_init_kernel(int32_t *piv, int32_t *pos, PstrfState *st, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { piv[i] = i; pos[i] = i; }
    if (i == 0) { st->rank = 0; st->info = 0; st->dstop = 0.0; }
}

#define PSTRF_BETTER(v, i, bv, bi) ((v) > (bv) || ((v) == (bv) && (i) < (bi)))   // first maximum in logical order: a total order
__device__ __forceinline__ void pstrf_block_argmax(double &bv, int &bi, double *redv, int *redi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (PSTRF_BETTER(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { redv[w] = bv; redi[w] = bi; }
    __syncthreads();
    bv = redv[0]; bi = redi[0];
#pragma unroll
    for (int x = 1; x < PSTRF_ROWS / 32; x++) if (PSTRF_BETTER(redv[x], redi[x], bv, bi)) { bv = redv[x]; bi = redi[x]; }
    __syncthreads();
}

// One block of up to 64 columns starting at logical column k: a COOPERATIVE launch of ceil(n / 128) CTAs, each owning 128
// physical rows (one per thread: running diagonal, dpstrf's work entry, the logical position and the most recent factor
// entry live in registers; the row's entries of the current block sit in shared memory for the per-column dgemv).
// Per column: every CTA publishes its best pivot candidate, ONE grid barrier, every CTA reduces the same candidate list
// to the same pivot (the order is total, so the result does not depend on the reduction tree), swaps its private copy
// of piv[], fetches the pivot row's block entries from Pt (global, written before the barrier) and updates its rows.
// The arithmetic per row is sequential in the block column index, exactly as in the single-CTA version this replaces.
__global__ void __launch_bounds__(PSTRF_ROWS) pstrf_panel_kernel(const double *__restrict__ Af, double *__restrict__ Lb,
                                                                 double *Pt, int64_t ld, int n, int k, int32_t *piv, int32_t *posg,
                                                                 PstrfState *st, PstrfSlot *slots) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sm[];
    double *ptl = sm;                                  // [64][128]: this CTA's rows of the block's factor columns, transposed
    int *pivs = (int *)(sm + PSTRF_NB * PSTRF_ROWS);   // [n]: private copy of logical position -> physical row
    __shared__ double redv[PSTRF_ROWS / 32], lrow[PSTRF_NB];
    __shared__ int redi[PSTRF_ROWS / 32];
    const int tid = threadIdx.x, p = blockIdx.x * PSTRF_ROWS + tid, ncta = gridDim.x;
    const bool live = p < n;
    if (st->info != 0) return;                         // an earlier block already stopped (uniform over the grid)
    const int jb = min(PSTRF_NB, n - k);
    for (int i = tid; i < n; i += PSTRF_ROWS) pivs[i] = piv[i];
    double work = 0.0, col = 0.0, dstop = st->dstop;
    const double dg = live ? Af[(int64_t)p * ld + p] : 0.0;
    int pos = live ? posg[p] : -1;
    __syncthreads();
    for (int j = k; j < k + jb; j++) {
        // ---- running diagonal + pivot candidate: first maximum in LOGICAL order among positions >= j -------------
        double bv = -INFINITY; int bi = 0x7fffffff;
        if (live && pos >= j) {
            if (j > k) work += col * col;
            bv = dg - work; bi = pos;
        }
        pstrf_block_argmax(bv, bi, redv, redi);
        PstrfSlot *sl = slots + (j & 1) * ncta;
        if (tid == 0) { PstrfSlot s; s.v = bv; s.i = bi; s.pad = 0; sl[blockIdx.x] = s; }
        grid.sync();
        bv = -INFINITY; bi = 0x7fffffff;
        for (int x = tid; x < ncta; x += PSTRF_ROWS) {
            const double ov = __ldcg(&sl[x].v);
            const int oi = __ldcg(&sl[x].i);
            if (PSTRF_BETTER(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        pstrf_block_argmax(bv, bi, redv, redi);
        if (j == 0) dstop = (double)n * 1.1102230246251565e-16 * bv;              // N * DLAMCH('Epsilon') * max diag
        const bool bad = (j == 0) ? !(bv > 0.0) : !(bv > dstop);
        if (bad || bi == 0x7fffffff) {                                              // uniform: every CTA sees the same pivot
            if (blockIdx.x == 0 && tid == 0) { st->info = 1; st->rank = j; if (j == 0) st->dstop = dstop; }
            if (blockIdx.x == 0) for (int i = tid; i < n; i += PSTRF_ROWS) piv[i] = pivs[i];
            if (live) posg[p] = pos;
            return;
        }
        const int pj = pivs[bi], pold = pivs[j];                                    // interchange logical positions j and pvt
        __syncthreads();
        if (tid == 0) { pivs[bi] = pold; pivs[j] = pj; }
        if (live) { if (p == pold) pos = bi; if (p == pj) pos = j; }
        const double ajj = sqrt(bv), inv = 1.0 / ajj;
        const int nprev = j - k;
        if (tid < nprev) lrow[tid] = __ldcg(Pt + (int64_t)tid * ld + pj);
        __syncthreads();
        // ---- column j (dgemv with the block's previous columns, then dscal by 1/ajj) ---------------------------
        if (live && pos >= j) {                 // rows already pivoted keep an exact 0 in column j
            double v;
            if (pos == j) v = ajj;
            else {
                // the trailing update maintains the LOWER triangle only (half the rank-64 update's flops and bytes): entry (pj, p) of the
                // symmetric matrix is read where it is kept — coalesced along row pj for the rows above it, one strided element otherwise
                v = pj >= p ? Af[(int64_t)pj * ld + p] : Af[(int64_t)p * ld + pj];
#pragma unroll 8
                for (int m = 0; m < nprev; m++) v = fma(-ptl[m * PSTRF_ROWS + tid], lrow[m], v);
                v *= inv;
            }
            ptl[nprev * PSTRF_ROWS + tid] = v;
            Pt[(int64_t)nprev * ld + p] = v;
            Lb[(int64_t)p * ld + j] = v;
            col = (pos == j) ? 0.0 : v;
        }
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int i = tid; i < n; i += PSTRF_ROWS) piv[i] = pivs[i];
        if (tid == 0) { st->rank = k + jb; if (k == 0) st->dstop = dstop; }
    }
    if (live) posg[p] = pos;
}

// The same panel for n <= 4096 as ONE thread-block cluster of up to 16 CTAs x 256 rows: the per-column exchange of the pivot
// candidates goes through distributed shared memory (every CTA writes its candidate into every peer's slot array) and the
// per-column barrier is the cluster's hardware barrier (~0.2 us) instead of a grid barrier through global memory (~2 us):
// the panel is latency-bound by construction — the pivot of column j+1 needs column j — so the barrier IS the kernel.
// Arithmetic, reduction order (a total order on (value, logical position)) and therefore the pivots are those of the kernel above.
#define PSTRF_CROWS 256
#define PSTRF_CMAX 16
__device__ __forceinline__ void pstrf_block_argmax_c(double &bv, int &bi, double *redv, int *redi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (PSTRF_BETTER(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { redv[w] = bv; redi[w] = bi; }
    __syncthreads();
    bv = redv[0]; bi = redi[0];
#pragma unroll
    for (int x = 1; x < PSTRF_CROWS / 32; x++) if (PSTRF_BETTER(redv[x], redi[x], bv, bi)) { bv = redv[x]; bi = redi[x]; }
    __syncthreads();
}
__global__ void __launch_bounds__(PSTRF_CROWS) pstrf_panel_cluster_kernel(const double *__restrict__ Af, double *__restrict__ Lb,
                                                                          double *Pt, int64_t ld, int n, int k, int32_t *piv, int32_t *posg,
                                                                          PstrfState *st) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    double *ptl = sm;                                  // [64][256]: this CTA's rows of the block's factor columns, transposed
    int *pivs = (int *)(sm + PSTRF_NB * PSTRF_CROWS);  // [n]: private copy of logical position -> physical row
    __shared__ double redv[PSTRF_CROWS / 32], lrow[PSTRF_NB];
    __shared__ int redi[PSTRF_CROWS / 32];
    __shared__ double slot_v[2][PSTRF_CMAX];
    __shared__ int slot_i[2][PSTRF_CMAX];
    const int tid = threadIdx.x, p = blockIdx.x * PSTRF_CROWS + tid, ncta = gridDim.x;
    const unsigned rank = cluster.block_rank();
    const bool live = p < n;
    if (st->info != 0) return;                         // an earlier block already stopped (uniform over the cluster)
    const int jb = min(PSTRF_NB, n - k);
    for (int i = tid; i < n; i += PSTRF_CROWS) pivs[i] = piv[i];
    double work = 0.0, col = 0.0, dstop = st->dstop;
    const double dg = live ? Af[(int64_t)p * ld + p] : 0.0;
    int pos = live ? posg[p] : -1;
    __syncthreads();
    cluster.sync();                                    // every CTA of the cluster is resident before the first remote store
    for (int j = k; j < k + jb; j++) {
        double bv = -INFINITY; int bi = 0x7fffffff;
        if (live && pos >= j) {
            if (j > k) work += col * col;
            bv = dg - work; bi = pos;
        }
        pstrf_block_argmax_c(bv, bi, redv, redi);
        if (tid < ncta) {                              // this CTA's candidate into slot [rank] of CTA tid
            *cluster.map_shared_rank(&slot_v[j & 1][rank], tid) = bv;
            *cluster.map_shared_rank(&slot_i[j & 1][rank], tid) = bi;
        }
        cluster.sync();                                // the one barrier of the column (release / acquire at cluster scope: also Pt)
        bv = -INFINITY; bi = 0x7fffffff;
        if (tid < ncta) { bv = slot_v[j & 1][tid]; bi = slot_i[j & 1][tid]; }
        pstrf_block_argmax_c(bv, bi, redv, redi);
        if (j == 0) dstop = (double)n * 1.1102230246251565e-16 * bv;              // N * DLAMCH('Epsilon') * max diag
        const bool bad = (j == 0) ? !(bv > 0.0) : !(bv > dstop);
        if (bad || bi == 0x7fffffff) {                                              // uniform: every CTA sees the same pivot
            if (blockIdx.x == 0 && tid == 0) { st->info = 1; st->rank = j; if (j == 0) st->dstop = dstop; }
            if (blockIdx.x == 0) for (int i = tid; i < n; i += PSTRF_CROWS) piv[i] = pivs[i];
            if (live) posg[p] = pos;
            cluster.sync();                            // nobody leaves while a peer may still write into its slots
            return;
        }
        const int pj = pivs[bi], pold = pivs[j];                                    // interchange logical positions j and pvt
        __syncthreads();
        if (tid == 0) { pivs[bi] = pold; pivs[j] = pj; }
        if (live) { if (p == pold) pos = bi; if (p == pj) pos = j; }
        const double ajj = sqrt(bv), inv = 1.0 / ajj;
        const int nprev = j - k;
        // this row's entry of the pivot column: issued before the pivot row's block entries are fetched, so the two L2 round trips overlap
        double a_in = 0.0;
        if (live && pos > j) a_in = pj >= p ? Af[(int64_t)pj * ld + p] : Af[(int64_t)p * ld + pj];
        if (tid < nprev) lrow[tid] = __ldcg(Pt + (int64_t)tid * ld + pj);
        __syncthreads();
        if (live && pos >= j) {                 // rows already pivoted keep an exact 0 in column j
            double v;
            if (pos == j) v = ajj;
            else {
                v = a_in;
#pragma unroll 8
                for (int m = 0; m < nprev; m++) v = fma(-ptl[m * PSTRF_CROWS + tid], lrow[m], v);
                v *= inv;
            }
            ptl[nprev * PSTRF_CROWS + tid] = v;
            Pt[(int64_t)nprev * ld + p] = v;
            Lb[(int64_t)p * ld + j] = v;
            col = (pos == j) ? 0.0 : v;
        }
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int i = tid; i < n; i += PSTRF_CROWS) piv[i] = pivs[i];
        if (tid == 0) { st->rank = k + jb; if (k == 0) st->dstop = dstop; }
    }
    if (live) posg[p] = pos;
    cluster.sync();                                    // nobody leaves while a peer may still write into its slots
}

// Lp[i][c] = Lb[piv[i]][c] for c <= i (zero above): LAPACK's factor of P^T M P.
__global__ void pstrf_gather_kernel(const double *__restrict__ Lb, int64_t ld, const int32_t *__restrict__ piv, int n,
                                    double *__restrict__ Lp) {
    const int i = blockIdx.y;
    const int p = piv[i];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
        Lp[(int64_t)i * n + c] = c <= i ? Lb[(int64_t)p * ld + c] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Draws  Yt[dr][i] = mean[i] + sum_{j<=i} Z[j][dr] L[i][j]   (gsum/diagnostics.py:82, models.py:872: m + L z)
// Zn holds -z transposed (one draw per row) so the shared tile recurrence  acc -= A B^T  applies unchanged.
// ------------------------------------------------------------------------------------------------
struct DrawArgs {
    const double *Zn;     // (n_draws_pad x ld) rows = -z
    const double *L;      // (np x ld) lower factor, zeros above the diagonal, identity padding
    const double *mean;   // (n) or null
    double *Yt;           // (n_draws_pad x ld)
    int64_t ld;
    int n;
};
__global__ void __launch_bounds__(CHOL_THREADS, 2) draws_kernel(DrawArgs P) {
    extern __shared__ __align__(16) double smem[];
    const int i = blockIdx.x, dr = blockIdx.y;
    const double *Zi = P.Zn + (int64_t)dr * GSUM_TILE * P.ld;
    const double *Li = P.L + (int64_t)i * GSUM_TILE * P.ld;
    const int lane = threadIdx.x & 31, t = lane & 3;
    Acc acc;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const int c = i * GSUM_TILE + nt * 8 + 2 * t;
        const double m0 = (P.mean && c < P.n) ? P.mean[c] : 0.0;
        const double m1 = (P.mean && c + 1 < P.n) ? P.mean[c + 1] : 0.0;
#pragma unroll
        for (int mt = 0; mt < 2; mt++) { acc[mt][nt][0] = m0; acc[mt][nt][1] = m1; }
    }
    tile_accumulate(acc, Zi, Li, P.ld, P.ld, i + 1, false, 8, smem);
    tile_store_acc(acc, P.Yt + (int64_t)dr * GSUM_TILE * P.ld + i * GSUM_TILE, P.ld);
}

// Philox4x32-10 counter-based generator + Box-Muller: rows of -z (z ~ N(0,1)), zero beyond (n_draws, n).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// The counter is (pair index, GLOBAL draw index = first_draw + row): a shard of the draw axis reproduces the columns the
// unsharded call would generate.  `scale` (n_draws,) or NULL multiplies draw j's normals (multivariate-t draws).
__global__ void __launch_bounds__(256) normal_rows_kernel(double *__restrict__ Zn, int64_t ld, int64_t rows_pad, int64_t n_draws, int n,
                                                          uint64_t seed, int64_t first_draw, const double *__restrict__ scale) {
    const int64_t row = blockIdx.x, grow = first_draw + row;
    const double sc = (scale && row < n_draws) ? scale[row] : 1.0;
    for (int64_t pr = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; 2 * pr < ld; pr += (int64_t)gridDim.y * blockDim.x) {
        uint32_t c[4] = {(uint32_t)pr, (uint32_t)(pr >> 32), (uint32_t)grow, (uint32_t)(grow >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);   // (0,1)
        const double u2 = ((double)(((uint64_t)c[2] << 21) ^ (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const int64_t x = 2 * pr;
        const bool live = row < n_draws;
        Zn[row * ld + x] = (live && x < n) ? -(rad * cs) * sc : 0.0;
        if (x + 1 < ld) Zn[row * ld + x + 1] = (live && x + 1 < n) ? -(rad * sn) * sc : 0.0;
    }
}
// rows of caller-supplied normals scaled per draw (multivariate-t with a caller Z)
__global__ void __launch_bounds__(256) scale_rows_kernel(double *__restrict__ Zn, int64_t ld, int64_t n_draws, const double *__restrict__ scale) {
    const int64_t row = blockIdx.y;
    const double sc = scale[row];
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < ld; x += (int64_t)gridDim.x * blockDim.x) Zn[row * ld + x] *= sc;
}

// ------------------------------------------------------------------------------------------------
// Credible-interval coverage  (gsum/diagnostics.py:148-171)
// cov[dr][a] = mean_i 1[lower[a][i] < y[dr][i] < upper[a][i]].  128 draws per CTA share each staged
// (n_alpha x 32 points) slice of the bounds; counts are integers, so the result is order independent.
// Central intervals at increasing levels are NESTED at every point (lower non-increasing, upper non-decreasing in a);
// `coverage_nested_kernel` checks that on the device and the counting kernel then finds, per (draw, point), the first
// interval containing y by bisection (7 probes instead of n_alpha comparisons) and builds a histogram whose prefix
// sums are the counts.  Bounds in any other order take the comparison-per-interval path; both give the same integers.
// ------------------------------------------------------------------------------------------------
#define COVG_WARPS 16
#define COVG_DPW 8               // draws per warp: 128 draws per CTA share each staged slice of the bounds (the bounds would
                                 // otherwise cost 13x the L2 traffic of the draws themselves at n_alpha = 101)
#define COVG_ROWS (COVG_WARPS * COVG_DPW)
#define COVG_MAXA 128
__global__ void __launch_bounds__(256) coverage_nested_kernel(const double *__restrict__ lower, const double *__restrict__ upper, int n_alpha,
                                                              int n, int *__restrict__ nested) {
    const int64_t total = (int64_t)(n_alpha - 1) * n;
    bool ok = true;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        ok = ok && (lower[e + n] <= lower[e]) && (upper[e + n] >= upper[e]);          // false for NaN as well
    if (!ok) *nested = 0;
}
__global__ void __launch_bounds__(COVG_WARPS * 32) coverage_rows_kernel(const double *__restrict__ Yt, int64_t ld, int64_t n_draws, int n,
                                                                        const double *__restrict__ lower, const double *__restrict__ upper,
                                                                        int n_alpha, double *__restrict__ out,
                                                                        unsigned long long *__restrict__ counts,
                                                                        const int *__restrict__ nested_flag) {
    extern __shared__ __align__(16) double sm[];
    double *lo = sm, *up = sm + (size_t)n_alpha * 32;
    int *cnt = (int *)(up + (size_t)n_alpha * 32);       // [COVG_ROWS][n_alpha]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t dr0 = ((int64_t)blockIdx.x * COVG_WARPS + w) * COVG_DPW;       // this warp's first draw
    const bool nested = *nested_flag != 0;
    for (int e = tid; e < COVG_ROWS * n_alpha; e += COVG_WARPS * 32) cnt[e] = 0;
    for (int x0 = 0; x0 < n; x0 += 32) {
        __syncthreads();
        for (int e = tid; e < n_alpha * 32; e += COVG_WARPS * 32) {
            const int a = e >> 5, xx = x0 + (e & 31);
            lo[e] = xx < n ? lower[(int64_t)a * n + xx] : INFINITY;     // out-of-range points never count
            up[e] = xx < n ? upper[(int64_t)a * n + xx] : -INFINITY;
        }
        __syncthreads();
        const int xx = x0 + lane;
        double y[COVG_DPW];
#pragma unroll
        for (int u = 0; u < COVG_DPW; u++) y[u] = (xx < n && dr0 + u < n_draws) ? Yt[(dr0 + u) * ld + xx] : 0.0;
#pragma unroll
        for (int u = 0; u < COVG_DPW; u++) {
            if (dr0 + u >= n_draws) break;
            int *c = cnt + (w * COVG_DPW + u) * n_alpha;
            if (nested) {
                // first a with lower[a] < y < upper[a] (the predicate is monotone in a); n_alpha if there is none
                int lo_a = 0, hi_a = n_alpha;
                while (lo_a < hi_a) {
                    const int mid = (lo_a + hi_a) >> 1;
                    const bool in = (lo[mid * 32 + lane] < y[u]) && (y[u] < up[mid * 32 + lane]);
                    if (in) hi_a = mid; else lo_a = mid + 1;
                }
                if (xx < n && lo_a < n_alpha) atomicAdd(&c[lo_a], 1);
            } else {
                for (int a = 0; a < n_alpha; a++) {
                    const bool in = (lo[a * 32 + lane] < y[u]) && (y[u] < up[a * 32 + lane]);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (lane == 0) c[a] += __popc(m);
                }
            }
        }
    }
    __syncwarp();
    for (int u = 0; u < COVG_DPW; u++) {
        const int64_t dr = dr0 + u;
        if (dr >= n_draws) break;
        int *c = cnt + (w * COVG_DPW + u) * n_alpha;
        if (nested) {                                    // histogram of first-containing intervals -> counts: inclusive prefix sums
            int carry = 0;
            for (int a0 = 0; a0 < n_alpha; a0 += 32) {
                const int a = a0 + lane;
                int v = a < n_alpha ? c[a] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
                v += carry;
                if (a < n_alpha) c[a] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
            __syncwarp();
        }
        if (out)
            for (int a = lane; a < n_alpha; a += 32) out[dr * n_alpha + a] = (double)c[a] / (double)n;
    }
    if (counts) {                                        // integer totals over the rows of this launch: order independent
        __syncthreads();
        const int64_t first = (int64_t)blockIdx.x * COVG_ROWS;
        for (int a = tid; a < n_alpha; a += COVG_WARPS * 32) {
            unsigned long long t = 0;
            for (int rr = 0; rr < COVG_ROWS; rr++)
                if (first + rr < n_draws) t += (unsigned long long)cnt[rr * n_alpha + a];
            if (t) atomicAdd(&counts[a], t);
        }
    }
}

// max-shift normalisation of a log-likelihood grid (notebook cell 54): post = exp(ll - max), lse = max + log(sum post)
// Cluster version of the normalisation below: 8 CTAs of one thread-block cluster split the grid, exchange their partial
// maxima / sums through distributed shared memory (every CTA writes its partial into every peer's array) and meet at
// cluster barriers — one launch, no global scratch, fixed summation order.
#define GN_CLUSTER 8
__global__ void __cluster_dims__(GN_CLUSTER, 1, 1) __launch_bounds__(1024)
grid_normalize_cluster_kernel(const double *__restrict__ ll, int64_t count, double *__restrict__ post, double *__restrict__ lse) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double red[32];
    __shared__ double part_max[GN_CLUSTER], part_sum[GN_CLUSTER];
    const unsigned rank = cluster.block_rank();
    const int64_t stride = (int64_t)GN_CLUSTER * 1024, first = (int64_t)rank * 1024 + threadIdx.x;
    double mx = -INFINITY;
    for (int64_t i = first; i < count; i += stride) { const double v = ll[i]; if (v > mx) mx = v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < GN_CLUSTER) {
        double m = red[0];
        for (int i = 1; i < 32; i++) m = fmax(m, red[i]);
        *cluster.map_shared_rank(&part_max[rank], threadIdx.x) = m;
    }
    cluster.sync();
    mx = part_max[0];
#pragma unroll
    for (int i = 1; i < GN_CLUSTER; i++) mx = fmax(mx, part_max[i]);
    double s = 0.0;
    for (int64_t i = first; i < count; i += stride) {
        const double e = exp(ll[i] - mx);
        if (post) post[i] = e;
        s += e;
    }
    s = block_sum(s, red);
    if (threadIdx.x < GN_CLUSTER) *cluster.map_shared_rank(&part_sum[rank], threadIdx.x) = s;
    cluster.sync();
    if (lse && rank == 0 && threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < GN_CLUSTER; i++) t += part_sum[i];
        *lse = mx + log(t);
    }
}

__global__ void __launch_bounds__(1024) grid_normalize_kernel(const double *__restrict__ ll, int64_t count, double *__restrict__ post,
                                                              double *__restrict__ lse) {
    __shared__ double red[32];
    double mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < count; i += 1024) { const double v = ll[i]; if (v > mx) mx = v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 32; i++) mx = fmax(mx, red[i]);
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += 1024) {
        const double e = exp(ll[i] - mx);
        if (post) post[i] = e;
        s += e;
    }
    s = block_sum(s, red);
    if (lse && threadIdx.x == 0) *lse = mx + log(s);
}
