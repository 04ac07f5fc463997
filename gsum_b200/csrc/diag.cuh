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
#define PSTRF_THREADS 1024
#define PSTRF_NB 64

struct PstrfState {          // device-resident scalars
    int rank;                // columns completed
    int info;                // 0 ok, 1 = stopped early (matrix not positive definite to working precision)
    double dstop;
};

__global__ void pstrf_init_kernel(int32_t *piv, PstrfState *st, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) piv[i] = i;
    if (i == 0) { st->rank = 0; st->info = 0; st->dstop = 0.0; }
}

// One block of up to 64 columns starting at logical column k.  Single CTA; shared memory holds the running
// diagonal (dg), dpstrf's work array, the most recent factor column and the inverse permutation pos[] — all by
// physical row, so every sweep over the rows is a unit-stride sweep.  Pt (64 x ld, global, L2 resident) keeps the
// current block's factor columns TRANSPOSED so the per-column dgemv reads it coalesced.
__global__ void __launch_bounds__(PSTRF_THREADS, 1) pstrf_panel_kernel(const double *__restrict__ Af, double *__restrict__ Lb,
                                                                       double *__restrict__ Pt, int64_t ld, int n, int k,
                                                                       int32_t *piv, PstrfState *st) {
    extern __shared__ __align__(16) double sm[];
    double *work = sm, *dg = sm + n, *col = sm + 2 * n;
    double *lrow = sm + 3 * n;                       // 64: the pivot row's factor entries of this block
    double *redv = lrow + PSTRF_NB;                  // 32
    int *redi = (int *)(redv + 32);                  // 32
    int *pos = redi + 32;                            // n: logical position of physical row p
    __shared__ int s_pvt, s_stop;
    __shared__ double s_ajj;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (st->info != 0) return;                       // an earlier block already stopped
    const int jb = min(PSTRF_NB, n - k);
    for (int p = tid; p < n; p += PSTRF_THREADS) {
        work[p] = 0.0; dg[p] = Af[(int64_t)p * ld + p]; col[p] = 0.0;
        pos[piv[p]] = p;
    }
    if (tid == 0) s_stop = 0;
    __syncthreads();
    for (int j = k; j < k + jb; j++) {
        // ---- running diagonal + pivot search: first maximum in LOGICAL order among positions >= j ---------------
        double bv = -INFINITY; int bi = 0x7fffffff;
        for (int p = tid; p < n; p += PSTRF_THREADS) {
            const int i = pos[p];
            if (i < j) continue;
            if (j > k) work[p] += col[p] * col[p];
            const double c = dg[p] - work[p];
            if (c > bv || (c == bv && i < bi)) { bv = c; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { redv[w] = bv; redi[w] = bi; }
        __syncthreads();
        if (w == 0) {
            bv = redv[lane]; bi = redi[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                double dstop = st->dstop;
                if (j == 0) { dstop = (double)n * 1.1102230246251565e-16 * bv; st->dstop = dstop; }   // N * DLAMCH('Epsilon') * max diag
                const bool bad = (j == 0) ? !(bv > 0.0) : !(bv > dstop);
                if (bad || bi == 0x7fffffff) { s_stop = 1; st->info = 1; st->rank = j; }
                else {
                    const int pj = piv[bi], pold = piv[j];      // interchange logical positions j and pvt
                    piv[bi] = pold; piv[j] = pj;
                    pos[pold] = bi; pos[pj] = j;
                    s_pvt = pj; s_ajj = bv;
                }
            }
        }
        __syncthreads();
        if (s_stop) return;
        const int pj = s_pvt;
        const double ajj = sqrt(s_ajj), inv = 1.0 / ajj;
        const int nprev = j - k;
        if (tid < nprev) lrow[tid] = Pt[(int64_t)tid * ld + pj];
        __syncthreads();
        // ---- column j (dgemv with the block's previous columns, then dscal by 1/ajj) ---------------------------
        for (int p = tid; p < n; p += PSTRF_THREADS) {
            const int i = pos[p];
            if (i < j) continue;                    // already pivoted: its entry in column j stays an exact 0
            double v;
            if (i == j) v = ajj;
            else {
                v = Af[(int64_t)pj * ld + p];
                const double *pp = Pt + p;
#pragma unroll 8
                for (int m = 0; m < nprev; m++) v = fma(-pp[(int64_t)m * ld], lrow[m], v);
                v *= inv;
            }
            Pt[(int64_t)nprev * ld + p] = v;
            Lb[(int64_t)p * ld + j] = v;
            col[p] = (i == j) ? 0.0 : v;
        }
        __syncthreads();
    }
    if (tid == 0) st->rank = k + jb;
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
// cov[dr][a] = mean_i 1[lower[a][i] < y[dr][i] < upper[a][i]].  16 draws per CTA share each staged
// (n_alpha x 32 points) slice of the bounds; counts are integers, so the result is order independent.
// ------------------------------------------------------------------------------------------------
#define COVG_WARPS 16
#define COVG_MAXA 128
__global__ void __launch_bounds__(COVG_WARPS * 32) coverage_rows_kernel(const double *__restrict__ Yt, int64_t ld, int64_t n_draws, int n,
                                                                        const double *__restrict__ lower, const double *__restrict__ upper,
                                                                        int n_alpha, double *__restrict__ out,
                                                                        unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) double sm[];
    double *lo = sm, *up = sm + (size_t)n_alpha * 32;
    int *cnt = (int *)(up + (size_t)n_alpha * 32);       // [COVG_WARPS][n_alpha]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t dr = (int64_t)blockIdx.x * COVG_WARPS + w;
    for (int e = tid; e < COVG_WARPS * n_alpha; e += COVG_WARPS * 32) cnt[e] = 0;
    for (int x0 = 0; x0 < n; x0 += 32) {
        __syncthreads();
        for (int e = tid; e < n_alpha * 32; e += COVG_WARPS * 32) {
            const int a = e >> 5, xx = x0 + (e & 31);
            lo[e] = xx < n ? lower[(int64_t)a * n + xx] : INFINITY;     // out-of-range points never count
            up[e] = xx < n ? upper[(int64_t)a * n + xx] : -INFINITY;
        }
        __syncthreads();
        if (dr < n_draws) {
            const int xx = x0 + lane;
            const double y = xx < n ? Yt[dr * ld + xx] : 0.0;
            for (int a = 0; a < n_alpha; a++) {
                const bool in = (lo[a * 32 + lane] < y) && (y < up[a * 32 + lane]);
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (lane == 0) cnt[w * n_alpha + a] += __popc(m);
            }
        }
    }
    __syncwarp();
    if (out && dr < n_draws)
        for (int a = lane; a < n_alpha; a += 32) out[dr * n_alpha + a] = (double)cnt[w * n_alpha + a] / (double)n;
    if (counts) {                                        // integer totals over the rows of this launch: order independent
        __syncthreads();
        for (int a = tid; a < n_alpha; a += COVG_WARPS * 32) {
            unsigned long long t = 0;
            for (int ww = 0; ww < COVG_WARPS; ww++)
                if ((int64_t)blockIdx.x * COVG_WARPS + ww < n_draws) t += (unsigned long long)cnt[ww * n_alpha + a];
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
