// K2/K3: blocked FP64 Cholesky + forward solves on a *bordered* batch of matrices.
//
// Layout (HBM): a batch of row-major matrices  A[b] : (Trows*64) x ld,  ld = T*64 (= N padded to 64).
//   rows [0, T*64)          the symmetric correlation matrix R (only tiles i >= k are read/written);
//                           padding rows/cols (>= N) hold the identity;
//   rows [T*64, Trows*64)   optional "border" rows: right-hand sides stored TRANSPOSED (one RHS per
//                           row, zero padded).  Running the same left-looking tile recurrence over
//                           them yields  Wt = RHSt * L^{-T},  i.e. the forward solve L^{-1} RHS for
//                           every RHS with the factor read once (reference: scipy cho_solve /
//                           solve_triangular call sites gsum/models.py:432-439,831,836,1032;
//                           gsum/helpers.py:505).
//
// The factorisation itself is the heterogeneous schedule of hetero.cuh / hetero_tma.cuh / chain.cuh.  This file holds the
// batch descriptor, the 16x64 warp-tile helpers (FP64 tensor-core DMMA.8x8x4, operands staged by cp.async into a 3-stage
// shared-memory ring) and the Schur-complement kernel built on them (posterior covariance, pivoted-Cholesky updates, draws).
#pragma once
#include "common.cuh"

#define CHOL_THREADS 128
// Barrier over one 128-thread group: the math threads (0..127) of the multi-launch / dataflow kernels use id 1, the
// epilogue groups of the pipeline kernel (threads 128.., 256..) ids 2, 3.  EPI_TID: thread index within the group.
// bar.sync is the ALIGNED barrier: every lane of a warp has to execute it together.  The callers reach it from code with
// lane-divergent work (a single writer lane, a partial warp of solvers); nothing obliges the compiler to reconverge a
// warp in front of an inline-asm barrier, and a warp that arrives with lanes missing releases the barrier early (seen
// as sporadic non-positive pivots in late 8-column blocks).  __syncwarp() makes the convergence explicit.
#define CONS_SYNC() do { __syncwarp(); asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory"); } while (0)
#define EPI_TID ((int)(threadIdx.x & 127))
#ifndef CHOL_NST
#define CHOL_NST 3
#endif
#ifndef CHOL_CTAS_PER_SM
#define CHOL_CTAS_PER_SM 2
#endif
#define CHOL_STAGE_DOUBLES (2 * GSUM_TILE * GSUM_LDH)                 // A half-slab + B half-slab
#define CHOL_SMEM_BYTES (CHOL_NST * CHOL_STAGE_DOUBLES * 8)           // 110592 B -> 2 CTAs / SM

struct BorderedBatch {
    double *A;          // factor part: base of the batch, (T*64) x ld per matrix
    int64_t ld;         // leading dimension (= T*64), shared by factor and border rows
    int64_t bstride;    // elements between consecutive matrices (factor part)
    double *W;          // border rows: base of the batch, ((Trows-T)*64) x ld per matrix (may alias A + T*64*ld)
    int64_t wstride;    // elements between consecutive border blocks (0: one block shared... not allowed for writes)
    int T;              // factor tile columns
    int Trows;          // total tile rows (>= T)
    int *info;          // per-matrix status: 0 ok, j+1 = first non-positive pivot (LAPACK potrf convention)
    double *logdet_part;  // (batch, T): sum_j 2*log(L_jj) over the 64 columns of diagonal tile k
    int n;              // true order N (columns >= n are identity padding and excluded from logdet)
    int border_used;    // border rows in use, counted from the first border row (0 = unknown: every border tile row is full)
};

// ---- operand staging ---------------------------------------------------------------------
__device__ __forceinline__ void chol_load_stage(double *st, const double *Ai, const double *Bk, int64_t lda,
                                                int64_t ldb, int h, bool same, int tid) {
    const int col0 = (h >> 1) * GSUM_TILE + (h & 1) * GSUM_KH;
    double *As = st, *Bs = st + GSUM_TILE * GSUM_LDH;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int c = tid + q * CHOL_THREADS;
        int row = c >> 4, ch = (c & 15) * 2;
        cp_async16(As + row * GSUM_LDH + ch, Ai + (int64_t)row * lda + col0 + ch);
    }
    if (!same) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int c = tid + q * CHOL_THREADS;
            int row = c >> 4, ch = (c & 15) * 2;
            cp_async16(Bs + row * GSUM_LDH + ch, Bk + (int64_t)row * ldb + col0 + ch);
        }
    }
}
// A whole 64x64 tile (the diagonal factor L_kk needed by the epilogue) into one stage buffer, row stride GSUM_LDS.
__device__ __forceinline__ void chol_load_tail(double *st, const double *src, int64_t ld, int tid) {
#pragma unroll
    for (int q = 0; q < 16; q++) {
        int c = tid + q * CHOL_THREADS;
        int row = c >> 5, ch = (c & 31) * 2;
        cp_async16(st + row * GSUM_LDS + ch, src + (int64_t)row * ld + ch);
    }
}

// ---- warp tile ------------------------------------------------------------------------------------------------
// Each of the 4 math warps owns 16 rows x 64 columns of the 64x64 CTA tile, held as DMMA C fragments:
//     acc[mt][nt][e]  <->  row = 16*w + 8*mt + g,  col = 8*nt + 2*t + e        (lane = 4*g + t)
// A whole tile row per warp is what lets the triangular solve of the epilogue run in registers (rows are independent).
typedef double Acc[2][8][2];

// One pipeline stage (K depth 32): acc -= A[16 rows of this warp] * B[8*ntm rows]^T.  `ntm` = n-tiles this warp needs
// (8 everywhere except on a diagonal tile, where warp w only owns columns < 16*(w+1)); FULL drops the predicates.
// MMA_PASSES > 1 walks the stage in passes over the n tiles (every accumulator then is a dependent chain over the 8
// k-steps and the warp has at most 16 / MMA_PASSES DMMAs in flight).  Kept as a switch for the record: measured
// (profiles/r01_fp64_latency.txt) it changes neither the stage time (2441 cycles) nor what a co-resident warp's FP64
// chain suffers beside the main loop — that is arbitration at the pipe, not queue depth.
#ifndef MMA_PASSES
#define MMA_PASSES 1
#endif
template <bool FULL>
__device__ __forceinline__ void stage_mma(Acc &acc, const double *As, const double *Bs, int ntm) {
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 3, g = lane >> 2, t = lane & 3;
    const double *ap = As + (w * 16 + g) * GSUM_LDH + t;
    const double *bp = Bs + g * GSUM_LDH + t;
    constexpr int NTP = 8 / MMA_PASSES;                  // n tiles per pass
#pragma unroll
    for (int n0 = 0; n0 < 8; n0 += NTP) {
#pragma unroll
        for (int ks = 0; ks < GSUM_KH / 4; ks++) {
            double a[2], b[NTP];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) a[mt] = -ap[mt * 8 * GSUM_LDH + ks * 4];
#pragma unroll
            for (int q = 0; q < NTP; q++) if (FULL || n0 + q < ntm) b[q] = bp[(n0 + q) * 8 * GSUM_LDH + ks * 4];
#pragma unroll
            for (int q = 0; q < NTP; q++)
                if (FULL || n0 + q < ntm) {
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) dmma884(acc[mt][n0 + q][0], acc[mt][n0 + q][1], a[mt], b[q]);
                }
        }
    }
}

// acc -= Ai[64 x 64*nslab] * Bk[64 x 64*nslab]^T     (acc preloaded by the caller; all 128 threads copy and sync)
// `tail` (optional) is one more 64x64 tile streamed through the same ring right behind the last operand slab — the
// epilogue's L_kk — so its latency hides under the main loop; the function returns the stage buffer it landed in.
__device__ __forceinline__ double *tile_accumulate(Acc &acc, const double *Ai, const double *Bk, int64_t lda, int64_t ldb, int nslab,
                                                   bool same, int ntm, double *smem, const double *tail = nullptr,
                                                   int64_t tail_ld = 0) {
    const int tid = threadIdx.x;
    const int nh = nslab * 2;
    const int total = nh + (tail ? 1 : 0);
#pragma unroll
    for (int s = 0; s < CHOL_NST - 1; s++) {
        if (s < nh) chol_load_stage(smem + s * CHOL_STAGE_DOUBLES, Ai, Bk, lda, ldb, s, same, tid);
        else if (s < total) chol_load_tail(smem + s * CHOL_STAGE_DOUBLES, tail, tail_ld, tid);
        cp_async_commit();
    }
    for (int h = 0; h < nh; h++) {
        cp_async_wait<CHOL_NST - 2>();
        __syncthreads();
        {
            const int hn = h + CHOL_NST - 1;
            double *dst = smem + (hn % CHOL_NST) * CHOL_STAGE_DOUBLES;
            if (hn < nh) chol_load_stage(dst, Ai, Bk, lda, ldb, hn, same, tid);
            else if (hn < total) chol_load_tail(dst, tail, tail_ld, tid);
            cp_async_commit();
        }
        const double *As = smem + (h % CHOL_NST) * CHOL_STAGE_DOUBLES;
        const double *Bs = same ? As : As + GSUM_TILE * GSUM_LDH;
        if (ntm == 8) stage_mma<true>(acc, As, Bs, 8);
        else stage_mma<false>(acc, As, Bs, ntm);
    }
    cp_async_wait<0>();
    __syncthreads();
    return tail ? smem + (nh % CHOL_NST) * CHOL_STAGE_DOUBLES : nullptr;
}

__device__ __forceinline__ void tile_load_acc(Acc &acc, const double *C, int64_t ldc, int ntm = 8) {
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 3, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            if (nt < ntm) {
                const double2 v = *reinterpret_cast<const double2 *>(C + (int64_t)(w * 16 + mt * 8 + g) * ldc + nt * 8 + 2 * t);
                acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
            } else { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        }
}
__device__ __forceinline__ void tile_store_acc(const Acc &acc, double *C, int64_t ldc) {
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 3, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
            *reinterpret_cast<double2 *>(C + (int64_t)(w * 16 + mt * 8 + g) * ldc + nt * 8 + 2 * t) = v;
        }
}

// Schur-complement tile over the border rows:  C(i,i') -= Wt_i Wt_i'^T  summed over all T factor columns.
// (Gram matrices of forward-solved RHS, and R_nn - V^T V for the posterior covariance, gsum/models.py:836.)
struct SchurArgs {
    const double *W;      // border rows base (row-major, ld), one batch entry
    int64_t ld, bstride;  // of W
    int T;                // slabs to contract over
    double *C;            // output (rows x ldc), updated in place
    int64_t ldc, cstride;
    int lower_only;       // skip tiles with i' > i
};
__global__ void __launch_bounds__(CHOL_THREADS, CHOL_CTAS_PER_SM) schur_kernel(SchurArgs P) {
    extern __shared__ __align__(16) double smem[];
    const int i = blockIdx.y, ip = blockIdx.x, b = blockIdx.z;
    if (P.lower_only && ip > i) return;
    const double *Wi = P.W + (int64_t)b * P.bstride + (int64_t)i * GSUM_TILE * P.ld;
    const double *Wp = P.W + (int64_t)b * P.bstride + (int64_t)ip * GSUM_TILE * P.ld;
    double *C = P.C + (int64_t)b * P.cstride + (int64_t)i * GSUM_TILE * P.ldc + ip * GSUM_TILE;
    Acc acc;
    tile_load_acc(acc, C, P.ldc);
    tile_accumulate(acc, Wi, Wp, P.ld, P.ld, P.T, i == ip, 8, smem);
    tile_store_acc(acc, C, P.ldc);
}

static inline int chol_set_attrs(gsum_ctx *ctx) {
    // per-device function attributes (a process may hold contexts on several devices)
    static bool done[64] = {false};
    int dev = ctx->device & 63;
    if (!done[dev]) {
        GSUM_CUDA(ctx, cudaFuncSetAttribute(schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
        done[dev] = true;
    }
    return 0;
}

static inline int schur_run(gsum_ctx *ctx, const SchurArgs &S, int tiles_rows, int batch) {
    GSUM_TRY(chol_set_attrs(ctx));
    ctx->launches += 1;
    dim3 grid(tiles_rows, tiles_rows, batch);
    schur_kernel<<<grid, CHOL_THREADS, CHOL_SMEM_BYTES, ctx->stream>>>(S);
    GSUM_CUDA(ctx, cudaGetLastError());
    return 0;
}
