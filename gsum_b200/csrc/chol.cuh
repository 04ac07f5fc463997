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
// One tile task (i,k), i >= k, computes   S = A_ik - sum_{j<k} L_ij L_kj^T   with the accumulator held in
// registers across the whole k-loop (FP64 tensor-core DMMA.8x8x4, operands staged by cp.async into a
// 3-stage shared-memory ring), then finishes with POTRF (i == k) or the triangular solve
// X = S L_kk^{-T} (i > k) and writes the tile exactly once.  Algorithmic traffic per task:
// 2*k*32 KiB of operand reads + one 32 KiB tile read/write.
#pragma once
#include "common.cuh"

#define CHOL_THREADS 128
// barrier over the 128 math threads only (the dataflow kernel adds a producer warp that must not take part)
#define CONS_SYNC() asm volatile("bar.sync 1, 128;" ::: "memory")
#define CHOL_NST 3
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

// acc(64x64, warp tile 32x32) -= Ai[64 x 64*nslab] * Bk[64 x 64*nslab]^T     (acc preloaded by the caller)
// `skip` lets a warp sit out the DMMA work (strict upper block of a diagonal tile) while still taking part in the
// copies and barriers.  `tail` (optional) is one more 64x64 tile streamed through the same ring right behind the last
// operand slab — the epilogue's L_kk — so its latency hides under the main loop; the function returns the stage buffer
// it landed in.  Fragments are double-buffered in registers so the LDS latency of step ks+1 hides under the DMMAs of ks.
__device__ __forceinline__ double *tile_accumulate(double (&acc)[4][4][2], const double *Ai, const double *Bk,
                                                   int64_t lda, int64_t ldb, int nslab, bool same, bool skip,
                                                   double *smem, const double *tail = nullptr, int64_t tail_ld = 0) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, t = lane & 3;
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
        if (!skip) {
            const double *As = smem + (h % CHOL_NST) * CHOL_STAGE_DOUBLES;
            const double *Bs = same ? As : As + GSUM_TILE * GSUM_LDH;
            const double *ap = As + (wm * 32 + g) * GSUM_LDH + t;
            const double *bp = Bs + (wn * 32 + g) * GSUM_LDH + t;
            double a[2][4], b[2][4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) { a[0][mi] = ap[mi * 8 * GSUM_LDH]; b[0][mi] = bp[mi * 8 * GSUM_LDH]; }
#pragma unroll
            for (int ks = 0; ks < GSUM_KH / 4; ks++) {
                const int cur = ks & 1, nxt = cur ^ 1;
                if (ks + 1 < GSUM_KH / 4) {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) {
                        a[nxt][mi] = ap[mi * 8 * GSUM_LDH + (ks + 1) * 4];
                        b[nxt][mi] = bp[mi * 8 * GSUM_LDH + (ks + 1) * 4];
                    }
                }
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], -a[cur][mi], b[cur][ni]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    return tail ? smem + (nh % CHOL_NST) * CHOL_STAGE_DOUBLES : nullptr;
}

__device__ __forceinline__ void tile_load_acc(double (&acc)[4][4][2], const double *C, int64_t ldc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const double2 v = *reinterpret_cast<const double2 *>(C + (int64_t)(wm * 32 + mi * 8 + g) * ldc + wn * 32 + ni * 8 + 2 * t);
            acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
        }
}

__device__ __forceinline__ void tile_store_acc_smem(const double (&acc)[4][4][2], double *S, int lds) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double *p = S + (wm * 32 + mi * 8 + g) * lds + wn * 32 + ni * 8 + 2 * t;
            p[0] = acc[mi][ni][0]; p[1] = acc[mi][ni][1];
        }
}

// ---- epilogue 1: POTRF of a 64x64 tile held in smem (stride GSUM_LDS), blocked by 8 columns ------------------
// Per 8-column block: (1) warp 0 factors the 8x8 diagonal block with its rows in registers (one lane per row, pivots and
// column entries exchanged by shuffles; column scaled by the reciprocal as LAPACK dpotf2 does); (2) the rows below are
// solved against it, one thread per row; (3) the trailing 8x8 blocks get a rank-8 DMMA update.  Writes dg = diag(L) and
// the failing column (1-based, LAPACK potrf convention) to *s_fail (0 = ok; must be zeroed by the caller).
__device__ __forceinline__ void tile_potrf_blocked(double *S, double *dg, int *s_fail) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int cb = 0; cb < 8; cb++) {
        const int c0 = cb * 8;
        if (w == 0) {
            const int r = lane & 7;                      // lanes >= 8 mirror lanes 0..7 (full-mask shuffles)
            double row[8];
#pragma unroll
            for (int c = 0; c < 8; c++) row[c] = S[(c0 + r) * GSUM_LDS + c0 + c];
            int fail = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                if (!(d > 0.0) && fail == 0) fail = c0 + j + 1;
                const double sj = sqrt(d);
                const double rs = 1.0 / sj;
                const double lj = row[j] * rs;
                row[j] = (r == j) ? sj : lj;
#pragma unroll
                for (int m = j + 1; m < 8; m++) {
                    const double lm = __shfl_sync(0xffffffffu, lj, m);
                    row[m] = fma(-lj, lm, row[m]);
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int c = 0; c < 8; c++) S[(c0 + r) * GSUM_LDS + c0 + c] = row[c];
                double dv = row[0];
#pragma unroll
                for (int c = 1; c < 8; c++) dv = (c == r) ? row[c] : dv;      // row[r] without a runtime-indexed register array
                dg[c0 + r] = dv;
            }
            if (lane == 0 && fail && *s_fail == 0) *s_fail = fail;
        }
        CONS_SYNC();
        if (cb == 7) break;
        {   // (2) rows below the diagonal block: X = S L_D^{-T}
            const int rr = c0 + 8 + tid;
            if (rr < GSUM_TILE) {
                double *row = S + rr * GSUM_LDS + c0;
                double x[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double v = row[c];
                    const double *lrow = S + (c0 + c) * GSUM_LDS + c0;
#pragma unroll
                    for (int m = 0; m < c; m++) v = fma(-x[m], lrow[m], v);
                    x[c] = v * (1.0 / dg[c0 + c]);
                }
#pragma unroll
                for (int c = 0; c < 8; c++) row[c] = x[c];
            }
        }
        CONS_SYNC();
        {   // (3) trailing update of the 8x8 blocks (rb, cb2), cb < cb2 <= rb <= 7
            const int nt = 7 - cb, nblk = nt * (nt + 1) / 2;
            for (int blk = w; blk < nblk; blk += CHOL_THREADS / 32) {
                int rbi = 0, rem = blk;
                while (rem > rbi) { rem -= rbi + 1; rbi++; }        // blk -> (rbi, rem) with rem <= rbi
                const int rb = cb + 1 + rbi, cb2 = cb + 1 + rem;
                double *cp = S + (rb * 8 + g) * GSUM_LDS + cb2 * 8 + 2 * t;
                double cc0 = cp[0], cc1 = cp[1];
#pragma unroll
                for (int k0 = 0; k0 < 8; k0 += 4) {
                    const double a = S[(rb * 8 + g) * GSUM_LDS + c0 + k0 + t];
                    const double b = S[(cb2 * 8 + g) * GSUM_LDS + c0 + k0 + t];
                    dmma884(cc0, cc1, -a, b);
                }
                cp[0] = cc0; cp[1] = cc1;
            }
        }
        CONS_SYNC();
    }
}

// ---- epilogue 2: X = S * Lkk^{-T} (rows independent; each warp owns 16 rows) --------------------------
// Blocked by 8 columns: DMMA update with the already-solved columns, then an 8x8 forward substitution
// per row.  True substitution (no explicit inverse of the diagonal tile): keeps the row-wise backward
// stability the rtol 1e-10 parity relies on.
__device__ __forceinline__ void tile_trsm_smem(double *S, const double *Lk, const double *rdiag) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = w * 16;
    for (int cb = 0; cb < 8; cb++) {
        if (cb > 0) {
            double c[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const double *p = S + (r0 + mt * 8 + g) * GSUM_LDS + cb * 8 + 2 * t;
                c[mt][0] = p[0]; c[mt][1] = p[1];
            }
            for (int k0 = 0; k0 < cb * 8; k0 += 4) {
                const double b = Lk[(cb * 8 + g) * GSUM_LDS + k0 + t];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    const double a = -S[(r0 + mt * 8 + g) * GSUM_LDS + k0 + t];
                    dmma884(c[mt][0], c[mt][1], a, b);
                }
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                double *p = S + (r0 + mt * 8 + g) * GSUM_LDS + cb * 8 + 2 * t;
                p[0] = c[mt][0]; p[1] = c[mt][1];
            }
            __syncwarp();
        }
        if (lane < 16) {
            double *row = S + (r0 + lane) * GSUM_LDS + cb * 8;
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                double v = row[c];
                const double *lrow = Lk + (cb * 8 + c) * GSUM_LDS + cb * 8;
#pragma unroll
                for (int m = 0; m < c; m++) v -= x[m] * lrow[m];
                x[c] = v * rdiag[cb * 8 + c];
            }
#pragma unroll
            for (int c = 0; c < 8; c++) row[c] = x[c];
        }
        __syncwarp();
    }
}

// ---- one tile task (i, k) of matrix b: accumulate, then POTRF (i == k) or TRSM (i > k); tile written once ---------
// Epilogue shared by the multi-launch and the dataflow schedules: S (a free 64x68 smem buffer with 160 spare doubles
// behind it), Lk = L_kk staged in smem (panel tasks).  Called by the 128 math threads.
__device__ __forceinline__ void tile_epilogue(const BorderedBatch &P, int i, int k, int b, double (&acc)[4][4][2], double *S,
                                              double *Lk, double *C, bool skip) {
    const int tid = threadIdx.x, w = tid >> 5;
    const bool diag = (i == k);
    double *dg = S + GSUM_TILE * GSUM_LDS;                                    // 64 doubles behind the tile
    int *s_fail = reinterpret_cast<int *>(dg + 2 * GSUM_TILE);
    if (!skip) tile_store_acc_smem(acc, S, GSUM_LDS);
    if (diag) {
        if (tid == 0) *s_fail = 0;
        CONS_SYNC();
        tile_potrf_blocked(S, dg, s_fail);
        const int fail = *s_fail;
        if (fail && tid == 0 && P.info[b] == 0) P.info[b] = k * GSUM_TILE + fail;
        // write L_kk: lower triangle, exact zeros above the diagonal (numpy.linalg.cholesky convention)
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v;
            v.x = (c <= r) ? S[r * GSUM_LDS + c] : 0.0;
            v.y = (c + 1 <= r) ? S[r * GSUM_LDS + c + 1] : 0.0;
            if (fail) { v.x = v.y = nan(""); }
            *reinterpret_cast<double2 *>(C + (int64_t)r * P.ld + c) = v;
        }
        if (P.logdet_part && w == 0) {
            // 2 * sum log(L_jj), same form as gsum/models.py:1015,1250; padding columns (>= n) contribute log 1 = 0
            double v = 0.0;
            for (int j = (tid & 31); j < GSUM_TILE; j += 32)
                if (k * GSUM_TILE + j < P.n) v += log(dg[j]);
            v = warp_sum(v);
            if (tid == 0) P.logdet_part[(int64_t)b * P.T + k] = fail ? nan("") : 2.0 * v;
        }
    } else {
        double *rdiag = dg;
        if (tid < GSUM_TILE) rdiag[tid] = 1.0 / Lk[tid * GSUM_LDS + tid];
        CONS_SYNC();
        tile_trsm_smem(S, Lk, rdiag);
        CONS_SYNC();
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v; v.x = S[r * GSUM_LDS + c]; v.y = S[r * GSUM_LDS + c + 1];
            *reinterpret_cast<double2 *>(C + (int64_t)r * P.ld + c) = v;
        }
    }
}


__device__ __forceinline__ void tile_task(const BorderedBatch &P, int i, int k, int b, double *smem) {
    const int tid = threadIdx.x, w = tid >> 5;
    double *Ab = P.A + (int64_t)b * P.bstride;
    double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                           : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
    const double *Ak = Ab + (int64_t)k * GSUM_TILE * P.ld;
    double *C = Ri + k * GSUM_TILE;
    const bool diag = (i == k);
    const bool skip = diag && (w == 1);                  // warp (wm=0, wn=1): strictly upper block of a diagonal tile
    double acc[4][4][2];
    if (!skip) tile_load_acc(acc, C, P.ld);
    double *Lk = tile_accumulate(acc, Ri, Ak, P.ld, P.ld, k, diag, skip, smem, diag ? nullptr : Ak + k * GSUM_TILE, P.ld);
    double *S = smem + ((2 * k + 1) % CHOL_NST) * CHOL_STAGE_DOUBLES;      // a stage buffer the ring is done with
    tile_epilogue(P, i, k, b, acc, S, Lk, C, skip);
}

// ---- multi-launch schedule: per tile column k one diagonal launch + one panel launch ------------------------------
__global__ void __launch_bounds__(CHOL_THREADS, 2) chol_diag_kernel(BorderedBatch P, int k) {
    extern __shared__ __align__(16) double smem[];
    tile_task(P, k, k, blockIdx.x, smem);
}
// Tile rows i = i0 + blockIdx.x of tile column k (i0 = k+1 during a factorisation; i0 = T for a solve with an
// existing factor).  Rows >= T live in the border block W.
__global__ void __launch_bounds__(CHOL_THREADS, 2) chol_panel_kernel(BorderedBatch P, int k, int i0) {
    extern __shared__ __align__(16) double smem[];
    tile_task(P, i0 + blockIdx.x, k, blockIdx.y, smem);
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
__global__ void __launch_bounds__(CHOL_THREADS, 2) schur_kernel(SchurArgs P) {
    extern __shared__ __align__(16) double smem[];
    const int i = blockIdx.y, ip = blockIdx.x, b = blockIdx.z;
    if (P.lower_only && ip > i) return;
    const double *Wi = P.W + (int64_t)b * P.bstride + (int64_t)i * GSUM_TILE * P.ld;
    const double *Wp = P.W + (int64_t)b * P.bstride + (int64_t)ip * GSUM_TILE * P.ld;
    double *C = P.C + (int64_t)b * P.cstride + (int64_t)i * GSUM_TILE * P.ldc + ip * GSUM_TILE;
    double acc[4][4][2];
    tile_load_acc(acc, C, P.ldc);
    tile_accumulate(acc, Wi, Wp, P.ld, P.ld, P.T, i == ip, false, smem);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double2 v; v.x = acc[mi][ni][0]; v.y = acc[mi][ni][1];
            *reinterpret_cast<double2 *>(C + (int64_t)(wm * 32 + mi * 8 + g) * P.ldc + wn * 32 + ni * 8 + 2 * t) = v;
        }
}

static inline int chol_set_attrs(gsum_ctx *ctx);

// Forward solve of the border rows against an existing factor (tile rows T..Trows-1 only).
static inline int chol_solve_border_run(gsum_ctx *ctx, const BorderedBatch &P, int batch) {
    GSUM_TRY(chol_set_attrs(ctx));
    const int nb = P.Trows - P.T;
    if (nb <= 0) return 0;
    for (int k = 0; k < P.T; k++) {
        dim3 grid(nb, batch);
        chol_panel_kernel<<<grid, CHOL_THREADS, CHOL_SMEM_BYTES, ctx->stream>>>(P, k, P.T);
    }
    ctx->launches += P.T;
    GSUM_CUDA(ctx, cudaGetLastError());
    return 0;
}

// Host-side schedule of the factorisation (+ border rows riding along).
static inline int chol_bordered_run(gsum_ctx *ctx, const BorderedBatch &P, int batch) {
    GSUM_TRY(chol_set_attrs(ctx));
    for (int k = 0; k < P.T; k++) {
        chol_diag_kernel<<<batch, CHOL_THREADS, CHOL_SMEM_BYTES, ctx->stream>>>(P, k);
        const int below = P.Trows - k - 1;
        if (below > 0) {
            dim3 grid(below, batch);
            chol_panel_kernel<<<grid, CHOL_THREADS, CHOL_SMEM_BYTES, ctx->stream>>>(P, k, k + 1);
            ctx->launches += 1;
        }
        ctx->launches += 1;
    }
    GSUM_CUDA(ctx, cudaGetLastError());
    return 0;
}

static inline int chol_set_attrs(gsum_ctx *ctx) {
    // per-device function attributes (a process may hold contexts on several devices)
    static bool done[64] = {false};
    int dev = ctx->device & 63;
    if (!done[dev]) {
        GSUM_CUDA(ctx, cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
        GSUM_CUDA(ctx, cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
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
