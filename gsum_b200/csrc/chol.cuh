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

// ---- epilogue 1: POTRF of a 64x64 tile held in smem (stride GSUM_LDS), blocked by 8 columns ------------------
// Per 8-column block: (1) every thread that owns a row below the block factors the 8x8 diagonal block redundantly in
// registers (no shuffles, no barrier: the dependent chain per column is rsqrt -> mul -> fma) and forward-substitutes its
// own row against it; warp 3 does the same factorisation and writes the block, diag(L) and the failure column back;
// (2) the trailing 8x8 blocks get a rank-8 DMMA update.  The column is scaled by the reciprocal square root, as LAPACK
// dpotf2 scales by the reciprocal of the pivot's square root.  *s_fail: failing column (1-based, LAPACK potrf
// convention), 0 = ok; must be zeroed by the caller.
__device__ __forceinline__ void tile_potrf_blocked_inl(double *S, double *dg, int *s_fail) {
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *wb = dg + GSUM_TILE;                        // 8x8 scratch for the factored diagonal block (behind diag(L))
#pragma unroll 1
    for (int cb = 0; cb < 8; cb++) {
        const int c0 = cb * 8;
        const int rr = c0 + 8 + tid;
        const bool solver = rr < GSUM_TILE;
        const bool writer = (tid == 96);            // one otherwise idle thread writes the factored block back
        if (solver || writer) {
            double a[8][8], x[8];
            const double *blk = S + c0 * GSUM_LDS + c0;
            double *row = S + (solver ? rr : 0) * GSUM_LDS + c0;
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int n = 0; n <= m; n += 2) {                   // 16-byte loads where both entries are in the lower triangle
                    if (n + 1 <= m) {
                        const double2 v = *reinterpret_cast<const double2 *>(blk + m * GSUM_LDS + n);
                        a[m][n] = v.x; a[m][n + 1] = v.y;
                    } else a[m][n] = blk[m * GSUM_LDS + n];
                }
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(row + c);
                x[c] = v.x; x[c + 1] = v.y;
            }
            int fail = 0;
            // column by column: factor column j of the block, then eliminate it from this thread's row (right-looking, so
            // a factor column is dead as soon as it has been applied)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double d = a[j][j];
                if (!(d > 0.0) && fail == 0) fail = c0 + j + 1;
                const double rs = rsqrt(d);
                x[j] *= rs;
                if (writer) { wb[j * 8 + j] = d * rs; dg[c0 + j] = d * rs; }
#pragma unroll
                for (int m = j + 1; m < 8; m++) {
                    a[m][j] *= rs;
                    if (writer) wb[m * 8 + j] = a[m][j];
                    x[m] = fma(-x[j], a[m][j], x[m]);
                }
#pragma unroll
                for (int m = j + 1; m < 8; m++)
#pragma unroll
                    for (int n = j + 1; n <= m; n++) a[m][n] = fma(-a[m][j], a[n][j], a[m][n]);
            }
            if (solver) {
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 v; v.x = x[c]; v.y = x[c + 1];
                    *reinterpret_cast<double2 *>(row + c) = v;
                }
            }
            if (writer && fail && *s_fail == 0) *s_fail = fail;
        }
        CONS_SYNC();
        // The factored block goes back only now: the solvers read the unfactored block at the start of the step, and a
        // writer that stored straight into it could overtake a solver warp that is still loading (seen with three POTRFs
        // sharing an SM: sporadic garbage rows -> non-positive pivots in later blocks).
        if (tid < 64 && (tid & 7) <= (tid >> 3)) S[(c0 + (tid >> 3)) * GSUM_LDS + c0 + (tid & 7)] = wb[tid];
        if (cb == 7) break;
        {   // trailing update of the 8x8 blocks (rb, cb2), cb < cb2 <= rb <= 7; warp w takes blocks w, w+4, ... (4 in flight)
            const int nt = 7 - cb, nblk = nt * (nt + 1) / 2;
#pragma unroll 1
            for (int q0 = 0; w + 4 * q0 < nblk; q0 += 4) {
                double cc[4][2], fa[4][2], fb[4][2];
                int off[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int blk = w + 4 * (q0 + q);
                    if (blk < nblk) {
                        int rbi = 0, rem = blk;
                        while (rem > rbi) { rem -= rbi + 1; rbi++; }        // blk -> (rbi, rem) with rem <= rbi
                        const int rb = cb + 1 + rbi, cb2 = cb + 1 + rem;
                        off[q] = (rb * 8 + g) * GSUM_LDS + cb2 * 8 + 2 * t;
                        const double2 v = *reinterpret_cast<const double2 *>(S + off[q]);
                        cc[q][0] = v.x; cc[q][1] = v.y;
                        fa[q][0] = S[(rb * 8 + g) * GSUM_LDS + c0 + t]; fa[q][1] = S[(rb * 8 + g) * GSUM_LDS + c0 + 4 + t];
                        fb[q][0] = S[(cb2 * 8 + g) * GSUM_LDS + c0 + t]; fb[q][1] = S[(cb2 * 8 + g) * GSUM_LDS + c0 + 4 + t];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (w + 4 * (q0 + q) < nblk) {
                        dmma884(cc[q][0], cc[q][1], -fa[q][0], fb[q][0]);
                        dmma884(cc[q][0], cc[q][1], -fa[q][1], fb[q][1]);
                    }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (w + 4 * (q0 + q) < nblk) {
                        double2 v; v.x = cc[q][0]; v.y = cc[q][1];
                        *reinterpret_cast<double2 *>(S + off[q]) = v;
                    }
            }
        }
        CONS_SYNC();
    }
    CONS_SYNC();
}

// Out-of-line copy for kernels whose register budget (168 at two CTAs per SM) the block step would overrun: the call
// costs one save/restore of the callee-saved registers per diagonal tile.  NOT for use under setmaxnreg (the pipeline
// kernel): caller regions compiled for a larger register count and this function disagree on the callee-saved set.
__device__ __noinline__ void tile_potrf_blocked(double *S, double *dg, int *s_fail) { tile_potrf_blocked_inl(S, dg, s_fail); }

// ---- epilogue 2: X = T * Lk^{-T} on a warp's 16 x 64 register block (rows are independent) ---------------------
// Right-looking over 8-column blocks: (1) forward-substitute the 8x8 diagonal block — a row's 8 entries live in the 4
// lanes of a quad, so each solved entry is broadcast with one quad shuffle and applied by fma (true substitution, no
// explicit inverse: keeps the row-wise backward stability the rtol 1e-10 parity relies on); (2) re-layout the solved
// block from C- to A-fragments with quad shuffles; (3) update the later column blocks with DMMAs (two per block and m
// tile, all independent).  Lk: L_kk in smem (stride GSUM_LDS), rdiag[j] = 1 / L_kk[j][j].  MT = m-tiles in use.
template <int MT>
__device__ __forceinline__ void trsm_regs(Acc &T, const double *Lk, const double *rdiag) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const unsigned FULLMASK = 0xffffffffu;
#pragma unroll
    for (int cb = 0; cb < 8; cb++) {
        const int c0 = cb * 8;
        asm volatile("" ::: "memory");         // keep the smem loads of later blocks from being hoisted (register pressure)
        // rows 2t and 2t+1 of the diagonal block (entries left of the diagonal) and their reciprocal pivots
        double l0[8], l1[8];
#pragma unroll
        for (int m = 0; m < 8; m += 2) {
            const double2 u = *reinterpret_cast<const double2 *>(Lk + (c0 + 2 * t) * GSUM_LDS + c0 + m);
            const double2 v = *reinterpret_cast<const double2 *>(Lk + (c0 + 2 * t + 1) * GSUM_LDS + c0 + m);
            l0[m] = u.x; l0[m + 1] = u.y; l1[m] = v.x; l1[m + 1] = v.y;
        }
        const double2 rd = *reinterpret_cast<const double2 *>(rdiag + c0 + 2 * t);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int owner = c >> 1, e = c & 1;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const double xv = T[mt][cb][e] * (e ? rd.y : rd.x);
                const double xc = __shfl_sync(FULLMASK, xv, owner, 4);
                if (t == owner) T[mt][cb][e] = xc;
                if (2 * t > c) T[mt][cb][0] = fma(-xc, l0[c], T[mt][cb][0]);
                if (2 * t + 1 > c) T[mt][cb][1] = fma(-xc, l1[c], T[mt][cb][1]);
            }
        }
        if (cb == 7) break;
        double a0[MT], a1[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const double p0 = __shfl_sync(FULLMASK, T[mt][cb][0], t >> 1, 4), p1 = __shfl_sync(FULLMASK, T[mt][cb][1], t >> 1, 4);
            const double q0 = __shfl_sync(FULLMASK, T[mt][cb][0], 2 + (t >> 1), 4), q1 = __shfl_sync(FULLMASK, T[mt][cb][1], 2 + (t >> 1), 4);
            a0[mt] = -((t & 1) ? p1 : p0);
            a1[mt] = -((t & 1) ? q1 : q0);
        }
#pragma unroll
        for (int j = cb + 1; j < 8; j++) {
            const double b0 = Lk[(j * 8 + g) * GSUM_LDS + c0 + t], b1 = Lk[(j * 8 + g) * GSUM_LDS + c0 + 4 + t];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                dmma884(T[mt][j][0], T[mt][j][1], a0[mt], b0);
                dmma884(T[mt][j][0], T[mt][j][1], a1[mt], b1);
            }
        }
    }
}

// Row-per-thread variant of the same solve.  DFMA latency on B200 is 32 cycles and a quad shuffle of a double 54, so the
// substitution chain is shortest when one thread owns a whole row of the 8-column block: per block the warp drops its
// 16 x 8 slice into a per-warp smem scratch (C-fragment layout -> row major), 16 lanes substitute one row each against
// the PRESCALED diagonal block  Lp[c][m] = L[c][m] / L[c][c]  (x_c = s_c / L_cc - sum_m x_m Lp[c][m]: one fma per step
// on the chain instead of fma + mul; still a true substitution), and the solved block comes back both as C fragments
// and — straight from the row-major scratch — as the A fragments of the DMMA update of the later blocks.
//   Lp:    [8 blocks][8][8] prescaled strictly-lower entries (smem), rdiag[64] = 1 / L_jj (smem)
//   scr:   this warp's scratch, 16 rows x TRSM_SCR_LD doubles
#define TRSM_SCR_LD 10
template <int MT>
__device__ __forceinline__ void trsm_rows(Acc &T, const double *Lk, const double *Lp, const double *rdiag, double *scr) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int cb = 0; cb < 8; cb++) {
        const int c0 = cb * 8;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double2 v; v.x = T[mt][cb][0]; v.y = T[mt][cb][1];
            *reinterpret_cast<double2 *>(scr + (mt * 8 + g) * TRSM_SCR_LD + 2 * t) = v;
        }
        __syncwarp();
        if (lane < 8 * MT) {
            double *row = scr + lane * TRSM_SCR_LD;
            const double *lp = Lp + cb * 64;
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(row + c);
                const double2 r = *reinterpret_cast<const double2 *>(rdiag + c0 + c);
                x[c] = v.x * r.x; x[c + 1] = v.y * r.y;
            }
#pragma unroll
            for (int c = 1; c < 8; c++) {
                double v = x[c];
#pragma unroll
                for (int m = 0; m < c; m++) v = fma(-x[m], lp[c * 8 + m], v);
                x[c] = v;
            }
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                double2 v; v.x = x[c]; v.y = x[c + 1];
                *reinterpret_cast<double2 *>(row + c) = v;
            }
        }
        __syncwarp();
        double a0[MT], a1[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const double2 v = *reinterpret_cast<const double2 *>(scr + (mt * 8 + g) * TRSM_SCR_LD + 2 * t);
            T[mt][cb][0] = v.x; T[mt][cb][1] = v.y;
            a0[mt] = -scr[(mt * 8 + g) * TRSM_SCR_LD + t];
            a1[mt] = -scr[(mt * 8 + g) * TRSM_SCR_LD + 4 + t];
        }
        __syncwarp();
        if (cb == 7) break;
#pragma unroll
        for (int j = cb + 1; j < 8; j++) {
            const double b0 = Lk[(j * 8 + g) * GSUM_LDS + c0 + t], b1 = Lk[(j * 8 + g) * GSUM_LDS + c0 + 4 + t];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                dmma884(T[mt][j][0], T[mt][j][1], a0[mt], b0);
                dmma884(T[mt][j][0], T[mt][j][1], a1[mt], b1);
            }
        }
    }
}
// rdiag[j] = 1 / L_jj and the prescaled diagonal blocks, by the 128 math threads (caller syncs afterwards)
__device__ __forceinline__ void trsm_prepare(const double *Lk, double *Lp, double *rdiag) {
    const int tid = EPI_TID;
    if (tid < GSUM_TILE) rdiag[tid] = 1.0 / Lk[tid * GSUM_LDS + tid];
    for (int e = tid; e < 512; e += CHOL_THREADS) {
        const int cb = e >> 6, c = (e >> 3) & 7, m = e & 7;
        const double d = Lk[(cb * 8 + c) * GSUM_LDS + cb * 8 + c];
        Lp[e] = m < c ? Lk[(cb * 8 + c) * GSUM_LDS + cb * 8 + m] * (1.0 / d) : 0.0;
    }
}

// ---- one tile task (i, k) of matrix b: accumulate, then POTRF (i == k) or TRSM (i > k); tile written once ---------
// Epilogue shared by the multi-launch and the dataflow schedules.  S: a free 64x68 smem buffer with 160 spare doubles
// behind it (diagonal tasks stage the tile there; panel tasks only use the spare doubles), Lk = L_kk staged in smem
// (panel tasks).  Called by the 128 math threads.  `es`: optional dev instrumentation (cycle counters).
template <bool INLINE_POTRF = false>
__device__ __forceinline__ void tile_epilogue(const BorderedBatch &P, int i, int k, int b, Acc &acc, double *S, double *Lk, double *C,
                                              long long *es = nullptr) {
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const bool diag = (i == k);
    double *dg = S + GSUM_TILE * GSUM_LDS;                                    // 64 doubles behind the tile
    int *s_fail = reinterpret_cast<int *>(dg + 2 * GSUM_TILE);
    const long long e0 = es ? clock64() : 0;
    if (diag) {
        // stage the lower part: warp w owns rows 16w.., columns < 16(w+1)
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
                if (nt < 2 * (w + 1)) {
                    double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                    *reinterpret_cast<double2 *>(S + (w * 16 + mt * 8 + g) * GSUM_LDS + nt * 8 + 2 * t) = v;
                }
        if (tid == 0) *s_fail = 0;
        CONS_SYNC();
        const long long e1 = es ? clock64() : 0;
        if (INLINE_POTRF) tile_potrf_blocked_inl(S, dg, s_fail); else tile_potrf_blocked(S, dg, s_fail);
        const long long e2 = es ? clock64() : 0;
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
        if (es && tid == 0) { es[0] += e1 - e0; es[1] += e2 - e1; es[2] += clock64() - e2; es[3] += 1; }
    } else {
        // panel tasks never stage the tile: S only carries rdiag, the prescaled diagonal blocks and the per-warp scratch
        double *rdiag = S, *Lp = S + GSUM_TILE, *scr = S + GSUM_TILE + 512 + w * (16 * TRSM_SCR_LD);
        trsm_prepare(Lk, Lp, rdiag);
        CONS_SYNC();
        const long long e1 = es ? clock64() : 0;
        trsm_rows<2>(acc, Lk, Lp, rdiag, scr);
        const long long e2 = es ? clock64() : 0;
        tile_store_acc(acc, C, P.ld);
        if (es && tid == 0) { es[4] += e1 - e0; es[5] += e2 - e1; es[6] += clock64() - e2; es[7] += 1; }
    }
}

__device__ __forceinline__ void tile_task(const BorderedBatch &P, int i, int k, int b, double *smem) {
    const int w = threadIdx.x >> 5;
    double *Ab = P.A + (int64_t)b * P.bstride;
    double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                           : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
    const double *Ak = Ab + (int64_t)k * GSUM_TILE * P.ld;
    double *C = Ri + k * GSUM_TILE;
    const bool diag = (i == k);
    const int ntm = diag ? 2 * (w + 1) : 8;              // diagonal tile: warp w owns columns < 16 (w + 1)
    Acc acc;
    tile_load_acc(acc, C, P.ld, ntm);
    double *Lk = tile_accumulate(acc, Ri, Ak, P.ld, P.ld, k, diag, ntm, smem, diag ? nullptr : Ak + k * GSUM_TILE, P.ld);
    double *S = smem + ((2 * k + 1) % CHOL_NST) * CHOL_STAGE_DOUBLES;      // a stage buffer the ring is done with
    tile_epilogue(P, i, k, b, acc, S, Lk, C);
}

// ---- multi-launch schedule: per tile column k one diagonal launch + one panel launch ------------------------------
__global__ void __launch_bounds__(CHOL_THREADS, CHOL_CTAS_PER_SM) chol_diag_kernel(BorderedBatch P, int k) {
    extern __shared__ __align__(16) double smem[];
    tile_task(P, k, k, blockIdx.x, smem);
}
// Tile rows i = i0 + blockIdx.x of tile column k (i0 = k+1 during a factorisation; i0 = T for a solve with an
// existing factor).  Rows >= T live in the border block W.
__global__ void __launch_bounds__(CHOL_THREADS, CHOL_CTAS_PER_SM) chol_panel_kernel(BorderedBatch P, int k, int i0) {
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
