// Heterogeneous persistent schedule of the bordered Cholesky (K2 + K3 in ONE launch): roles, task lists and the factor
// workers.  The kernel itself is chol_hetero_tma_kernel (hetero_tma.cuh); the few-matrices regime is chain.cuh.
//
// What the measurements say (profiles/r01_notes.md, r02_notes.md, tools/fp64_latency.cu, tools/mma_bench.cu, tools/fp64_hog_ilp.cu):
//   * one DMMA-streaming warp per SM sub-partition reaches 84 % of the FP64 tensor issue rate, two reach 99.5 %;
//   * any FP64 dependency chain (POTRF, row substitution) that shares a sub-partition with a warp streaming INDEPENDENT
//     DMMAs is starved (275 cycles per dependent instruction beside one stream, ~10^4 beside two) — a warp running ONE
//     dependent DMMA chain (27 cycles a link) leaves it alone.
// So the latency chains and the streams must not share an SM, and whatever stays next to the streams must itself be DMMA:
//
//   GEMM CTAs   (most SMs; three math groups of four warps, one per sub-partition).  A task (i, k, b) runs
//               S = C_ik - sum_j L_ij L_kj^T with 16x64 warp tiles, operands streamed by TMA through an mbarrier ring, then
//                 i >  k : the triangular solve X = S L_kk^{-T} as DMMAs only — block substitution over 8-column blocks
//                          against M_kk, the tile L_kk whose 8x8 diagonal blocks were replaced by their inverses (8
//                          dependent steps of two DMMAs each, warp-local on an 8x64 row block, no FP64 scalar chain).
//                          Inverting only the 8x8 diagonal blocks keeps the accuracy of a true substitution (measured
//                          against extended precision: same error as LAPACK's dtrsm, where a full 64x64 inverse loses a
//                          digit);
//                 i == k : S goes back to global memory, flag := 1 (the SYRK half of a diagonal task).
//   factor CTAs (a few SMs, four independent 128-thread workers each) take the diagonal tiles: wait for S, POTRF in
//               shared memory, write L_kk, invert the eight 8x8 diagonal blocks, write M_kk, flag := 2.
//               Nothing else runs on their SM, so the chain sees the bare 32-cycle DFMA latency.
//
// Both roles claim their tasks in order from two lists derived from ONE topologically ordered list (df_build_tasks), so
// the earliest unfinished task of the joint order is always claimed and never waits: no deadlock with all CTAs
// co-resident (cooperative launch).  Every wait is bounded by the watchdog / abort flag of flags.cuh.
#pragma once
#include "flags.cuh"

#define HT_QD 4                         // task queue depth
#define HT_STR2(x) #x
#define HT_STR(x) HT_STR2(x)
#define HT_WORKER_DOUBLES (GSUM_TILE * GSUM_LDS + 4 * GSUM_TILE)     // factor worker: tile + diag + scratch + reciprocal pivots
#define HT_RINV (3 * GSUM_TILE)                                      // offset of the reciprocal pivots 1 / L_jj behind dg
#define HT_NSTAT 40
#ifndef HT_FACTOR_CTAS
#define HT_FACTOR_CTAS 12               // SMs given to the diagonal tiles (GSUM_B200_FACTOR_CTAS overrides)
#endif
#ifndef HT_FACTOR_WORKERS
#define HT_FACTOR_WORKERS 4             // 128-thread workers per factor CTA (measured on C4: 3 -> 1.740 ms, 4 -> 1.723 ms)
#endif
#ifndef HT_CHAIN_MAX
#define HT_CHAIN_MAX 40                 // batches up to this many matrices run in chain mode (chain.cuh)
#endif
#ifndef HT_DIAG_DELAY
#define HT_DIAG_DELAY 0
#endif

struct HeteroArgs {
    BorderedBatch P;
    const int4 *gtasks;     // GEMM tasks (i, k, b, flags) in schedule order; flags bit 0: thin border task (<= 8 rows in use)
    int ngtasks;
    const int4 *ftasks;     // factor tasks (k, b, 0, 0) in schedule order
    int nftasks;
    int nf0;                // leading entries of ftasks that belong to tile column 0 (no dependencies)
    int *ctl;               // [0] GEMM task counter, [1] abort flag, [2] sticky abort, [3] factor task counter, [4] column-0 factor counter
    int *flags;             // per (b, i, k): index (b * Trows + i) * T + k.  i > k: 1 = tile final.  i == k: 1 = S ready, 2 = L_kk and M_kk final
    double *M;              // (batch, T, 64, 64): L_kk with its 8x8 diagonal blocks inverted
    int nfactor_ctas;       // CTAs [0, nfactor_ctas) are factor CTAs
    int nworkers;           // 128-thread workers per factor CTA (1..4)
    long long *stats;       // optional per-CTA cycle counters [grid][HT_NSTAT]
    int chain;              // chain mode (chain.cuh): CTA c < nfactor_ctas is the chain worker of matrix c
    long long *trace; int trace_cta;      // STATS build: phase timestamps of one GEMM CTA's groups (tools/group_trace.py)
    int *pre;               // chain mode, per (b, k): 1 = the pre-panel tile (k+1, k) holds S' (everything but the triangular solve)
};

__device__ __forceinline__ bool flag_wait_ge(const int *flag, int want, int *abort_flag) {
    if (ld_relaxed(flag) >= want) return true;
    const long long t0 = clock64();
    while (ld_relaxed(flag) < want) {
        __nanosleep(64);
        if (ld_relaxed(abort_flag)) return false;
        if (clock64() - t0 > DF_WATCHDOG_CYCLES) { atomicExch(abort_flag, 1); return false; }
    }
    return true;
}
// M_kk (dense 64 x 64, ld 64) from the factored tile S (smem, stride GSUM_LDS; dg[j] = L_jj): L_kk below the 8x8
// diagonal blocks, the INVERSES of the diagonal blocks on them, zeros above.  Threads 64..127 copy the off-diagonal
// part while thread (cb, j) < 64 solves  L_blk x = e_j  by substitution in registers (reciprocal pivots, one multiply
// per step on the chain) and stores its column.  S is not modified.
__device__ __forceinline__ void ht_write_mkk(const double *S, const double *rinv, double *Mt, bool fail) {
    const int tid = EPI_TID;
    if (tid < 64) {
        const int cb = tid >> 3, j = tid & 7;
        const double *blk = S + (cb * 8) * GSUM_LDS + cb * 8;
        // right-looking: as soon as x[n] is final every later row takes its term (independent FMAs), so the dependent chain
        // is one multiply and one FMA per row; each row still sums its terms in the order n = 0, 1, ...
        double x[8], sacc[8];
#pragma unroll
        for (int m = 0; m < 8; m++) sacc[m] = (m == j) ? 1.0 : 0.0;
#pragma unroll
        for (int n = 0; n < 8; n++) {
            x[n] = (n >= j) ? sacc[n] * rinv[cb * 8 + n] : 0.0;
#pragma unroll
            for (int m = n + 1; m < 8; m++) sacc[m] = fma(-blk[m * GSUM_LDS + n], x[n], sacc[m]);
        }
#pragma unroll
        for (int m = 0; m < 8; m++) Mt[(cb * 8 + m) * GSUM_TILE + cb * 8 + j] = fail ? nan("") : x[m];
    } else {
        for (int e = tid - 64; e < GSUM_TILE * GSUM_TILE / 2; e += 64) {
            const int r = e >> 5, c = (e & 31) * 2;
            if ((c >> 3) == (r >> 3)) continue;            // diagonal block: written by the solvers
            double2 v;
            const bool below = (c >> 3) < (r >> 3);
            v.x = below ? S[r * GSUM_LDS + c] : 0.0;
            v.y = below ? S[r * GSUM_LDS + c + 1] : 0.0;
            if (fail) { v.x = v.y = nan(""); }
            *reinterpret_cast<double2 *>(Mt + r * GSUM_TILE + c) = v;
        }
    }
}

// POTRF of a 64x64 tile in shared memory (stride GSUM_LDS) by one 128-thread group, blocked by 8 columns, written so
// that neither its speed nor its correctness depends on how ptxas schedules it (an earlier version kept 36 + 8
// doubles live per thread and wants ~180 registers; in a kernel whose register target is lower, ptxas serialises its
// dependency chain and the tile takes 2x longer).  Per 8-column block:
//   F  warp 3 factors the 8x8 diagonal block in registers (every lane redundantly; chain rsqrt -> mul -> fma per column)
//      and leaves the block, its reciprocal pivots and diag(L) in a scratch area;
//   S  one thread per row below the block substitutes its 8 entries against the scratch block (64-cycle steps);
//   B  rank-8 DMMA update of the trailing 8x8 blocks — with one block of look-ahead: warp 3 updates the NEXT diagonal
//      block first and factors it (F of the next step) while warps 0-2 update the rest.
// Scratch behind the tile: dg[0..63] diag(L), dg[64..127] the factored block (row major 8x8), dg[136..143] 1 / L_jj of the
// current block, dg[192..255] 1 / L_jj of the whole tile (= rsqrt of the pivot: what the block inverses are scaled with).
// *s_fail: failing column (1-based, LAPACK potrf convention), 0 = ok; zeroed by the caller.
__device__ __forceinline__ void potrf_lean_factor_block(double *S, double *dg, int *s_fail, int cb, int lane) {
    const int c0 = cb * 8;
    double *wb = dg + GSUM_TILE, *rsd = dg + 2 * GSUM_TILE + 8;
    double a[8][8];
    const double *blk = S + c0 * GSUM_LDS + c0;
#pragma unroll
    for (int m = 0; m < 8; m++)
#pragma unroll
        for (int n = 0; n <= m; n += 2) {
            if (n + 1 <= m) {
                const double2 v = *reinterpret_cast<const double2 *>(blk + m * GSUM_LDS + n);
                a[m][n] = v.x; a[m][n + 1] = v.y;
            } else a[m][n] = blk[m * GSUM_LDS + n];
        }
    int fail = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double d = a[j][j];
        if (!(d > 0.0) && fail == 0) fail = c0 + j + 1;
        const double rs = rsqrt(d);
        a[j][j] = d * rs;
        if (lane == j) { rsd[j] = rs; dg[c0 + j] = d * rs; dg[HT_RINV + c0 + j] = rs; }
#pragma unroll
        for (int m = j + 1; m < 8; m++) a[m][j] *= rs;
#pragma unroll
        for (int m = j + 1; m < 8; m++)
#pragma unroll
            for (int n = j + 1; n <= m; n++) a[m][n] = fma(-a[m][j], a[n][j], a[m][n]);
    }
    // lane m < 8 writes row m of the factored block (scratch and tile)
#pragma unroll
    for (int m = 0; m < 8; m++)
        if (lane == m) {
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const double v = n <= m ? a[m][n] : 0.0;
                wb[m * 8 + n] = v;
                if (n <= m) S[(c0 + m) * GSUM_LDS + c0 + n] = v;
            }
        }
    if (lane == 0 && fail && *s_fail == 0) *s_fail = fail;
}
// rank-8 update of the 8x8 block (rb, cb2) by column block cb: one DMMA pair per warp
__device__ __forceinline__ void potrf_lean_update_block(double *S, int cb, int rb, int cb2, int g, int t) {
    const int c0 = cb * 8, off = (rb * 8 + g) * GSUM_LDS + cb2 * 8 + 2 * t;
    const double2 v = *reinterpret_cast<const double2 *>(S + off);
    double c0v = v.x, c1v = v.y;
    const double fa0 = S[(rb * 8 + g) * GSUM_LDS + c0 + t], fa1 = S[(rb * 8 + g) * GSUM_LDS + c0 + 4 + t];
    const double fb0 = S[(cb2 * 8 + g) * GSUM_LDS + c0 + t], fb1 = S[(cb2 * 8 + g) * GSUM_LDS + c0 + 4 + t];
    dmma884(c0v, c1v, -fa0, fb0);
    dmma884(c0v, c1v, -fa1, fb1);
    double2 o; o.x = c0v; o.y = c1v;
    *reinterpret_cast<double2 *>(S + off) = o;
}
__device__ __forceinline__ void tile_potrf_lean(double *S, double *dg, int *s_fail) {
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const double *wb = dg + GSUM_TILE, *rsd = dg + 2 * GSUM_TILE + 8;
    if (w == 3) potrf_lean_factor_block(S, dg, s_fail, 0, lane);
#pragma unroll 1
    for (int cb = 0; cb < 7; cb++) {
        const int c0 = cb * 8;
        CONS_SYNC();
        {
            // ---- S ----  row rr = c0 + 8 + tid
            const int rr = c0 + 8 + tid;
            if (rr < GSUM_TILE) {
                double *row = S + rr * GSUM_LDS + c0;
                double x[8];
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(row + c);
                    x[c] = v.x; x[c + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double v = x[c];
#pragma unroll
                    for (int m = 0; m < c; m++) v = fma(-x[m], wb[c * 8 + m], v);
                    x[c] = v * rsd[c];
                }
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 v; v.x = x[c]; v.y = x[c + 1];
                    *reinterpret_cast<double2 *>(row + c) = v;
                }
            }
        }
        CONS_SYNC();
        // ---- B with look-ahead ----  blocks (rb, cb2), cb < cb2 <= rb <= 7, numbered row by row; block 0 = (cb+1, cb+1)
        if (w == 3) {
            potrf_lean_update_block(S, cb, cb + 1, cb + 1, g, t);
            __syncwarp();
            potrf_lean_factor_block(S, dg, s_fail, cb + 1, lane);
        } else {
            const int nt = 7 - cb, nblk = nt * (nt + 1) / 2;
#pragma unroll 1
            for (int blk = 1 + w; blk < nblk; blk += 3) {
                int rbi = 0, rem = blk;
                while (rem > rbi) { rem -= rbi + 1; rbi++; }        // blk -> (rbi, rem) with rem <= rbi
                potrf_lean_update_block(S, cb, cb + 1 + rbi, cb + 1 + rem, g, t);
            }
        }
    }
    CONS_SYNC();
}

// ---- factor worker: one 128-thread group of a factor CTA ------------------------------------------------------------
// The first nf0 entries of the factor list are the column-0 tiles: nothing precedes them, so at the start of the launch
// EVERY CTA can take some (phase 0, own counter; `helper` = a math group of a GEMM CTA, which leaves after that phase)
// — the whole batch's first POTRFs then run in one round instead of queueing on the few factor CTAs while the GEMM CTAs
// have nothing to do.  Phase 1 is the rest of the list, factor workers only.
__device__ __forceinline__ void ht_factor_worker(const HeteroArgs &D, double *S, long long *st, bool helper = false) {
    const BorderedBatch &P = D.P;
    const int tid = EPI_TID;
    double *dg = S + GSUM_TILE * GSUM_LDS;
    int *s_fail = reinterpret_cast<int *>(dg + 2 * GSUM_TILE);
    int *s_tk = s_fail + 2;
    bool dead = false;
    for (int phase = 0; phase < (helper ? 1 : 2) && !dead; phase++)
    for (;;) {
        if (tid == 0) s_tk[0] = (phase == 0) ? atomicAdd(D.ctl + 4, 1) : D.nf0 + atomicAdd(D.ctl + 3, 1);
        CONS_SYNC();
        const int tix = s_tk[0];
        CONS_SYNC();                                  // everyone has read the claim before thread 0 writes the next one
        if (tix >= (phase == 0 ? D.nf0 : D.nftasks)) break;
        const int4 tk = D.ftasks[tix];
        const int k = tk.x, b = tk.y;
        double *C = P.A + (int64_t)b * P.bstride + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
        int *flag = D.flags + ((int64_t)b * P.Trows + k) * P.T + k;
        int ok = 1;
        const long long t0 = st ? clock64() : 0;
        if (k > 0 && tid == 0) ok = flag_wait_ge(flag, 1, D.ctl + 1) ? 1 : 0;
        if (!cons_sync_and(ok != 0)) { dead = true; break; }
        const long long t1 = st ? clock64() : 0;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int c = tid + q * CHOL_THREADS, row = c >> 5, ch = (c & 31) * 2;
            cp_async16(S + row * GSUM_LDS + ch, C + (int64_t)row * P.ld + ch);
        }
        cp_async_commit();
        cp_async_wait<0>();
        if (tid == 0) *s_fail = 0;
        CONS_SYNC();
        const long long t2 = st ? clock64() : 0;
        tile_potrf_lean(S, dg, s_fail);
        const long long t3 = st ? clock64() : 0;
        const int fail = *s_fail;
        // Critical path first: M_kk (what the panel tasks of this column wait for), then the flag; L_kk itself, the
        // log-determinant and the status are outputs nobody inside the launch reads.
        ht_write_mkk(S, dg + HT_RINV, D.M + ((int64_t)b * P.T + k) * (GSUM_TILE * GSUM_TILE), fail != 0);
        CONS_SYNC();                                  // every thread's M stores are ordered before the release below
        if (tid == 0) st_release(flag, 2);
        if (fail && tid == 0 && P.info[b] == 0) P.info[b] = k * GSUM_TILE + fail;
        // L_kk: lower triangle, exact zeros above the diagonal (numpy.linalg.cholesky convention)
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v;
            v.x = (c <= r) ? S[r * GSUM_LDS + c] : 0.0;
            v.y = (c + 1 <= r) ? S[r * GSUM_LDS + c + 1] : 0.0;
            if (fail) { v.x = v.y = nan(""); }
            *reinterpret_cast<double2 *>(C + (int64_t)r * P.ld + c) = v;
        }
        CONS_SYNC();                                  // S and dg are reused by the next tile
        if (st && tid == 0) { st[0] += t1 - t0; st[1] += clock64() - t1; st[2] += 1; st[3] += t2 - t1; st[4] += t3 - t2; }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
// flags for a fresh factorisation: all zero.  With an existing factor (solve_only): tiles i < T are final (1), diagonal
// tiles carry 2 (L_kk and M_kk final).
__global__ void ht_init_kernel(int *flags, int *ctl, int64_t batch, int Trows, int T, int factor_done) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2) ctl[idx] = 0;
    if (idx == 3 || idx == 4) ctl[idx] = 0;
    if (idx >= batch * Trows * T) {
        if (idx < batch * Trows * T + batch * T) flags[idx] = 0;      // the pre flags of chain mode
        return;
    }
    const int k = (int)(idx % T), i = (int)((idx / T) % Trows);
    flags[idx] = (factor_done && i < T) ? (i == k ? 2 : 1) : 0;
}
// Log-determinant parts from the finished factor: (batch, T) entries 2 * sum_j log(L_jj) over the 64 columns of diagonal tile
// k — same form as gsum/models.py:1015,1250; padding columns (>= n) contribute log 1 = 0; a failed tile was written as NaN and
// gives NaN.  A kernel of its own so that log() is not part of the factor / chain workers' code (instruction-cache footprint).
__global__ void __launch_bounds__(32) ht_logdet_kernel(BorderedBatch P) {
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double *C = P.A + (int64_t)b * P.bstride + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
    double v = 0.0;
    for (int j = tid; j < GSUM_TILE; j += 32)
        if (k * GSUM_TILE + j < P.n) v += log(C[(int64_t)j * P.ld + j]);
    v = warp_sum(v);
    if (tid == 0) P.logdet_part[(int64_t)b * P.T + k] = 2.0 * v;
}
// M_kk tiles from an existing factor (solve_only calls): one 128-thread CTA per diagonal tile
__global__ void __launch_bounds__(CHOL_THREADS) ht_mkk_from_factor_kernel(BorderedBatch P, double *M) {
    __shared__ __align__(16) double S[GSUM_TILE * GSUM_LDS];
    __shared__ double dg[GSUM_TILE];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double *C = P.A + (int64_t)b * P.bstride + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
    for (int e = tid; e < GSUM_TILE * GSUM_TILE; e += CHOL_THREADS) {
        const int r = e >> 6, c = e & 63;
        S[r * GSUM_LDS + c] = C[(int64_t)r * P.ld + c];
    }
    if (tid < GSUM_TILE) dg[tid] = 1.0 / C[(int64_t)tid * P.ld + tid];
    __syncthreads();
    ht_write_mkk(S, dg, M + ((int64_t)b * P.T + k) * (GSUM_TILE * GSUM_TILE), false);
}

// Chain mode (chain.cuh): the diagonal band belongs to the chain CTAs; the GEMM list holds the panel tasks of rows >= k+2
// and of the border rows, and the pre tasks (flag bit 1) that prepare the band tiles without their last two terms.
static inline void ht_build_chain_tasks(std::vector<int4> &gemm, int T, int Trows, int batch, bool thin_last) {
    gemm.clear();
    auto fl = [&](int i) { return (thin_last && i == Trows - 1 && i >= T) ? 1 : 0; };
    for (int k = 0; k < T; k++) {
        // panel tasks of column k: rows >= k+2 (row k+1 is the chain CTA's) and the border rows.  Right behind tile (k+3, k)
        // — the youngest operand they need — come the two pre tasks of the band step k+2 -> k+3, tiles (k+3, k+2) and
        // (k+3, k+3) without their last two terms: claimed two columns ahead of their use.
        for (int i = (k + 1 < T ? k + 2 : k + 1); i < Trows; i++) {
            for (int b = 0; b < batch; b++) gemm.push_back(make_int4(i, k, b, fl(i)));
            if (i == k + 3 && k + 3 < T) {
                for (int b = 0; b < batch; b++) gemm.push_back(make_int4(k + 3, k + 2, b, 2));
                for (int b = 0; b < batch; b++) gemm.push_back(make_int4(k + 3, k + 3, b, 2));
            }
        }
    }
}

// Split the joint topological order into the two claim lists.
// (Tried in round 2 and dropped: cutting the batch into groups of matrices that start a fraction of a factorisation apart,
// lists merged by cumulative work.  Workers claim in list order and block on the claimed task, so a group's tail tasks
// just park workers: 1 % at best, 5 % slower with four groups a quarter apart; profiles/r02_notes.md.)
static inline void ht_build_tasks(std::vector<int4> &gemm, std::vector<int4> &fact, int T, int Trows, int batch, bool solve_only,
                                  bool thin_last, int diag_delay) {
    std::vector<int4> all;
    df_build_tasks(all, T, Trows, batch, solve_only, thin_last, diag_delay);
    gemm.clear(); fact.clear();
    for (const int4 &tk : all) {
        if (tk.x == tk.y) {
            if (tk.y > 0) gemm.push_back(tk);                    // SYRK half (column 0 needs none: S = C)
            fact.push_back(make_int4(tk.y, tk.z, 0, 0));
        } else gemm.push_back(tk);
    }
}
