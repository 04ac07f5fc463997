// Chain mode of the heterogeneous schedule: the few-matrices regime (a strong-scaled grid shard of 16 length scales, one
// N = 4096 covariance, one N = 2500 fit).
//
// With few matrices the launch is bounded by the per-column dependency chain, not by throughput.  In the many-matrices
// schedule a column costs  POTRF(k) -> flag -> [GEMM CTA: M_kk by TMA, triangular solve of tile (k+1, k), store, flag] ->
// [GEMM CTA: last SYRK stage of tile (k+1, k+1), store, flag] -> [factor CTA: load, POTRF(k+1)]:  six trips through L2 and
// three hand-overs between SMs on top of the POTRF (measured: 28 us per column of which the POTRF is 11).
//
// Here ONE chain CTA per matrix (alone on its SM) owns the diagonal band and keeps it in shared memory:
//   chain  (warps 0-3):  POTRF(k) | inverses of the 8x8 diagonal blocks | L_{k+1,k} = S' L_kk^{-T} as DMMAs from shared
//                        memory | its share of  D' -= L_{k+1,k} L_{k+1,k}^T | POTRF(k+1) ...   — no global round trip.
//   helper (warps 4-7):  under the POTRF it brings the two tiles of the next band step into shared memory and applies the
//                        last-but-one term to them (Y = L_{k+1,k-1}, the freshest tile a GEMM CTA produces for this step:
//                        P -= Y L_{k,k-1}^T,  D -= Y Y^T);  after the POTRF it writes M_kk (what the panel tasks of the
//                        GEMM CTAs wait for), takes its share of the rank-64 update, then writes L_kk / log-determinant.
//   publisher (warp 8):  the release store of the M_kk flag (its fence would otherwise sit on the helper's path).
// The GEMM CTAs prepare the two band tiles without their last TWO terms (`pre` tasks, flag bit 1):
//     pre-panel (k+1, k):   S' = A_{k+1,k}   - sum_{j<k-1} L_{k+1,j} L_{k,j}^T        (no triangular solve)  -> pre flag
//     pre-diag  (k+1, k+1): S' = A_{k+1,k+1} - sum_{j<k-1} L_{k+1,j} L_{k+1,j}^T                             -> flag 1
// which depend on nothing younger than column k-2: they are done long before they are needed.  Every other tile of column
// k (rows >= k+2, border rows) is a normal panel task of a GEMM CTA waiting for M_kk, as in the many-matrices schedule.
// All partial sums run over j in increasing order with the DMMA k order of the GEMM CTAs, so every tile equals bit for bit
// what the many-matrices schedule produces: a grid cell does not depend on how many length scales share its launch.
#pragma once
#include "hetero.cuh"

#define CH_TILE_DOUBLES (GSUM_TILE * GSUM_LDS)
#define CH_DV_LD 12                                     // row stride of an inverted 8x8 diagonal block: 12 % 16 -> conflict-free B fragments
#define CH_DV_BLOCK (8 * CH_DV_LD)
// per-column scratch, double-buffered by the parity of the column (the helper reads column k's while the chain factors k+1):
// [0, 256) the POTRF scratch of hetero.cuh (diag(L), factored block, status, reciprocal pivots), [256, 1024) the inverses
#define CH_SCR_DOUBLES 1024
#define CH_SCR_DV 256
#define CH_NTILES 5                                     // L_kk | next diagonal tile | sub-diagonal tile | previous sub-diagonal tile | Y
#define CH_SMEM_DOUBLES (CH_NTILES * CH_TILE_DOUBLES + 2 * CH_SCR_DOUBLES + 8)
#define CH_THREADS 256                                  // warps 0-3: the chain; warps 4-7: the output helper
#define CH_BAR_POTRF 3                                  // chain -> helper: L_kk, diag(L) and the inverted blocks are in shared memory
#define CH_BAR_LOADED 5                                 // helper -> chain: the two pre tiles of the column are in shared memory
#define CH_BAR_X 6                                      // both: L_{k+1,k} is complete in shared memory
#define CH_BAR_UPD 7                                    // both: the next diagonal tile is updated, Pt and the old L_kk buffer are free
#define CH_BAR_PUB 10                                   // helper -> publisher warp: M_kk is written (128 + 32 threads)
#define CH_BAR_XPUB 11                                  // chain -> publisher warp: L_{k+1,k} is written (128 + 32 threads)

__device__ __noinline__ void chain_load_tile(double *S, const double *C, int64_t ld, int tid) {
#pragma unroll 4
    for (int q = 0; q < 16; q++) {
        const int c = tid + q * CHOL_THREADS, row = c >> 5, ch = (c & 31) * 2;
        cp_async16(S + row * GSUM_LDS + ch, C + (int64_t)row * ld + ch);
    }
    cp_async_commit();
}
__device__ __forceinline__ void chain_bar_arrive(int id) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(CH_THREADS) : "memory"); }
__device__ __forceinline__ void chain_bar_sync(int id) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(CH_THREADS) : "memory"); }

// Inverses of the eight 8x8 diagonal blocks of the factored tile S (dg[j] = L_jj) into Dv: thread (cb, j) < 64 solves
// L_blk x = e_j by substitution in registers (same arithmetic as ht_write_mkk) and stores its column.
__device__ __forceinline__ void chain_invert_blocks(const double *S, const double *rinv, double *Dv, bool fail) {
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
        for (int m = 0; m < 8; m++) Dv[cb * CH_DV_BLOCK + m * CH_DV_LD + j] = fail ? nan("") : x[m];
    }
}

// hx_trsm_dinv (hetero_tma.cuh) with the inverted diagonal blocks in Dv and the blocks below them in the factored tile Ls (stride GSUM_LDS):
// register-resident (quad shuffles for the C -> A re-layouts).  A loop form working from shared memory (1.5 KB of code instead
// of 11) was measured slower: 6.2k instead of 4.7k cycles per tile (profiles/r02_notes.md).
template <int MT>
__device__ __forceinline__ void chain_trsm(Acc &T, const double *Ls, const double *Dv, int g, int t) {
    const unsigned FULLMASK = 0xffffffffu;
    const int s0 = t >> 1, s1 = 2 + (t >> 1);
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int cb = 0; cb < 8; cb++) {
        const int c0 = cb * 8;
        const double b0 = Dv[cb * CH_DV_BLOCK + g * CH_DV_LD + t], b1 = Dv[cb * CH_DV_BLOCK + g * CH_DV_LD + 4 + t];
        double a0[MT], a1[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const double p0 = __shfl_sync(FULLMASK, T[mt][cb][0], s0, 4), p1 = __shfl_sync(FULLMASK, T[mt][cb][1], s0, 4);
            const double q0 = __shfl_sync(FULLMASK, T[mt][cb][0], s1, 4), q1 = __shfl_sync(FULLMASK, T[mt][cb][1], s1, 4);
            a0[mt] = odd ? p1 : p0;
            a1[mt] = odd ? q1 : q0;
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double x0 = 0.0, x1 = 0.0;
            dmma884(x0, x1, a0[mt], b0);
            dmma884(x0, x1, a1[mt], b1);
            T[mt][cb][0] = x0; T[mt][cb][1] = x1;
        }
        if (cb == 7) break;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const double p0 = __shfl_sync(FULLMASK, T[mt][cb][0], s0, 4), p1 = __shfl_sync(FULLMASK, T[mt][cb][1], s0, 4);
            const double q0 = __shfl_sync(FULLMASK, T[mt][cb][0], s1, 4), q1 = __shfl_sync(FULLMASK, T[mt][cb][1], s1, 4);
            a0[mt] = -(odd ? p1 : p0);
            a1[mt] = -(odd ? q1 : q0);
        }
#pragma unroll
        for (int j = cb + 1; j < 8; j++) {
            const double l0 = Ls[(j * 8 + g) * GSUM_LDS + c0 + t], l1 = Ls[(j * 8 + g) * GSUM_LDS + c0 + 4 + t];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                dmma884(T[mt][j][0], T[mt][j][1], a0[mt], l0);
                dmma884(T[mt][j][0], T[mt][j][1], a1[mt], l1);
            }
        }
    }
}
// Dn -= X X^T on the 36 blocks on and below the diagonal, dealt over the eight warps of the chain CTA: chain warp v takes
// blocks 0..4 of block row 7 - v, helper warp v the rest of that row and block row v (5 and 4 blocks).
// The contraction runs in EXACTLY the order of a GEMM CTA's diagonal task (hetero_tma.cuh: 16-column boxes in sequence,
// k-step a of a box contracts columns {2a, 2a+1, 2a+8, 2a+9} in the DMMA's four k slots), so the tile equals bit for bit
// what the many-matrices schedule would have produced for it: a cell does not depend on the size of its batch.
__device__ __forceinline__ void chain_syrk_part(double *Dn, const double *X, int v, bool helper, int g, int t) {
    const int r0 = (7 - v) * 8, r1 = v * 8;
    const int lo0 = helper ? 5 : 0, hi0 = helper ? 7 - v : 4;          // block row 7 - v: blocks lo0..hi0
    const int hi1 = helper ? v : -1;                                   // block row v: blocks 0..hi1
    double c[2][8][2];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        if (nt >= lo0 && nt <= hi0) {
            const double2 q = *reinterpret_cast<const double2 *>(Dn + (r0 + g) * GSUM_LDS + nt * 8 + 2 * t);
            c[0][nt][0] = q.x; c[0][nt][1] = q.y;
        } else { c[0][nt][0] = 0.0; c[0][nt][1] = 0.0; }
        if (nt <= hi1) {
            const double2 q = *reinterpret_cast<const double2 *>(Dn + (r1 + g) * GSUM_LDS + nt * 8 + 2 * t);
            c[1][nt][0] = q.x; c[1][nt][1] = q.y;
        } else { c[1][nt][0] = 0.0; c[1][nt][1] = 0.0; }
    }
    const int ko = 8 * (t >> 1) + (t & 1);
    const double *ap0 = X + (r0 + g) * GSUM_LDS + ko, *ap1 = X + (r1 + g) * GSUM_LDS + ko, *bp = X + g * GSUM_LDS + ko;
#pragma unroll 4
    for (int ks = 0; ks < GSUM_TILE / 4; ks++) {
        const int kc = (ks >> 2) * 16 + (ks & 3) * 2;                // box base + 2a
        const double a0 = -ap0[kc], a1 = -ap1[kc];
        double b[8];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) if ((nt >= lo0 && nt <= hi0) || nt <= hi1) b[nt] = bp[nt * 8 * GSUM_LDS + kc];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            if (nt >= lo0 && nt <= hi0) dmma884(c[0][nt][0], c[0][nt][1], a0, b[nt]);
            if (nt <= hi1) dmma884(c[1][nt][0], c[1][nt][1], a1, b[nt]);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        if (nt >= lo0 && nt <= hi0) { double2 q; q.x = c[0][nt][0]; q.y = c[0][nt][1]; *reinterpret_cast<double2 *>(Dn + (r0 + g) * GSUM_LDS + nt * 8 + 2 * t) = q; }
        if (nt <= hi1) { double2 q; q.x = c[1][nt][0]; q.y = c[1][nt][1]; *reinterpret_cast<double2 *>(Dn + (r1 + g) * GSUM_LDS + nt * 8 + 2 * t) = q; }
    }
}

// One 8x8 block of  C -= A B^T  (K = 64) as ONE dependent DMMA chain, in the DMMA k order of the GEMM CTAs.  The helper runs
// these under the chain's POTRF: measured (tools/fp64_hog_ilp.cu), a warp with a single dependent DMMA chain still issues
// one DMMA per 27 cycles (63 % of the pipe) while a co-resident FP64 latency chain keeps its speed (DFMA 24 cycles,
// rsqrt + DADD 131 instead of 99) — two or more independent accumulators saturate the pipe and slow the neighbour's chain
// by 5-10x (DFMA 153, rsqrt 1095 cycles).
__device__ __forceinline__ void chain_block_update(double *Cm, const double *ap, const double *bp) {
    const double2 q = *reinterpret_cast<const double2 *>(Cm);
    double c0 = q.x, c1 = q.y;
#pragma unroll
    for (int ks = 0; ks < GSUM_TILE / 4; ks++) {
        const int kc = (ks >> 2) * 16 + (ks & 3) * 2;
        dmma884(c0, c1, -ap[kc], bp[kc]);
    }
    double2 o; o.x = c0; o.y = c1;
    *reinterpret_cast<double2 *>(Cm) = o;
}
// P -= Y Xp^T (64 x 64 x 64; warp w owns rows 16 w .. 16 w + 15) and D -= Y Y^T (the 36 lower blocks; block rows w and 7 - w):
// 25 blocks per helper warp, one dependent chain at a time, ONE copy of the chain code.
__device__ __forceinline__ void chain_helper_updates(double *Pm, double *Dm, const double *Y, const double *Xp, int w, int g, int t) {
    const int ko = 8 * (t >> 1) + (t & 1);
#pragma unroll 1
    for (int blk = 0; blk < 25; blk++) {
        int rb, nt;
        double *Cm;
        const double *Bm;
        if (blk < 16) { rb = 2 * w + (blk >> 3); nt = blk & 7; Cm = Pm; Bm = Xp; }
        else { const int q = blk - 16; rb = q <= w ? w : 7 - w; nt = q <= w ? q : q - w - 1; Cm = Dm; Bm = Y; }
        chain_block_update(Cm + (rb * 8 + g) * GSUM_LDS + nt * 8 + 2 * t, Y + (rb * 8 + g) * GSUM_LDS + ko, Bm + (nt * 8 + g) * GSUM_LDS + ko);
    }
}
// One place for the helper's three global waits.
__device__ __noinline__ int chain_wait_flag(const int *flag, int *abort_flag) {
    const int ok = flag_wait_ge(flag, 1, abort_flag) ? 1 : 0;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    return ok;
}

// ---- publisher (warp 8 of a chain CTA) ------------------------------------------------------------------------------------
__device__ __forceinline__ void ht_chain_publisher(const HeteroArgs &D, double *smem, int b) {
    const BorderedBatch &P = D.P;
    const int *s_ctl = reinterpret_cast<const int *>(smem + CH_NTILES * CH_TILE_DOUBLES + 2 * CH_SCR_DOUBLES);
    int *frow = D.flags + (int64_t)b * P.Trows * P.T;
    for (int k = 0; k < P.T; k++) {
        __syncwarp();
        asm volatile("bar.sync %0, 160;" ::"r"(CH_BAR_PUB) : "memory");
        if (s_ctl[0]) return;
        if ((threadIdx.x & 31) == 0) st_release(frow + (int64_t)k * P.T + k, 2);
        if (k + 1 < P.T) {
            __syncwarp();
            asm volatile("bar.sync %0, 160;" ::"r"(CH_BAR_XPUB) : "memory");
            if (s_ctl[0]) return;
            if ((threadIdx.x & 31) == 0) st_release(frow + (int64_t)(k + 1) * P.T + k, 1);
        }
    }
}

// ---- helper (warps 4-7 of a chain CTA) -----------------------------------------------------------------------------------
__device__ __forceinline__ void ht_chain_helper(const HeteroArgs &D, double *smem, int b) {
    const BorderedBatch &P = D.P;
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    double *Dk = smem, *Dn = smem + CH_TILE_DOUBLES, *Pt = smem + 2 * CH_TILE_DOUBLES, *Xp = smem + 3 * CH_TILE_DOUBLES;
    double *Y = smem + 4 * CH_TILE_DOUBLES;
    double *scr = smem + CH_NTILES * CH_TILE_DOUBLES;
    int *s_ctl = reinterpret_cast<int *>(scr + 2 * CH_SCR_DOUBLES);     // [0] abort
    double *Ab = P.A + (int64_t)b * P.bstride;
    int *abort_flag = D.ctl + 1;
    int *frow = D.flags + (int64_t)b * P.Trows * P.T;
    const int *pre = D.pre + (int64_t)b * P.T;
    if (tid == 0) s_ctl[0] = 0;
    CONS_SYNC();
    for (int k = 0; k < P.T; k++) {
        const bool more = (k + 1 < P.T);
        if (more) {
            // ---- the two tiles of the next band step, with every term but the chain's own ----------------------------------
            const double *C10 = Ab + (int64_t)(k + 1) * GSUM_TILE * P.ld + k * GSUM_TILE;
            int ok = 1;
            if (tid == 0 && k >= 2) ok = chain_wait_flag(pre + k, abort_flag) && chain_wait_flag(frow + (int64_t)(k + 1) * P.T + k + 1, abort_flag);
            if (k >= 2) ok = cons_sync_and(ok != 0) ? 1 : 0;
            if (ok) {
                chain_load_tile(Pt, C10, P.ld, tid);
                chain_load_tile(Dn, C10 + GSUM_TILE, P.ld, tid);
                if (k >= 1) {
                    // Y = L_{k+1,k-1}: the panel task of a GEMM CTA that waited for M_{k-1,k-1}
                    if (tid == 0) ok = chain_wait_flag(frow + (int64_t)(k + 1) * P.T + k - 1, abort_flag);
                    ok = cons_sync_and(ok != 0) ? 1 : 0;
                    if (ok) chain_load_tile(Y, C10 - GSUM_TILE, P.ld, tid);
                }
                cp_async_wait<0>();
                if (ok && k >= 1) {
                    CONS_SYNC();
                    chain_helper_updates(Pt, Dn, Y, Xp, w, g, t);   // P -= L_{k+1,k-1} L_{k,k-1}^T,  D -= L_{k+1,k-1} L_{k+1,k-1}^T
                }
            }
            if (!ok && tid == 0) s_ctl[0] = 1;
            chain_bar_arrive(CH_BAR_LOADED);
            if (!ok) {
                __syncwarp();
                asm volatile("bar.arrive %0, 160;" ::"r"(CH_BAR_PUB) : "memory");       // releases the publisher, which leaves
                return;
            }
        }
        chain_bar_sync(CH_BAR_POTRF);
        const double *dg = scr + (k & 1) * CH_SCR_DOUBLES, *Dv = dg + CH_SCR_DV;
        const int fail = reinterpret_cast<const int *>(dg + 2 * GSUM_TILE)[0];
        // M_kk (L_kk below its 8x8 diagonal blocks, their inverses on them, zeros above) and L_kk itself (lower triangle, exact
        // zeros above the diagonal: numpy.linalg.cholesky convention) in one pass over the tile
        double *Mt = D.M + ((int64_t)b * P.T + k) * (GSUM_TILE * GSUM_TILE);
        double *C = Ab + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
#pragma unroll 2
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            const bool blockdiag = (c >> 3) == (r >> 3), below = (c >> 3) < (r >> 3);
            double2 l = *reinterpret_cast<const double2 *>(Dk + r * GSUM_LDS + c), q;
            l.x = (c <= r) ? l.x : 0.0;
            l.y = (c + 1 <= r) ? l.y : 0.0;
            q = l;
            if (blockdiag) {
                const double *p = Dv + (r >> 3) * CH_DV_BLOCK + (r & 7) * CH_DV_LD + (c & 7);
                q.x = p[0]; q.y = p[1];
            } else if (!below) { q.x = 0.0; q.y = 0.0; }
            if (fail) { q.x = q.y = l.x = l.y = nan(""); }
            *reinterpret_cast<double2 *>(Mt + r * GSUM_TILE + c) = q;
            *reinterpret_cast<double2 *>(C + (int64_t)r * P.ld + c) = l;
        }
        if (fail && tid == 0 && P.info[b] == 0) P.info[b] = k * GSUM_TILE + fail;
        __syncwarp();
        asm volatile("bar.arrive %0, 160;" ::"r"(CH_BAR_PUB) : "memory");   // the publisher warp releases the M_kk flag
        if (more) {
            chain_bar_sync(CH_BAR_X);
            chain_syrk_part(Dn, Pt, w, true, g, t);
            chain_bar_sync(CH_BAR_UPD);
        }
        if (!more) break;
        CONS_SYNC();                                  // every helper thread has read L_kk and L_{k,k-1}: their buffers take the next prefetch
        double *tmp = Dk; Dk = Dn; Dn = tmp;
        tmp = Pt; Pt = Xp; Xp = tmp;                  // L_{k+1,k} is the "previous sub-diagonal tile" of the next step
    }
}

// ---- the chain (warps 0-3) ---------------------------------------------------------------------------------------------
// st[0] cycles waiting for the band tiles, st[1] POTRF, st[2] block inverses, st[3] solve, st[4] update, st[5] columns,
// st[6] waiting for the helper before the update
__device__ __forceinline__ void ht_chain_worker(const HeteroArgs &D, double *smem, int b, long long *st) {
    const BorderedBatch &P = D.P;
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    double *Dk = smem, *Dn = smem + CH_TILE_DOUBLES, *Pt = smem + 2 * CH_TILE_DOUBLES, *Xp = smem + 3 * CH_TILE_DOUBLES;
    double *scr = smem + CH_NTILES * CH_TILE_DOUBLES;
    const int *s_ctl = reinterpret_cast<const int *>(scr + 2 * CH_SCR_DOUBLES);
    double *Ab = P.A + (int64_t)b * P.bstride;
    int *frow = D.flags + (int64_t)b * P.Trows * P.T;
    chain_load_tile(Dk, Ab, P.ld, tid);
    cp_async_wait<0>();
    for (int k = 0; k < P.T; k++) {
        const bool more = (k + 1 < P.T);
        double *dg = scr + (k & 1) * CH_SCR_DOUBLES, *Dv = dg + CH_SCR_DV;
        int *s_fail = reinterpret_cast<int *>(dg + 2 * GSUM_TILE);
        if (tid == 0) s_fail[0] = 0;
        CONS_SYNC();
        long long t0 = st ? clock64() : 0;
        tile_potrf_lean(Dk, dg, s_fail);
        long long t1 = st ? clock64() : 0;
        chain_invert_blocks(Dk, dg + HT_RINV, Dv, s_fail[0] != 0);
        long long t1b = st ? clock64() : 0;
        chain_bar_arrive(CH_BAR_POTRF);               // the helper takes the outputs from here
        if (!more) break;
        long long t2 = st ? clock64() : 0;
        chain_bar_sync(CH_BAR_LOADED);                // the band tiles are in Pt and Dn (and the block inverses complete)
        if (s_ctl[0]) {
            __syncwarp();
            asm volatile("bar.arrive %0, 160;" ::"r"(CH_BAR_XPUB) : "memory");  // the publisher may already wait there; it leaves
            return;
        }
        long long t3 = st ? clock64() : 0;
        // ---- L_{k+1,k} = S' L_kk^{-T}: warp w owns rows 16 w .. 16 w + 15 ------------------------------------------------
        {
            Acc acc;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const double2 q = *reinterpret_cast<const double2 *>(Pt + (w * 16 + mt * 8 + g) * GSUM_LDS + nt * 8 + 2 * t);
                    acc[mt][nt][0] = q.x; acc[mt][nt][1] = q.y;
                }
            chain_trsm<2>(acc, Dk, Dv, g, t);
            double *C = Ab + (int64_t)(k + 1) * GSUM_TILE * P.ld + k * GSUM_TILE;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    double2 q; q.x = acc[mt][nt][0]; q.y = acc[mt][nt][1];
                    *reinterpret_cast<double2 *>(Pt + (w * 16 + mt * 8 + g) * GSUM_LDS + nt * 8 + 2 * t) = q;
                    *reinterpret_cast<double2 *>(C + (int64_t)(w * 16 + mt * 8 + g) * P.ld + nt * 8 + 2 * t) = q;
                }
        }
        long long t3b = st ? clock64() : 0;
        __syncwarp();
        asm volatile("bar.arrive %0, 160;" ::"r"(CH_BAR_XPUB) : "memory");      // the publisher warp releases the flag of L_{k+1,k}
        chain_bar_sync(CH_BAR_X);                     // X complete in Pt
        long long t4 = st ? clock64() : 0;
        chain_syrk_part(Dn, Pt, w, false, g, t);
        long long t4b = st ? clock64() : 0;
        chain_bar_sync(CH_BAR_UPD);
        if (st && tid == 0) { st[7] += t1b - t1; st[8] += t4b - t4; }
        if (st && tid == 0) { st[0] += t3 - t2; st[1] += t1 - t0; st[2] += t2 - t1; st[3] += t3b - t3; st[4] += clock64() - t4; st[5] += 1; st[6] += t4 - t3b; }
        double *tmp = Dk; Dk = Dn; Dn = tmp;
        tmp = Pt; Pt = Xp; Xp = tmp;
    }
}
