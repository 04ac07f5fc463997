// Chain mode of the heterogeneous schedule: the few-matrices regime (a strong-scaled grid shard of 16 length scales, one
// N = 4096 covariance, one N = 2500 fit).
//
// With few matrices the launch is bounded by the per-column dependency chain, not by throughput.  In the many-matrices
// schedule a column costs  POTRF(k) -> flag -> [GEMM CTA: M_kk by TMA, triangular solve of tile (k+1, k), store, flag] ->
// [GEMM CTA: last SYRK stage of tile (k+1, k+1), store, flag] -> [factor CTA: load, POTRF(k+1)]:  six trips through L2 and
// three hand-overs between SMs on top of the POTRF (measured: 28 us per column of which the POTRF is 11).
//
// Here ONE 128-thread chain worker per matrix (alone on its SM) owns the whole diagonal band: it keeps L_kk in shared
// memory and performs the triangular solve of the sub-diagonal tile (k+1, k) and the rank-64 update of the next diagonal
// tile itself, as DMMAs from shared memory.  The GEMM CTAs prepare both tiles WITHOUT their last term, which depends on
// nothing of column k (`pre` tasks):
//     pre-panel (k+1, k):   S' = A_{k+1,k}   - sum_{j<k} L_{k+1,j} L_{k,j}^T        (no triangular solve)  -> pre flag
//     pre-diag  (k+1, k+1): S' = A_{k+1,k+1} - sum_{j<k} L_{k+1,j} L_{k+1,j}^T                             -> flag 1
// so both are ready long before POTRF(k) ends and are prefetched under it.  The chain per column is then
//     POTRF(k) | M_kk + flag | solve L_{k+1,k} = S' L_kk^{-T} | flag | S'' = S' - L_{k+1,k} L_{k+1,k}^T | POTRF(k+1)
// with no global-memory round trip on it.  Every other tile of column k (rows >= k+2, border rows) is a normal panel task
// of a GEMM CTA waiting for M_kk, exactly as in the many-matrices schedule.
#pragma once
#include "hetero.cuh"

#define CH_TILE_DOUBLES (GSUM_TILE * GSUM_LDS)
#define CH_DV_LD 12                                     // row stride of an inverted 8x8 diagonal block: 12 % 16 -> conflict-free B fragments
#define CH_DV_BLOCK (8 * CH_DV_LD)
#define CH_SMEM_DOUBLES (3 * CH_TILE_DOUBLES + 3 * GSUM_TILE + 8 * CH_DV_BLOCK)
#define CH_THREADS 256                                  // warps 0-3: the chain; warps 4-7: the output helper
#define CH_BAR_POTRF 4                                  // chain -> helper: L_kk, diag(L) and the inverted blocks are in shared memory
#define CH_BAR_FREE 5                                   // helper -> chain: the outputs of the column are written, its buffers are free

__device__ __forceinline__ void chain_load_tile(double *S, const double *C, int64_t ld, int tid) {
#pragma unroll
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
__device__ __forceinline__ void chain_invert_blocks(const double *S, const double *dg, double *Dv, bool fail) {
    const int tid = EPI_TID;
    if (tid < 64) {
        const int cb = tid >> 3, j = tid & 7;
        const double *blk = S + (cb * 8) * GSUM_LDS + cb * 8;
        double x[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            double s = (m == j) ? 1.0 : 0.0;
#pragma unroll
            for (int n = 0; n < m; n++) s = fma(-blk[m * GSUM_LDS + n], x[n], s);
            x[m] = (m >= j) ? s * (1.0 / dg[cb * 8 + m]) : 0.0;
        }
#pragma unroll
        for (int m = 0; m < 8; m++) Dv[cb * CH_DV_BLOCK + m * CH_DV_LD + j] = fail ? nan("") : x[m];
    }
}

// ht_trsm_dinv with the inverted diagonal blocks in Dv and the blocks below them in the factored tile Ls (stride GSUM_LDS).
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

// Dn -= X X^T on the 36 blocks on and below the diagonal; warp w takes the block rows w and 7 - w (nine blocks each).
// The contraction runs in EXACTLY the order of a GEMM CTA's diagonal task (hetero_tma.cuh: 16-column boxes in sequence,
// k-step a of a box contracts columns {2a, 2a+1, 2a+8, 2a+9} in the DMMA's four k slots), so the tile equals bit for bit
// what the many-matrices schedule would have produced for it: a cell does not depend on the size of its batch.
__device__ __forceinline__ void chain_syrk(double *Dn, const double *X, int w, int g, int t) {
    const int r0 = w * 8, r1 = (7 - w) * 8;
    double c[2][8][2];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        if (nt <= w) {
            const double2 v = *reinterpret_cast<const double2 *>(Dn + (r0 + g) * GSUM_LDS + nt * 8 + 2 * t);
            c[0][nt][0] = v.x; c[0][nt][1] = v.y;
        } else { c[0][nt][0] = 0.0; c[0][nt][1] = 0.0; }
        if (nt <= 7 - w) {
            const double2 v = *reinterpret_cast<const double2 *>(Dn + (r1 + g) * GSUM_LDS + nt * 8 + 2 * t);
            c[1][nt][0] = v.x; c[1][nt][1] = v.y;
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
        for (int nt = 0; nt < 8; nt++) if (nt <= w || nt <= 7 - w) b[nt] = bp[nt * 8 * GSUM_LDS + kc];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            if (nt <= w) dmma884(c[0][nt][0], c[0][nt][1], a0, b[nt]);
            if (nt <= 7 - w) dmma884(c[1][nt][0], c[1][nt][1], a1, b[nt]);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        if (nt <= w) { double2 v; v.x = c[0][nt][0]; v.y = c[0][nt][1]; *reinterpret_cast<double2 *>(Dn + (r0 + g) * GSUM_LDS + nt * 8 + 2 * t) = v; }
        if (nt <= 7 - w) { double2 v; v.x = c[1][nt][0]; v.y = c[1][nt][1]; *reinterpret_cast<double2 *>(Dn + (r1 + g) * GSUM_LDS + nt * 8 + 2 * t) = v; }
    }
}

// ---- output helper (warps 4-7 of a chain CTA): everything of a column that nothing on the chain waits for ---------------
// After the chain's POTRF of column k: M_kk to global memory and its flag (what the panel tasks of the GEMM CTAs wait for),
// then L_kk, the log-determinant and the status.
__device__ __forceinline__ void ht_chain_helper(const HeteroArgs &D, double *smem, int b) {
    const BorderedBatch &P = D.P;
    const int tid = EPI_TID;
    double *Dk = smem, *Dn = smem + 2 * CH_TILE_DOUBLES;
    const double *dg = smem + 3 * CH_TILE_DOUBLES, *Dv = dg + 3 * GSUM_TILE;
    const int *s_fail = reinterpret_cast<const int *>(dg + 2 * GSUM_TILE);
    double *Ab = P.A + (int64_t)b * P.bstride;
    int *frow = D.flags + (int64_t)b * P.Trows * P.T;
    for (int k = 0; k < P.T; k++) {
        chain_bar_sync(CH_BAR_POTRF);
        const int fail = s_fail[0];
        if (s_fail[1]) return;                          // the chain is aborting
        double *Mt = D.M + ((int64_t)b * P.T + k) * (GSUM_TILE * GSUM_TILE);
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v;
            if ((c >> 3) == (r >> 3)) {
                const double *q = Dv + (r >> 3) * CH_DV_BLOCK + (r & 7) * CH_DV_LD + (c & 7);
                v.x = q[0]; v.y = q[1];
            } else if ((c >> 3) < (r >> 3)) {
                v.x = Dk[r * GSUM_LDS + c]; v.y = Dk[r * GSUM_LDS + c + 1];
            } else { v.x = 0.0; v.y = 0.0; }
            if (fail) { v.x = v.y = nan(""); }
            *reinterpret_cast<double2 *>(Mt + r * GSUM_TILE + c) = v;
        }
        CONS_SYNC();                                  // every thread's M stores are ordered before the release below
        if (tid == 0) st_release(frow + (int64_t)k * P.T + k, 2);
        if (fail && tid == 0 && P.info[b] == 0) P.info[b] = k * GSUM_TILE + fail;
        // L_kk: lower triangle, exact zeros above the diagonal (numpy.linalg.cholesky convention)
        double *C = Ab + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
        for (int e = tid; e < GSUM_TILE * GSUM_TILE / 2; e += CHOL_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v;
            v.x = (c <= r) ? Dk[r * GSUM_LDS + c] : 0.0;
            v.y = (c + 1 <= r) ? Dk[r * GSUM_LDS + c + 1] : 0.0;
            if (fail) { v.x = v.y = nan(""); }
            *reinterpret_cast<double2 *>(C + (int64_t)r * P.ld + c) = v;
        }
        if (P.logdet_part && tid < 32) {
            // 2 * sum log(L_jj), same form as gsum/models.py:1015,1250; padding columns (>= n) contribute log 1 = 0
            double v = 0.0;
            for (int j = tid; j < GSUM_TILE; j += 32)
                if (k * GSUM_TILE + j < P.n) v += log(dg[j]);
            v = warp_sum(v);
            if (tid == 0) P.logdet_part[(int64_t)b * P.T + k] = fail ? nan("") : 2.0 * v;
        }
        if (k + 1 < P.T) chain_bar_arrive(CH_BAR_FREE);
        double *tmp = Dk; Dk = Dn; Dn = tmp;
    }
}

// ---- the chain (warps 0-3) ---------------------------------------------------------------------------------------------
// st[0] cycles waiting for the pre tiles, st[1] POTRF, st[2] block inverses, st[3] solve, st[4] update, st[5] columns
__device__ __forceinline__ void ht_chain_worker(const HeteroArgs &D, double *smem, int b, long long *st) {
    const BorderedBatch &P = D.P;
    const int tid = EPI_TID, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    double *Dk = smem, *Pt = smem + CH_TILE_DOUBLES, *Dn = smem + 2 * CH_TILE_DOUBLES;
    double *dg = smem + 3 * CH_TILE_DOUBLES, *Dv = dg + 3 * GSUM_TILE;
    int *s_fail = reinterpret_cast<int *>(dg + 2 * GSUM_TILE);          // [0] failing column of this tile, [1] abort
    double *Ab = P.A + (int64_t)b * P.bstride;
    int *abort_flag = D.ctl + 1;
    int *frow = D.flags + (int64_t)b * P.Trows * P.T;
    const int *pre = D.pre + (int64_t)b * P.T;
    if (tid == 0) { s_fail[0] = 0; s_fail[1] = 0; }
    chain_load_tile(Dk, Ab, P.ld, tid);
    for (int k = 0; k < P.T; k++) {
        const bool more = (k + 1 < P.T);
        const double *C10 = Ab + (int64_t)(k + 1) * GSUM_TILE * P.ld + k * GSUM_TILE;
        // the helper has written the outputs of column k-1 (it read the buffer that is Dn now, diag(L) and the status), and
        // every chain thread is past the update of column k-1 (it read Pt)
        if (k > 0) chain_bar_sync(CH_BAR_FREE);
        // ---- prefetch the two pre tiles under the POTRF if they are ready -----------------------------------------------
        bool fetched = false;
        if (more) {
            int r = 1;
            if (k > 0) {
                if (tid == 0) {
                    r = (ld_relaxed(pre + k) >= 1 && ld_relaxed(frow + (int64_t)(k + 1) * P.T + k + 1) >= 1) ? 1 : 0;
                    if (r) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                }
                r = cons_sync_and(r != 0) ? 1 : 0;
            }
            if (r) {
                chain_load_tile(Pt, C10, P.ld, tid);
                chain_load_tile(Dn, C10 + GSUM_TILE, P.ld, tid);
                fetched = true;
            }
        }
        if (k == 0) { if (fetched) cp_async_wait<2>(); else cp_async_wait<0>(); }
        if (tid == 0) s_fail[0] = 0;
        CONS_SYNC();
        long long t0 = st ? clock64() : 0;
        tile_potrf_lean(Dk, dg, s_fail);
        long long t1 = st ? clock64() : 0;
        chain_invert_blocks(Dk, dg, Dv, s_fail[0] != 0);
        chain_bar_arrive(CH_BAR_POTRF);               // the helper takes the outputs from here
        if (!more) break;
        long long t2 = st ? clock64() : 0;
        if (!fetched) {
            int ok = 1;
            if (tid == 0) {
                ok = (flag_wait_ge(pre + k, 1, abort_flag) && flag_wait_ge(frow + (int64_t)(k + 1) * P.T + k + 1, 1, abort_flag)) ? 1 : 0;
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            if (!cons_sync_and(ok != 0)) {
                // aborting: release the helper (it leaves at its next barrier)
                if (tid == 0) s_fail[1] = 1;
                chain_bar_arrive(CH_BAR_POTRF);
                return;
            }
            chain_load_tile(Pt, C10, P.ld, tid);
            chain_load_tile(Dn, C10 + GSUM_TILE, P.ld, tid);
        }
        cp_async_wait<1>();                           // the sub-diagonal tile has landed (the diagonal one may still be in flight)
        CONS_SYNC();                                  // ... for every thread, and the block inverses are complete
        long long t3 = st ? clock64() : 0;
        // ---- L_{k+1,k} = S' L_kk^{-T}: warp w owns rows 16 w .. 16 w + 15 ------------------------------------------------
        {
            Acc acc;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const double2 v = *reinterpret_cast<const double2 *>(Pt + (w * 16 + mt * 8 + g) * GSUM_LDS + nt * 8 + 2 * t);
                    acc[mt][nt][0] = v.x; acc[mt][nt][1] = v.y;
                }
            chain_trsm<2>(acc, Dk, Dv, g, t);
            double *C = Ab + (int64_t)(k + 1) * GSUM_TILE * P.ld + k * GSUM_TILE;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                    *reinterpret_cast<double2 *>(C + (int64_t)(w * 16 + mt * 8 + g) * P.ld + nt * 8 + 2 * t) = v;
                    *reinterpret_cast<double2 *>(Pt + (w * 16 + mt * 8 + g) * GSUM_LDS + nt * 8 + 2 * t) = v;
                }
        }
        cp_async_wait<0>();
        CONS_SYNC();                                  // X complete in Pt, Dn landed, the tile stores ordered before the release
        if (tid == 0) st_release(frow + (int64_t)(k + 1) * P.T + k, 1);
        long long t4 = st ? clock64() : 0;
        chain_syrk(Dn, Pt, w, g, t);
        if (st && tid == 0) { st[0] += t3 - t2; st[1] += t1 - t0; st[2] += t2 - t1; st[3] += t4 - t3; st[4] += clock64() - t4; st[5] += 1; }
        double *tmp = Dk; Dk = Dn; Dn = tmp;
    }
}
