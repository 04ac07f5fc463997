// Warp-specialised pipeline schedule of the bordered Cholesky (K2 + K3 in ONE launch, one persistent CTA per SM).
//
// Same task list, done flags and dependency rule as dataflow.cuh, but the three phases of a tile task run on different
// warps of the CTA and overlap across consecutive tasks:
//
//   4 producer warps claim tasks in list order (global counter), publish them through a small shared-memory queue,
//                    wait on the dependency flags and stream the operand half-slabs into the mbarrier ring
//                    (cp.async.cg, 16 B) — they run ahead of the math warps by the depth of the ring, across task
//                    boundaries, so a task's first operands are already in flight while the previous one computes;
//   4 math warps     (one per SM sub-partition) run the DMMA main loops back to back, starting from a zero accumulator,
//                    and drop the finished 64x64 product P into one of two shared-memory buffers;
//   4 epilogue warps prefetch the C tile into registers, form S = C - P, and finish the task: POTRF (i == k) or the
//                    triangular solve against L_kk (i > k, L_kk staged by the group itself), tile store, flag.
//
// Why: measured on B200 (tools/fp64_latency.cu, tools/epi_bench.cu) a dependent DFMA takes 32 cycles (275 while another
// warp streams DMMAs on the same sub-partition), so the epilogues are long latency chains (POTRF 64x64 ~24k cycles,
// TRSM ~6k, more under contention) during which a CTA that does everything in the same warps leaves the FP64 tensor
// pipe idle.  Decoupling keeps the math warps issuing DMMAs while the chains of the previous tasks drain.
//
// Deadlock freedom as in dataflow.cuh: tasks are claimed in list order, every role handles its CTA's tasks in that
// order, and a task only waits for tasks earlier in the list; every wait is bounded by the watchdog / abort flag.
#pragma once
#include "dataflow.cuh"

#define PL_THREADS 384                  // warpgroups: 4 math warps | 4 epilogue warps | 4 producer warps
#define PL_PRODUCER_WARP 8
#define PL_NST 3                        // operand ring stages (K depth 32 each)
#ifndef PL_QD
#define PL_QD 4                         // task queue depth
#endif
#define PL_P_DOUBLES 4096               // one product buffer (fragment-major, 32 KiB)
#define PL_LKS_DOUBLES 4608             // L_kk (panel) / the tile being factored (diagonal) + diag / fail scratch
#define PL_SCR_DOUBLES 1280             // rdiag, prescaled diagonal blocks, per-warp TRSM scratch
#define PL_SMEM_DOUBLES (PL_NST * CHOL_STAGE_DOUBLES + 2 * PL_P_DOUBLES + PL_LKS_DOUBLES + PL_SCR_DOUBLES)
#define PL_SMEM_BYTES (PL_SMEM_DOUBLES * 8)

template <bool STATS>
__global__ void __launch_bounds__(PL_THREADS, 1) chol_pipeline_kernel(DataflowArgs D) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t full_bar[PL_NST], empty_bar[PL_NST], tq_full[PL_QD], tq_empty[PL_QD], p_full[2], p_empty[2];
    __shared__ int4 tq[PL_QD];
    const BorderedBatch &P = D.P;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool st_on = STATS && D.stats != nullptr;
    long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long st_t0 = STATS ? clock64() : 0;
#define PL_T0() const long long _t = st_on ? clock64() : 0
#define PL_ACC(q) do { if (st_on) st[q] += clock64() - _t; } while (0)
    double *ring_base = smem;
    double *pbuf = smem + PL_NST * CHOL_STAGE_DOUBLES;
    double *LkS = pbuf + 2 * PL_P_DOUBLES;
    double *scr = LkS + PL_LKS_DOUBLES;
    if (tid == 0) {
        for (int s = 0; s < PL_NST; s++) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 4); }
        for (int s = 0; s < PL_QD; s++) { mbar_init(&tq_full[s], 1); mbar_init(&tq_empty[s], 11); }
        for (int s = 0; s < 2; s++) { mbar_init(&p_full[s], 4); mbar_init(&p_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Registers are allocated per warpgroup: 12 warps at the launch bound get 168 each, which the math and epilogue
    // warps overrun.  The producer warpgroup hands half of its share back (not more: with 40 registers its copies
    // serialise on address-register reuse).
    if (w >= PL_PRODUCER_WARP) {
#ifndef PL_NO_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
#endif
        // ============================ producer warps =========================================================
        // One warp cannot keep enough 16-byte copies in flight to feed the ring (measured: the math warps waited 44 %
        // of the time on a single producer), so all four warps of the warpgroup stream: warp pw moves rows
        // [16 pw, 16 pw + 16) of both operand half-slabs.  Warp 0 of the group also claims the tasks.
        const int pw = w - PL_PRODUCER_WARP;
        RingState ring = {0, 0u};
        for (int n = 0;; n++) {
            const int slot = n % PL_QD;
            int4 tk = make_int4(-1, 0, 0, 0);
            int ok = 1;
            if (pw == 0) {
                if (lane == 0) {
                    { PL_T0(); ok = mbar_wait(&tq_empty[slot], (((unsigned)(n / PL_QD)) & 1u) ^ 1u, D.abort_flag); PL_ACC(0); }
                    if (ok) {
                        const int tix = atomicAdd(D.counter, 1);
                        if (tix < D.ntasks) tk = D.tasks[tix];
                        tq[slot] = tk;
                        mbar_arrive(&tq_full[slot]);
                    }
                }
                ok = __shfl_sync(0xffffffffu, ok, 0);
                tk.x = __shfl_sync(0xffffffffu, tk.x, 0); tk.y = __shfl_sync(0xffffffffu, tk.y, 0);
                tk.z = __shfl_sync(0xffffffffu, tk.z, 0); tk.w = __shfl_sync(0xffffffffu, tk.w, 0);
            } else {
                ok = mbar_wait(&tq_full[slot], ((unsigned)(n / PL_QD)) & 1u, D.abort_flag);
                if (ok) {
                    tk = tq[slot];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tq_empty[slot]);
                }
            }
            if (!ok || tk.x < 0) break;
            const int i = tk.x, k = tk.y, b = tk.z;
            const bool diag = (i == k), thin = (tk.w & 1) != 0;
            const double *Ab = P.A + (int64_t)b * P.bstride;
            const double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                                         : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
            const double *Ak = Ab + (int64_t)k * GSUM_TILE * P.ld;
            const int *frow_i = D.flags + ((int64_t)b * P.Trows + i) * P.T;
            const int *frow_k = D.flags + ((int64_t)b * P.Trows + k) * P.T;
            // this lane's chunk within a 16-row x 256-byte quarter: 8 chunks per lane, two rows per warp-wide copy
            const int r0 = pw * 16 + (lane >> 4), ch = (lane & 15) * 2;
            bool alive = true;
            // A finished tile (r, k-1) implies every (r, j < k-1): they were its operands.  One look at the last flag of
            // each operand row usually clears the whole task; only tasks on the critical path poll slab by slab.
            int done_i = 1, done_k = 1;
            if (lane == 0 && k > 0) {
                done_i = ld_relaxed(frow_i + k - 1);
                done_k = diag ? done_i : ld_relaxed(frow_k + k - 1);
            }
            for (int h = 0; h < 2 * k && alive; h++) {
                const int j = h >> 1;
                int good = 1;
                if (lane == 0) {
                    if ((h & 1) == 0 && !(done_i && done_k)) {
                        PL_T0();
                        good = (done_i || flag_wait(frow_i + j, D.abort_flag)) && (diag || done_k || flag_wait(frow_k + j, D.abort_flag));
                        PL_ACC(1);
                    }
                    if (good) { PL_T0(); good = mbar_wait(&empty_bar[ring.stage], ring.phase ^ 1u, D.abort_flag); PL_ACC(2); }
                }
                alive = __shfl_sync(0xffffffffu, good, 0) != 0;
                if (!alive) break;
                double *As = ring_base + ring.stage * CHOL_STAGE_DOUBLES, *Bs = As + GSUM_TILE * GSUM_LDH;
                const int col0 = j * GSUM_TILE + (h & 1) * GSUM_KH;
                if (!thin) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int row = r0 + 2 * q;
                        cp_async16(As + row * GSUM_LDH + ch, Ri + (int64_t)row * P.ld + col0 + ch);
                    }
                } else if (pw == 0) {                      // thin task: rows 0..7 of the A operand only
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int row = (lane >> 4) + 2 * q;
                        cp_async16(As + row * GSUM_LDH + ch, Ri + (int64_t)row * P.ld + col0 + ch);
                    }
                }
                if (!diag) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int row = r0 + 2 * q;
                        cp_async16(Bs + row * GSUM_LDH + ch, Ak + (int64_t)row * P.ld + col0 + ch);
                    }
                }
                cp_async_mbar_arrive(&full_bar[ring.stage]);
                ring_advance(ring);
            }
            if (!alive) break;
        }
        cp_async_wait<0>();
        if (st_on && pw == 0 && lane == 0) { long long *o = D.stats + (int64_t)blockIdx.x * DF_NSTAT; o[0] = clock64() - st_t0; o[1] = st[0]; o[2] = st[1]; o[3] = st[2]; }
    } else if (w < 4) {
#ifndef PL_NO_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#endif
        // ============================ math warps ============================================================
        const int g = lane >> 2, t = lane & 3;
        RingState ring = {0, 0u};
        int pcount = 0;
        for (int n = 0;; n++) {
            const int slot = n % PL_QD;
            { PL_T0(); const bool okq = mbar_wait(&tq_full[slot], ((unsigned)(n / PL_QD)) & 1u, D.abort_flag); PL_ACC(0); if (!okq) break; }
            const int4 tk = tq[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[slot]);
            if (tk.x < 0) break;
            const int i = tk.x, k = tk.y;
            const bool diag = (i == k), thin = (tk.w & 1) != 0;
            if (k == 0) continue;                                 // nothing to accumulate: the epilogue works on the C tile alone
            const int ntm = diag ? 2 * (w + 1) : 8;
            Acc acc;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
            bool alive = true;
            for (int h = 0; h < 2 * k; h++) {
                { PL_T0(); alive = mbar_wait(&full_bar[ring.stage], ring.phase, D.abort_flag); PL_ACC(1); if (thin) PL_ACC(5); else if (h == 0) PL_ACC(4); else if (diag) PL_ACC(6); }
                if (!alive) break;
                const double *As = ring_base + ring.stage * CHOL_STAGE_DOUBLES;
                const double *Bs = diag ? As : As + GSUM_TILE * GSUM_LDH;
                if (thin) {
                    const double *ap = As + g * GSUM_LDH + t;
                    const double *bp = Bs + (w * 16 + g) * GSUM_LDH + t;
#pragma unroll
                    for (int ks = 0; ks < GSUM_KH / 4; ks++) {
                        const double a = -ap[ks * 4];
#pragma unroll
                        for (int nt = 0; nt < 2; nt++)
                            dmma884(acc[ks & 1][nt][0], acc[ks & 1][nt][1], a, bp[nt * 8 * GSUM_LDH + ks * 4]);
                    }
                } else if (diag) stage_mma<false>(acc, As, Bs, ntm);
                else stage_mma<true>(acc, As, Bs, 8);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[ring.stage]);
                ring_advance(ring);
            }
            if (!alive) break;
            // hand the (negated) product over:  acc = -sum A B^T
            const int pb = pcount & 1;
            { PL_T0(); const bool okp = mbar_wait(&p_empty[pb], (((unsigned)(pcount >> 1)) & 1u) ^ 1u, D.abort_flag); PL_ACC(2); if (!okp) break; }
            st[3] += 1;
            double *Pb = pbuf + pb * PL_P_DOUBLES;
            if (thin) {
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    double2 v; v.x = acc[0][nt][0] + acc[1][nt][0]; v.y = acc[0][nt][1] + acc[1][nt][1];
                    *reinterpret_cast<double2 *>(Pb + g * GSUM_LDS + w * 16 + nt * 8 + 2 * t) = v;
                }
            } else {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                        *reinterpret_cast<double2 *>(Pb + (((w * 16 + mt * 8 + nt) * 32) + lane) * 2) = v;
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[pb]);
            pcount++;
        }
        if (st_on && tid == 0) { long long *o = D.stats + (int64_t)blockIdx.x * DF_NSTAT; o[4] = clock64() - st_t0; o[5] = st[0]; o[6] = st[1]; o[7] = st[2]; o[8] = st[3]; o[18] = st[4]; o[19] = st[5]; o[20] = st[6]; }
    } else {
#ifndef PL_NO_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#endif
        // ============================ epilogue warps ========================================================
        const int etid = EPI_TID, ew = etid >> 5, g = lane >> 2, t = lane & 3;
        int pcount = 0;
        for (int n = 0;; n++) {
            const int slot = n % PL_QD;
            bool alive;
            { PL_T0(); alive = mbar_wait(&tq_full[slot], ((unsigned)(n / PL_QD)) & 1u, D.abort_flag); PL_ACC(0); }
            int4 tk = make_int4(-1, 0, 0, 0);
            if (alive) {
                tk = tq[slot];
                __syncwarp();
                if (lane == 0) mbar_arrive(&tq_empty[slot]);
            }
            alive = cons_sync_and(alive);
            if (!alive || tk.x < 0) break;
            const int i = tk.x, k = tk.y, b = tk.z;
            const bool diag = (i == k), thin = (tk.w & 1) != 0;
            double *Ab = P.A + (int64_t)b * P.bstride;
            double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                                   : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
            double *C = Ri + k * GSUM_TILE;
            const double *Lg = Ab + (int64_t)k * GSUM_TILE * P.ld + k * GSUM_TILE;
            // 1. the C tile (original data: a tile is written exactly once, by its own task) into C fragments
            Acc T;
            if (!thin) tile_load_acc(T, C, P.ld, diag ? 2 * (ew + 1) : 8);
            else if (ew == 0) {
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const double2 v = *reinterpret_cast<const double2 *>(C + (int64_t)g * P.ld + nt * 8 + 2 * t);
                    T[0][nt][0] = v.x; T[0][nt][1] = v.y; T[1][nt][0] = 0.0; T[1][nt][1] = 0.0;
                }
            }
            // 2. L_kk for the triangular solve
            if (!diag) {
                int ok = 1;
                PL_T0();
                if (etid == 0) ok = flag_wait(D.flags + ((int64_t)b * P.Trows + k) * P.T + k, D.abort_flag);
                if (!cons_sync_and(ok != 0)) break;
                PL_ACC(1);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int c = etid + q * 128, row = c >> 5, ch = (c & 31) * 2;
                    cp_async16(LkS + row * GSUM_LDS + ch, Lg + (int64_t)row * P.ld + ch);
                }
                cp_async_commit();
            }
            // 3. S = C - sum A B^T
            if (k > 0) {
                const int pb = pcount & 1;
                { PL_T0(); alive = mbar_wait(&p_full[pb], ((unsigned)(pcount >> 1)) & 1u, D.abort_flag); PL_ACC(2); }
                if (alive) {
                    const double *Pb = pbuf + pb * PL_P_DOUBLES;
                    if (!thin) {
                        const int ntm = diag ? 2 * (ew + 1) : 8;
#pragma unroll
                        for (int mt = 0; mt < 2; mt++)
#pragma unroll
                            for (int nt = 0; nt < 8; nt++)
                                if (nt < ntm) {
                                    const double2 v = *reinterpret_cast<const double2 *>(Pb + (((ew * 16 + mt * 8 + nt) * 32) + lane) * 2);
                                    T[mt][nt][0] += v.x; T[mt][nt][1] += v.y;
                                }
                    } else if (ew == 0) {
#pragma unroll
                        for (int nt = 0; nt < 8; nt++) {
                            const double2 v = *reinterpret_cast<const double2 *>(Pb + g * GSUM_LDS + nt * 8 + 2 * t);
                            T[0][nt][0] += v.x; T[0][nt][1] += v.y;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_empty[pb]);
                }
                pcount++;
            }
            if (!diag) cp_async_wait<0>();
            if (!cons_sync_and(alive)) break;
            // 4. finish the tile
            PL_T0();
            if (!thin) tile_epilogue<true>(P, i, k, b, T, diag ? LkS : scr, LkS, C);
            else {
                double *rdiag = scr, *Lp = scr + GSUM_TILE, *wscr = scr + GSUM_TILE + 512;
                trsm_prepare(LkS, Lp, rdiag);
                CONS_SYNC();
                if (ew == 0) {
                    trsm_rows<1>(T, LkS, Lp, rdiag, wscr);
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        double2 v; v.x = T[0][nt][0]; v.y = T[0][nt][1];
                        *reinterpret_cast<double2 *>(C + (int64_t)g * P.ld + nt * 8 + 2 * t) = v;
                    }
                }
            }
            PL_ACC(diag ? 3 : (thin ? 5 : 4));
            if (diag) st[6] += 1;
            { PL_T0();
            __threadfence();                              // tile stores visible device-wide before the flag
            CONS_SYNC();
            if (etid == 0) st_release(D.flags + ((int64_t)b * P.Trows + i) * P.T + k, 1);
            PL_ACC(7); }
        }
        if (st_on && etid == 0) {
            long long *o = D.stats + (int64_t)blockIdx.x * DF_NSTAT;
            o[9] = clock64() - st_t0;
            for (int q = 0; q < 8; q++) o[10 + q] = st[q];
        }
    }
#undef PL_T0
#undef PL_ACC
}
