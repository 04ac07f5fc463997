// Persistent dataflow schedule of the bordered Cholesky (K2 + K3 in ONE launch).
//
// Every tile task (i, k, b) of the whole batch sits in one host-built list, ordered so that a task's dependencies
// always precede it; CTAs (two per SM, all co-resident: cooperative launch) claim tasks in list order from a global
// counter and synchronise through per-tile "done" flags in global memory — no kernel boundary between tile columns,
// so one matrix's latency-bound diagonal tile overlaps the DMMA-bound panel tiles of all the others.
//
// CTA = 4 math warps + 1 producer warp.  The producer waits on the dependency flags (acquire), then streams the
// operand half-slabs with 16-byte async copies (cp.async.cg -> SASS LDGSTS, L2 only) into a 3-stage shared-memory ring
// guarded by full/empty mbarriers (completion via cp.async.mbarrier.arrive); the math warps never issue a global load
// for operands and never spin on a flag.  The epilogue's L_kk rides through the same ring as the last stage.
// (Measured, profiles/r01_notes.md: feeding the ring with 1-D bulk copies of one 256-byte row each — cp.async.bulk /
// UBLKCP — costs ~73 cycles per request on the TMA unit and starves the math warps 50 % of the time; row-sized
// requests are too small for the TMA engine, and the padded, bank-conflict-free smem layout rules out tiled tensor maps.)
//
// Deadlock freedom: tasks are claimed in list order and a task only waits for tasks earlier in the list, which are
// therefore already claimed by resident CTAs; the earliest unfinished task never waits.  Every wait loop is bounded by
// a watchdog (abort flag) so a logic error surfaces as an error code, not as a hung GPU.
#pragma once
#include "chol.cuh"

#define DF_THREADS 160
#define DF_PRODUCER_WARP 4
#define DF_NSTAT 24
#define DF_WATCHDOG_CYCLES (4000000000LL)       // ~2 s at 1.9 GHz

struct DataflowArgs {
    BorderedBatch P;
    const int4 *tasks;      // (i, k, b, flags) in schedule order; flags bit 0: thin border task (<= 8 rows in use)
    int ntasks;
    int *counter;           // next task to claim (zeroed before launch)
    int *flags;             // done flag per (b, i, k): index (b * Trows + i) * T + k  (zeroed before launch)
    int *abort_flag;        // set by any thread whose wait exceeded the watchdog
    long long *stats;       // optional per-CTA cycle counters [grid][DF_NSTAT] (dev instrumentation; nullptr = off)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-charged at init: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Polling load: L2-coherent, no L1 invalidate (ld.acquire.gpu compiles to LD + CCTL.IVALL, ~1000 cycles a poll).
// Enough for the flags: whatever a set flag guards is read afterwards with cp.async.cg / from L2, never through L1.
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Bounded mbarrier wait; returns false if the kernel is aborting.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, unsigned parity, int *abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    int spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023) == 0) {
            if (ld_relaxed(abort_flag)) return false;
            // the intra-CTA handshakes are covered by the same watchdog as the flags: a protocol error must surface as an
            // error code, never as a hung GPU
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > DF_WATCHDOG_CYCLES) { atomicExch(abort_flag, 1); return false; }
        }
    }
    return true;
}
// Bounded wait for a done flag (one lane polls).  Returns false on abort / watchdog.
__device__ __forceinline__ bool flag_wait(const int *flag, int *abort_flag) {
    if (ld_relaxed(flag)) return true;
    const long long t0 = clock64();
    while (!ld_relaxed(flag)) {
        __nanosleep(64);
        if (ld_relaxed(abort_flag)) return false;
        if (clock64() - t0 > DF_WATCHDOG_CYCLES) { atomicExch(abort_flag, 1); return false; }
    }
    return true;
}

// AND-reduction + barrier over one 128-thread group (named barrier 8 + group index)
__device__ __forceinline__ bool cons_sync_and(bool v) {
    unsigned r;
    __syncwarp();                       // aligned barrier: the warp arrives converged (see CONS_SYNC)
    asm volatile("{\n .reg .pred p, q;\n setp.ne.u32 q, %1, 0;\n bar.red.and.pred p, %2, 128, q;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(r) : "r"((unsigned)v), "r"(8 + (int)(threadIdx.x >> 7)) : "memory");
    return r != 0;
}

struct RingState { int stage; unsigned phase; };
__device__ __forceinline__ void ring_advance(RingState &r) {
    if (++r.stage == CHOL_NST) { r.stage = 0; r.phase ^= 1u; }
}

template <bool STATS>
__global__ void __launch_bounds__(DF_THREADS, CHOL_CTAS_PER_SM) chol_dataflow_kernel(DataflowArgs D) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t full_bar[CHOL_NST], empty_bar[CHOL_NST];
    __shared__ int s_task;
    const BorderedBatch &P = D.P;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < CHOL_NST; s++) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], CHOL_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    RingState ring = {0, 0u};           // advanced identically by the producer and the math warps
    long long st_wait_full = 0, st_wait_flag = 0, st_wait_empty = 0, st_epi = 0, st_ntask = 0, st_acc0 = 0, st_fence = 0, st_claim = 0;
    long long es[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long st_t0 = STATS ? clock64() : 0;
    const bool st_on = STATS && D.stats != nullptr;
    bool alive = true;
    for (;;) {
        const long long tc0 = st_on ? clock64() : 0;
        if (tid == 0) s_task = atomicAdd(D.counter, 1);
        __syncthreads();
        if (st_on) st_claim += clock64() - tc0;
        const int tix = s_task;
        if (tix >= D.ntasks || !alive) break;
        const int4 tk = D.tasks[tix];
        const int i = tk.x, k = tk.y, b = tk.z;
        const bool diag = (i == k);
        const bool thin = (tk.w & 1) != 0;          // border tile row with <= 8 rows in use: 8 x 64 task
        double *Ab = P.A + (int64_t)b * P.bstride;
        double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                               : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
        const double *Ak = Ab + (int64_t)k * GSUM_TILE * P.ld;
        const int nh = 2 * k;
        const int *frow_i = D.flags + ((int64_t)b * P.Trows + i) * P.T;
        const int *frow_k = D.flags + ((int64_t)b * P.Trows + k) * P.T;

        if (w == DF_PRODUCER_WARP) {
            // ======================= producer: dependency flags -> bulk copies into the ring =======================
            for (int h = 0; h < nh && alive; h++) {
                const int j = h >> 1;
                if ((h & 1) == 0) {               // new slab j: tiles (i, j) and (k, j) must be final
                    int ok = 1;
                    long long tq = st_on ? clock64() : 0;
                    if (lane == 0) ok = flag_wait(frow_i + j, D.abort_flag) && (diag || flag_wait(frow_k + j, D.abort_flag));
                    if (st_on) st_wait_flag += clock64() - tq;
                    alive = __shfl_sync(0xffffffffu, ok, 0) != 0;
                    if (!alive) break;
                }
                if (lane == 0) {
                    long long tq = st_on ? clock64() : 0;
                    alive = mbar_wait(&empty_bar[ring.stage], ring.phase ^ 1u, D.abort_flag);
                    if (st_on) st_wait_empty += clock64() - tq;
                }
                alive = __shfl_sync(0xffffffffu, (int)alive, 0) != 0;
                if (!alive) break;
                double *As = smem + ring.stage * CHOL_STAGE_DOUBLES, *Bs = As + GSUM_TILE * GSUM_LDH;
                const int col0 = j * GSUM_TILE + (h & 1) * GSUM_KH;
                // 64 rows x 256 B per operand = 1024 16-byte chunks; a warp-wide LDGSTS moves two rows
                if (thin) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int c = lane + 32 * q, row = c >> 4, ch = (c & 15) * 2;
                        cp_async16(As + row * GSUM_LDH + ch, Ri + (int64_t)row * P.ld + col0 + ch);
                    }
                } else {
#pragma unroll 8
                    for (int q = 0; q < 32; q++) {
                        const int c = lane + 32 * q, row = c >> 4, ch = (c & 15) * 2;
                        cp_async16(As + row * GSUM_LDH + ch, Ri + (int64_t)row * P.ld + col0 + ch);
                    }
                }
                if (!diag) {
#pragma unroll 8
                    for (int q = 0; q < 32; q++) {
                        const int c = lane + 32 * q, row = c >> 4, ch = (c & 15) * 2;
                        cp_async16(Bs + row * GSUM_LDH + ch, Ak + (int64_t)row * P.ld + col0 + ch);
                    }
                }
                cp_async_mbar_arrive(&full_bar[ring.stage]);
                ring_advance(ring);
            }
            if (!diag && alive) {                 // tail: L_kk for the triangular solve
                int ok = 1;
                if (lane == 0) {
                    ok = flag_wait(frow_k + k, D.abort_flag);
                    if (ok) ok = mbar_wait(&empty_bar[ring.stage], ring.phase ^ 1u, D.abort_flag);
                }
                alive = __shfl_sync(0xffffffffu, ok, 0) != 0;
                if (alive) {
                    double *Ls = smem + ring.stage * CHOL_STAGE_DOUBLES;
                    const double *Lg = Ak + k * GSUM_TILE;
#pragma unroll 8
                    for (int q = 0; q < 64; q++) {
                        const int c = lane + 32 * q, row = c >> 5, ch = (c & 31) * 2;
                        cp_async16(Ls + row * GSUM_LDS + ch, Lg + (int64_t)row * P.ld + ch);
                    }
                    cp_async_mbar_arrive(&full_bar[ring.stage]);
                    ring_advance(ring);
                }
            }
        } else {
            // ======================= math warps: DMMA main loop + epilogue ==========================================
            const int g = lane >> 2, t = lane & 3;
            const int ntm = diag ? 2 * (w + 1) : 8;            // diagonal tile: warp w owns columns < 16 (w + 1)
            double *C = Ri + k * GSUM_TILE;
            Acc acc;                                            // thin task: acc[k-step parity][n tile 0..1] only
            long long tq0 = st_on ? clock64() : 0;
            if (!thin) tile_load_acc(acc, C, P.ld, ntm);
            else {
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    const double2 v = *reinterpret_cast<const double2 *>(C + (int64_t)g * P.ld + w * 16 + nt * 8 + 2 * t);
                    acc[0][nt][0] = v.x; acc[0][nt][1] = v.y; acc[1][nt][0] = 0.0; acc[1][nt][1] = 0.0;
                }
            }
            if (st_on) { st_acc0 += clock64() - tq0; st_ntask++; }
            for (int h = 0; h < nh && alive; h++) {
                long long tq = st_on ? clock64() : 0;
                alive = mbar_wait(&full_bar[ring.stage], ring.phase, D.abort_flag);
                if (st_on) st_wait_full += clock64() - tq;
                if (!alive) break;
                const double *As = smem + ring.stage * CHOL_STAGE_DOUBLES;
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
            double *Lk = nullptr;
            int tail_stage = -1;
            if (!diag && alive) {
                long long tq = st_on ? clock64() : 0;
                alive = mbar_wait(&full_bar[ring.stage], ring.phase, D.abort_flag);
                if (st_on) st_wait_full += clock64() - tq;
                Lk = smem + ring.stage * CHOL_STAGE_DOUBLES;
                tail_stage = ring.stage;
                ring_advance(ring);
            }
            // the block must agree on `alive` before the barriers inside the epilogue
            alive = cons_sync_and(alive);
            if (alive) {
                // every filled stage has been consumed, so any buffer but the tail's is free for the epilogue's scratch
                const int s_stage = (tail_stage + 1 + (tail_stage < 0 ? 1 : 0)) % CHOL_NST;
                double *S = smem + s_stage * CHOL_STAGE_DOUBLES;
                long long tq = st_on ? clock64() : 0;
                if (!thin) tile_epilogue(P, i, k, b, acc, S, Lk, C, st_on ? es : nullptr);
                else {
                    // 8 x 64 tile: gather the four 8 x 16 pieces in smem, then warp 0 solves the rows in registers
                    double *rdiag = S + 8 * GSUM_LDS, *Lp = rdiag + GSUM_TILE, *scr = Lp + 512;       // behind the 8 staged rows
#pragma unroll
                    for (int nt = 0; nt < 2; nt++) {
                        double2 v; v.x = acc[0][nt][0] + acc[1][nt][0]; v.y = acc[0][nt][1] + acc[1][nt][1];
                        *reinterpret_cast<double2 *>(S + g * GSUM_LDS + w * 16 + nt * 8 + 2 * t) = v;
                    }
                    trsm_prepare(Lk, Lp, rdiag);
                    CONS_SYNC();
                    if (w == 0) {
#pragma unroll
                        for (int nt = 0; nt < 8; nt++) {
                            const double2 v = *reinterpret_cast<const double2 *>(S + g * GSUM_LDS + nt * 8 + 2 * t);
                            acc[0][nt][0] = v.x; acc[0][nt][1] = v.y; acc[1][nt][0] = 0.0; acc[1][nt][1] = 0.0;
                        }
                        trsm_rows<1>(acc, Lk, Lp, rdiag, scr);
#pragma unroll
                        for (int nt = 0; nt < 8; nt++) {
                            double2 v; v.x = acc[0][nt][0]; v.y = acc[0][nt][1];
                            *reinterpret_cast<double2 *>(C + (int64_t)g * P.ld + nt * 8 + 2 * t) = v;
                        }
                    }
                }
                if (st_on) st_epi += clock64() - tq;
                const long long tf = st_on ? clock64() : 0;
                __threadfence();                              // tile stores visible device-wide before the flag
                CONS_SYNC();
                if (tid == 0) st_release(D.flags + ((int64_t)b * P.Trows + i) * P.T + k, 1);
                if (st_on) st_fence += clock64() - tf;
                if (tail_stage >= 0 && lane == 0) mbar_arrive(&empty_bar[tail_stage]);     // release the tail's stage
            }
        }
        // both roles learn whether anyone aborted; also separates tasks
        alive = __syncthreads_and(alive ? 1 : 0) != 0;
        if (!alive) break;
    }
    if (st_on && (tid == 0 || tid == DF_PRODUCER_WARP * 32)) {
        long long *o = D.stats + (int64_t)blockIdx.x * DF_NSTAT;
        if (tid == 0) {
            o[0] = clock64() - st_t0; o[1] = st_wait_full; o[2] = st_epi; o[3] = st_ntask; o[4] = st_acc0; o[7] = st_fence; o[8] = st_claim;
            for (int q = 0; q < 8; q++) o[9 + q] = es[q];
        } else { o[5] = st_wait_flag; o[6] = st_wait_empty; }
    }
}


// ---- host side ------------------------------------------------------------------------------------------------------
__global__ void df_init_kernel(int *flags, int *counter_abort, int64_t batch, int Trows, int T, int factor_done) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2) counter_abort[idx] = 0;
    if (idx >= batch * Trows * T) return;
    const int i = (int)((idx / T) % Trows);
    flags[idx] = (factor_done && i < T) ? 1 : 0;
}
// After the run: a watchdog abort marks every matrix as failed so that no caller consumes half-factored data.
__global__ void df_check_kernel(const int *abort_flag, int *info, int64_t batch, int *sticky) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (*abort_flag == 0) return;
    if (b == 0 && sticky) *sticky = 1;
    if (info && b < batch) info[b] = 0x7fffffff;
}

#include <vector>
// Task list in dependency order with one column of look-ahead.  Column k: first the sub-diagonal tiles (k+1, k, b) of
// every matrix — the only fresh operand of the next diagonal tile — then `DF_DIAG_DELAY` of the other tiles of the
// column, then the diagonal tiles (k+1, k+1, b), then the rest.  The diagonal tiles thus start as early as they can
// without stalling their CTA on the flag of a tile that is still being computed (the POTRF chain is the critical path).
#ifndef DF_DIAG_DELAY
#define DF_DIAG_DELAY 1184
#endif
static inline void df_build_tasks(std::vector<int4> &out, int T, int Trows, int batch, bool solve_only, bool thin_last, int diag_delay = DF_DIAG_DELAY) {
    out.clear();
    auto flags = [&](int i) { return (thin_last && i == Trows - 1 && i >= T) ? 1 : 0; };
    if (solve_only) {
        for (int k = 0; k < T; k++)
            for (int i = T; i < Trows; i++)
                for (int b = 0; b < batch; b++) out.push_back(make_int4(i, k, b, 0 | flags(i)));
        return;
    }
    for (int k = 0; k < T; k++) {
        if (k == 0) for (int b = 0; b < batch; b++) out.push_back(make_int4(0, 0, b, 0));
        if (k + 1 < Trows) for (int b = 0; b < batch; b++) out.push_back(make_int4(k + 1, k, b, flags(k + 1)));
        int emitted = 0;
        bool diag_done = !(k + 1 < T);
        auto emit_diag = [&]() { for (int b = 0; b < batch; b++) out.push_back(make_int4(k + 1, k + 1, b, 0)); diag_done = true; };
        for (int i = k + 2; i < Trows; i++)
            for (int b = 0; b < batch; b++) {
                if (!diag_done && emitted >= diag_delay) emit_diag();
                out.push_back(make_int4(i, k, b, flags(i)));
                emitted++;
            }
        if (!diag_done) emit_diag();
    }
}
