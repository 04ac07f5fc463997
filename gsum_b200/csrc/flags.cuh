// Synchronisation primitives of the persistent factorisation kernels (hetero.cuh, hetero_tma.cuh, chain.cuh): mbarrier
// wrappers, relaxed / release flag accesses at GPU scope, bounded waits (every wait is covered by a watchdog that turns a
// protocol error into an error code instead of a hung GPU), and the host-built task order.
#pragma once
#include "chol.cuh"

#define DF_WATCHDOG_CYCLES (4000000000LL)       // ~2 s at 1.9 GHz



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

// After the run: a watchdog abort marks every matrix as failed so that no caller consumes half-factored data.
__global__ void df_check_kernel(const int *abort_flag, int *info, int64_t batch, int *sticky) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (*abort_flag == 0) return;
    if (b == 0 && sticky) *sticky = 1;
    if (info && b < batch) info[b] = 0x7fffffff;
}

#include <vector>
#include <algorithm>
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
