// Shared device/host utilities for the gsum_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <nvtx3/nvToolsExt.h>

// NVTX range per C-ABI call and around the factorisation launch: visible in Nsight Systems / ncu --nvtx, free
// when no tool is attached (NVTX v3 is header-only and resolves its injection library lazily).
struct GsumRange {
    explicit GsumRange(const char *name) { nvtxRangePushA(name); }
    ~GsumRange() { nvtxRangePop(); }
};
#define GSUM_RANGE(name) GsumRange gsum_range_guard_(name)

#define GSUM_TILE 64            // tile edge of the blocked FP64 factorisation / solves
#define GSUM_LDS 68             // padded smem row stride (doubles) for 64-wide tiles: 68 % 16 == 4 -> conflict-free DMMA fragment loads
#define GSUM_KH 32              // K depth of one pipeline stage (half a tile)
#define GSUM_LDH 36             // padded smem row stride for 32-wide half-slabs: 36 % 16 == 4

#define GSUM_NWS 32             // workspace slots
struct gsum_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    char err[512];
    // grow-only device workspace arena (ctx-scoped; freed by gsum_ctx_destroy)
    void *ws[GSUM_NWS];
    size_t ws_bytes[GSUM_NWS];
    int64_t launches;           // kernels launched by this library on this context
    int *d_flag;                // small device int scratch
    // optional profiling of the factorisation phase (bench.py roofline): event pairs on ctx->stream
    int prof_enabled, prof_count;
    cudaEvent_t prof_ev[2 * 256];
    double prof_flops;          // algorithmic flops of the bracketed factorisations
    int64_t prof_border_rows;   // right-hand sides riding along with the current factorisation
    // factorisation schedule (hetero.cuh): cached claim lists and their key, tile flags, counters / sticky abort indicator
    void *df_flags; size_t df_flags_cap;
    int *df_ctl;                // [0] GEMM task counter, [1] abort flag, [2] sticky abort, [3], [4] factor task counters (device)
    int use_thin;               // GSUM_B200_THIN=0 disables the 8-row border tasks (debug / comparison)
    void *ht_gtasks, *ht_ftasks; size_t ht_gcap, ht_fcap; int ht_key[5]; int ht_ng, ht_nf;
    int hx_ready;               // function attributes of the factorisation kernel set
    // pinned staging arena for host-memory callers: small inputs go host -> pinned -> device with a truly asynchronous copy,
    // outputs come back device -> pinned and are handed to the caller after the call's single stream synchronisation
    char *pin; size_t pin_cap, pin_off;
    struct { void *dst; const void *src; size_t bytes; } pend[16];
    int npend;
    int ht_factor_ctas;         // GSUM_B200_FACTOR_CTAS (default HT_FACTOR_CTAS)
    int ht_factor_workers, ht_diag_delay;       // GSUM_B200_FACTOR_WORKERS, GSUM_B200_DIAG_DELAY (read at context creation)
    int use_smalln, sn_ready;   // small-N one-CTA grid path (smalln.cuh); GSUM_B200_SMALLN=0 disables
    int ht_chain_max;           // batches up to this size run in chain mode (chain.cuh); GSUM_B200_CHAIN_MAX, 0 disables
    // collective of the sharded grid (gsum_comm_init): an NCCL communicator owned by this context, NCCL resolved with dlopen
    void *comm; int comm_nranks, comm_rank;
};

// exp(a) for the RBF kernel's argument a = -d^2 / 2 <= 0, straight-line (K1 is bound by instruction issue: the library exp() with its
// branches was 48 of the 73 instructions per matrix entry; this one is ~25).  n = rint(a / ln 2) by the shift trick, Cody-Waite
// reduction with FMAs (|r| <= ln 2 / 2), Taylor polynomial of degree 13 by Horner (truncation 4e-18), scaling by 2^n through the
// exponent field — in two factors so that results below 2^-1022 round once, into the subnormal range or to zero exactly as exp()
// does.  Maximum error 0.85 ulp (CUDA's exp(): 1 ulp); against numpy's exp on 12 M arguments the product c * exp differs by at
// most one ulp (tools: tests/test_gpu_kernels.py::test_kernel_matrix keeps its 4e-16 bound).  NaN propagates; a > 0 is not supported.
__device__ __forceinline__ double rbf_exp_neg(double a) {
    const double ac = a < -750.0 ? -750.0 : a;                    // exp(-750) = 0 in FP64; keeps n inside the int range (NaN stays NaN)
    const double t = __fma_rn(ac, 1.4426950408889634, 6755399441055744.0);
    const int ni = __double2loint(t);
    const double n = t - 6755399441055744.0;
    double r = __fma_rn(n, -6.93147180369123816490e-01, ac);
    r = __fma_rn(n, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = __fma_rn(p, r, 1.0 / 479001600.0);
    p = __fma_rn(p, r, 1.0 / 39916800.0);
    p = __fma_rn(p, r, 1.0 / 3628800.0);
    p = __fma_rn(p, r, 1.0 / 362880.0);
    p = __fma_rn(p, r, 1.0 / 40320.0);
    p = __fma_rn(p, r, 1.0 / 5040.0);
    p = __fma_rn(p, r, 1.0 / 720.0);
    p = __fma_rn(p, r, 1.0 / 120.0);
    p = __fma_rn(p, r, 1.0 / 24.0);
    p = __fma_rn(p, r, 1.0 / 6.0);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    const int n1 = ni < -1021 ? -1021 : ni;                       // p in [0.70, 1.42): p * 2^n1 is a normal number
    const double big = __hiloint2double(__double2hiint(p) + (n1 << 20), __double2loint(p));
    const double rest = __hiloint2double((1023 + (ni - n1)) << 20, 0);      // 2^(ni - n1), 1.0 unless the result is subnormal
    const double v = big * rest;
    return a != a ? a : v;
}

static inline int gsum_fail(gsum_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        c->npend = 0; c->pin_off = 0;           // a failed call delivers nothing
        va_list ap; va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define GSUM_CUDA(ctx, call)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return gsum_fail((ctx), -100, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e),    \
                             __FILE__, __LINE__, cudaGetErrorString(_e));                         \
    } while (0)

#define GSUM_TRY(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != 0) return _rc;      \
    } while (0)

// Workspace slot `slot` of at least `bytes` (grow-only, contents undefined).
static inline int gsum_ws(gsum_ctx *c, int slot, size_t bytes, void **out) {
    if (c->ws_bytes[slot] < bytes) {
        if (c->ws[slot]) {
            GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
            GSUM_CUDA(c, cudaFree(c->ws[slot]));
            c->ws[slot] = nullptr; c->ws_bytes[slot] = 0;
        }
        size_t want = bytes + (bytes >> 3) + 256;
        GSUM_CUDA(c, cudaMalloc(&c->ws[slot], want));
        c->ws_bytes[slot] = want;
    }
    *out = c->ws[slot];
    return 0;
}

static inline int64_t gsum_pad64(int64_t n) { return (n + GSUM_TILE - 1) / GSUM_TILE * GSUM_TILE; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col)  — FP64 tensor-core MMA (SASS: DMMA.8x8x4).
// Fragment ownership for lane = 4*g + t:  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree order); result valid in every thread. `red` >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double *red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}

#endif  // __CUDACC__
