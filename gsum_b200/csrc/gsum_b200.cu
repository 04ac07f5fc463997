// C ABI of gsum_b200 (see include/gsum_b200.h).  Host-side orchestration only: every numerical step is a
// kernel from cov.cuh / chol.cuh / lml.cuh / solve.cuh / diag.cuh.  There is no CPU fallback.
#include "../../include/gsum_b200.h"
#include "common.cuh"
#include "cov.cuh"
#include "chol.cuh"
#include "lml.cuh"
#include "solve.cuh"
#include "diag.cuh"
#include "grad.cuh"
#include "pointwise.cuh"
#include "smalln.cuh"
#include "hetero.cuh"
#include "hetero_tma.cuh"
#include <cstdlib>
#include <vector>

#define LAUNCHED(ctx, n) ((ctx)->launches += (n))

enum {
    WS_X = 0, WS_XS, WS_DY, WS_REF, WS_ORD, WS_LS, WS_Q, WS_DETF, WS_MAT, WS_RHS, WS_GRAM, WS_LOGDET, WS_INFO,
    WS_LL, WS_IO0, WS_IO1, WS_IO2, WS_IO3, WS_MISC0, WS_MISC1, WS_MISC2, WS_MISC3, WS_MKK, WS_G0, WS_G1, WS_G2, WS_G3, WS_G4, WS_G5,
    WS_SCALE, WS_COUNTS, WS_NESTED
};
static_assert(WS_NESTED < GSUM_NWS, "workspace slots");

extern "C" int gsum_version(void) { return 100; }

extern "C" int gsum_ctx_create(int device, void *cuda_stream, gsum_ctx **out) {
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return -2;      // no CUDA device: there is no CPU path
    if (device < 0 || device >= count) return -3;
    if (cudaSetDevice(device) != cudaSuccess) return -4;
    gsum_ctx *c = new gsum_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return -5; }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {   // caller-held buffers (fit handles, resident factors) come from the device's stream-ordered pool and go back to it without a
        // device-wide synchronisation; the pool keeps what it has been given (a fit / Diagnostic per call would otherwise pay
        // 5-10 ms of cudaMalloc + cudaFree around 1 ms of kernels)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (cudaMallocHost((void **)&c->pin, (size_t)8 << 20) == cudaSuccess) c->pin_cap = (size_t)8 << 20; else { c->pin = nullptr; cudaGetLastError(); }
    const char *fc = getenv("GSUM_B200_FACTOR_CTAS");
    c->ht_factor_ctas = fc ? atoi(fc) : HT_FACTOR_CTAS;
    // schedule knobs are read ONCE, here; the per-call path reads no environment (GSUM_B200_DF_STATS, the instrumented build of the
    // factorisation kernel, is the one exception: a debugging switch tools/perf_chol.py flips between calls)
    c->ht_diag_delay = getenv("GSUM_B200_DIAG_DELAY") ? atoi(getenv("GSUM_B200_DIAG_DELAY")) : HT_DIAG_DELAY;
    c->ht_factor_workers = getenv("GSUM_B200_FACTOR_WORKERS") ? atoi(getenv("GSUM_B200_FACTOR_WORKERS")) : HT_FACTOR_WORKERS;
    const char *chn = getenv("GSUM_B200_CHAIN_MAX");
    c->ht_chain_max = chn ? atoi(chn) : HT_CHAIN_MAX;
    const char *sn = getenv("GSUM_B200_SMALLN");
    c->use_smalln = (sn && strcmp(sn, "0") == 0) ? 0 : 1;
    const char *thin = getenv("GSUM_B200_THIN");
    c->use_thin = (thin && strcmp(thin, "0") == 0) ? 0 : 1;
    *out = c;
    return 0;
}

extern "C" int gsum_ctx_destroy(gsum_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm) gsum_comm_destroy(c);
    for (int i = 0; i < GSUM_NWS; i++) if (c->ws[i]) cudaFree(c->ws[i]);
    if (c->ht_gtasks) cudaFree(c->ht_gtasks);
    if (c->ht_ftasks) cudaFree(c->ht_ftasks);
    if (c->df_flags) cudaFree(c->df_flags);
    if (c->df_ctl) cudaFree(c->df_ctl);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int gsum_ctx_synchronize(gsum_ctx *c) {
    if (!c) return -1;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int gsum_ctx_profile(gsum_ctx *c, int enable) {
    if (!c) return -1;
    c->prof_enabled = enable; c->prof_count = 0; c->prof_flops = 0.0;
    return 0;
}
extern "C" int gsum_ctx_profile_read(gsum_ctx *c, double *ms_total, double *flops_total, int64_t *n_brackets) {
    if (!c) return -1;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    double ms = 0.0;
    for (int i = 0; i < c->prof_count; i++) {
        float t = 0.f;
        GSUM_CUDA(c, cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        ms += t;
    }
    if (ms_total) *ms_total = ms;
    if (flops_total) *flops_total = c->prof_flops;
    if (n_brackets) *n_brackets = c->prof_count;
    c->prof_count = 0; c->prof_flops = 0.0;
    return 0;
}

// ---- caller-held device buffers (factors that stay in HBM between calls) -----------------------------------------
extern "C" int gsum_device_malloc(gsum_ctx *c, size_t bytes, void **out) {
    if (!c || !out || bytes == 0) return gsum_fail(c, -1, "gsum_device_malloc: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_CUDA(c, cudaMallocAsync(out, bytes, c->stream));
    return 0;
}
extern "C" int gsum_device_free(gsum_ctx *c, void *p) {
    if (!c) return -1;
    if (!p) return 0;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_CUDA(c, cudaFreeAsync(p, c->stream));               // stream-ordered: work enqueued on the buffer drains first
    return 0;
}
// direction: 0 host -> device, 1 device -> host, 2 device -> device; ordered on the context's stream, returns when done
extern "C" int gsum_device_copy(gsum_ctx *c, void *dst, const void *src, size_t bytes, int32_t direction) {
    if (!c || !dst || !src || direction < 0 || direction > 2) return gsum_fail(c, -1, "gsum_device_copy: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const cudaMemcpyKind kind = direction == 0 ? cudaMemcpyHostToDevice : direction == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    GSUM_CUDA(c, cudaMemcpyAsync(dst, src, bytes, kind, c->stream));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" const char *gsum_last_error(const gsum_ctx *c) { return c ? c->err : "null context"; }
extern "C" int64_t gsum_launch_count(const gsum_ctx *c) { return c ? c->launches : 0; }

// ---- buffer plumbing -------------------------------------------------------------------------------
// Input: device-resident view of a caller buffer (copy through workspace `slot` when it lives on the host).
static int dev_in(gsum_ctx *c, int slot, const void *p, size_t bytes, int mem_kind, const void **out) {
    if (!p) { *out = nullptr; return 0; }
    if (mem_kind == GSUM_MEM_DEVICE) { *out = p; return 0; }
    void *d;
    GSUM_TRY(gsum_ws(c, slot, bytes, &d));
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (c->pin && c->pin_off + need <= c->pin_cap) {
        memcpy(c->pin + c->pin_off, p, bytes);
        GSUM_CUDA(c, cudaMemcpyAsync(d, c->pin + c->pin_off, bytes, cudaMemcpyHostToDevice, c->stream));
        c->pin_off += need;
    } else GSUM_CUDA(c, cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream));
    *out = d;
    return 0;
}
// GSUM_MEM_FACTOR_DEVICE: the factor argument of a call is a device pointer while the other buffers are host pointers
static inline int factor_kind(int32_t &mem_kind) {
    const int fk = (mem_kind & GSUM_MEM_FACTOR_DEVICE) ? GSUM_MEM_DEVICE : (mem_kind & 1);
    mem_kind &= 1;
    return fk;
}
// Output: device buffer to write into (the caller's when it is a device pointer, workspace otherwise).
static int dev_out(gsum_ctx *c, int slot, void *p, size_t bytes, int mem_kind, void **out) {
    if (!p) { *out = nullptr; return 0; }
    if (mem_kind == GSUM_MEM_DEVICE) { *out = p; return 0; }
    return gsum_ws(c, slot, bytes, out);
}
static int dev_out_finish(gsum_ctx *c, void *host, const void *dev, size_t bytes, int mem_kind) {
    if (!host || mem_kind == GSUM_MEM_DEVICE) return 0;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (c->pin && c->npend < 16 && c->pin_off + need <= c->pin_cap) {
        GSUM_CUDA(c, cudaMemcpyAsync(c->pin + c->pin_off, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        c->pend[c->npend].dst = host; c->pend[c->npend].src = c->pin + c->pin_off; c->pend[c->npend].bytes = bytes;
        c->npend++;
        c->pin_off += need;
    } else GSUM_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
// hand the staged outputs to the caller (after a stream synchronisation)
static void flush_pending(gsum_ctx *c) {
    for (int i = 0; i < c->npend; i++) memcpy(c->pend[i].dst, c->pend[i].src, c->pend[i].bytes);
    c->npend = 0; c->pin_off = 0;
}
static int finish(gsum_ctx *c, int mem_kind) {
    GSUM_CUDA(c, cudaPeekAtLastError());
    GSUM_CUDA(c, cudaGetLastError());
    if (mem_kind == GSUM_MEM_HOST) {
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        flush_pending(c);
        if (c->df_ctl) {
            int sticky = 0;
            GSUM_CUDA(c, cudaMemcpy(&sticky, c->df_ctl + 2, sizeof(int), cudaMemcpyDeviceToHost));
            if (sticky) {
                cudaMemset(c->df_ctl + 2, 0, sizeof(int));
                return gsum_fail(c, -101, "dataflow Cholesky aborted by its watchdog (dependency wait exceeded %lld cycles)", (long long)DF_WATCHDOG_CYCLES);
            }
        }
    }
    return 0;
}

static int scale_coords(gsum_ctx *c, const double *dX, const double *dls, double *dXS, int64_t n, int d, int ls_dim,
                        int64_t batch) {
    int64_t total = batch * n * d;
    scale_coords_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(dX, dls, dXS, n, d, ls_dim, batch);
    LAUNCHED(c, 1);
    return 0;
}

// ---- heterogeneous schedule (hetero.cuh): GEMM CTAs + factor CTAs in one cooperative launch --------------------------
static int ht_upload(gsum_ctx *c, void **buf, size_t *cap, const std::vector<int4> &v) {
    const size_t bytes = v.size() * sizeof(int4);
    if (*cap < bytes || !*buf) {
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        if (*buf) GSUM_CUDA(c, cudaFree(*buf));
        GSUM_CUDA(c, cudaMalloc(buf, bytes + 4096));
        *cap = bytes + 4096;
    }
    if (bytes) GSUM_CUDA(c, cudaMemcpy(*buf, v.data(), bytes, cudaMemcpyHostToDevice));
    return 0;
}
// ---- tensor maps for the TMA-fed variant (driver entry point through the runtime: no -lcuda) -----------------------
typedef CUresult (*gsum_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int ht_make_map(gsum_ctx *c, CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    static gsum_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GSUM_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) return gsum_fail(c, -103, "cuTensorMapEncodeTiled not available");
        fn = (gsum_encode_tiled_fn)p;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld_elems * sizeof(double)};
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gsum_fail(c, -103, "cuTensorMapEncodeTiled failed (%d): rows %llu cols %llu ld %llu", (int)r,
                                            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems);
    return 0;
}

static int hetero_run(gsum_ctx *c, const BorderedBatch &P, int batch, bool solve_only) {
    GSUM_RANGE(solve_only ? "border solve (chol_hetero_tma_kernel)" : "factorisation (chol_hetero_tma_kernel)");
    const int nbt = P.Trows - P.T;
    const bool thin_last = c->use_thin && nbt > 0 && P.border_used > 0 && P.border_used - (nbt - 1) * GSUM_TILE <= 8 &&
                           P.border_used > (nbt - 1) * GSUM_TILE;
    const int delay = c->ht_diag_delay;
    // few matrices: chain mode (chain.cuh) — one chain worker CTA per matrix owns the diagonal band
    const bool chain = !solve_only && batch <= c->ht_chain_max;
    const int key[5] = {P.T, P.Trows, batch, (solve_only ? 1 : 0) | (thin_last ? 2 : 0) | (chain ? 4 : 0), delay};
    if (memcmp(key, c->ht_key, sizeof(key)) != 0 || !c->ht_gtasks) {
        std::vector<int4> gt, ft;
        if (chain) ht_build_chain_tasks(gt, P.T, P.Trows, batch, thin_last);
        else ht_build_tasks(gt, ft, P.T, P.Trows, batch, solve_only, thin_last, delay);
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));          // the previous lists may still be in use
        GSUM_TRY(ht_upload(c, &c->ht_gtasks, &c->ht_gcap, gt));
        GSUM_TRY(ht_upload(c, &c->ht_ftasks, &c->ht_fcap, ft));
        memcpy(c->ht_key, key, sizeof(key));
        c->ht_ng = (int)gt.size(); c->ht_nf = (int)ft.size();
    }
    const size_t fbytes = sizeof(int) * ((size_t)batch * P.Trows * P.T + (size_t)batch * P.T);       // tile flags + pre flags
    if (c->df_flags_cap < fbytes) {
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->df_flags) GSUM_CUDA(c, cudaFree(c->df_flags));
        GSUM_CUDA(c, cudaMalloc(&c->df_flags, fbytes + 4096));
        c->df_flags_cap = fbytes + 4096;
    }
    if (!c->df_ctl) {
        GSUM_CUDA(c, cudaMalloc((void **)&c->df_ctl, 8 * sizeof(int)));
        GSUM_CUDA(c, cudaMemsetAsync(c->df_ctl, 0, 8 * sizeof(int), c->stream));
    }
    void *dM;
    GSUM_TRY(gsum_ws(c, WS_MKK, sizeof(double) * (size_t)batch * P.T * GSUM_TILE * GSUM_TILE, &dM));
    const int64_t nflags = (int64_t)batch * P.Trows * P.T + (int64_t)batch * P.T;
    ht_init_kernel<<<(unsigned)((nflags + 255) / 256 + 1), 256, 0, c->stream>>>((int *)c->df_flags, c->df_ctl, batch, P.Trows, P.T, solve_only ? 1 : 0);
    c->launches += 1;
    if (solve_only) {
        ht_mkk_from_factor_kernel<<<dim3(P.T, batch), CHOL_THREADS, 0, c->stream>>>(P, (double *)dM);
        c->launches += 1;
    }
    HeteroArgs D;
    D.P = P; D.gtasks = (const int4 *)c->ht_gtasks; D.ngtasks = c->ht_ng; D.ftasks = (const int4 *)c->ht_ftasks; D.nftasks = c->ht_nf;
    D.nf0 = (!solve_only && !chain && c->ht_nf >= batch) ? batch : 0;          // df_build_tasks emits the column-0 diagonal tiles first
    D.ctl = c->df_ctl; D.flags = (int *)c->df_flags; D.M = (double *)dM; D.stats = nullptr;
    D.chain = chain ? 1 : 0; D.pre = (int *)c->df_flags + (int64_t)batch * P.Trows * P.T;
    // factor CTAs: HT_FACTOR_WORKERS (four) workers each; never more than the diagonal tiles can use, never all of the SMs
    int nf = 0;
    int nwk = c->ht_factor_workers;
    if (nwk < 1) nwk = 1;
    if (nwk > 4) nwk = 4;
    D.nworkers = nwk;
    if (c->ht_nf > 0) {
        nf = c->ht_factor_ctas;
        const int cap = (c->ht_nf + nwk - 1) / nwk;
        if (nf > cap) nf = cap;
        if (nf > c->sm_count / 2) nf = c->sm_count / 2;
        if (nf < 1) nf = 1;
    }
    if (chain) { nf = batch; D.nworkers = 1; }
    int ng = c->sm_count - nf;
    if (ng > c->ht_ng) ng = c->ht_ng;
    D.nfactor_ctas = nf;
    const int grid = nf + ng;
    if (grid <= 0) return 0;
    static long long *dbg_stats = nullptr;
    if (getenv("GSUM_B200_DF_STATS")) {
        if (!dbg_stats) cudaMalloc((void **)&dbg_stats, sizeof(long long) * HT_NSTAT * 1024);
        cudaMemsetAsync(dbg_stats, 0, sizeof(long long) * HT_NSTAT * 1024, c->stream);
        D.stats = dbg_stats;
    }
    static long long *dbg_trace = nullptr;
    D.trace = nullptr; D.trace_cta = -1;
    if (D.stats && getenv("GSUM_B200_TRACE_CTA")) {
        if (!dbg_trace) cudaMalloc((void **)&dbg_trace, sizeof(long long) * 3 * 64 * 8);
        cudaMemsetAsync(dbg_trace, 0, sizeof(long long) * 3 * 64 * 8, c->stream);
        D.trace = dbg_trace; D.trace_cta = atoi(getenv("GSUM_B200_TRACE_CTA"));
    }
    {
    if (!c->hx_ready) {
        GSUM_CUDA(c, cudaFuncSetAttribute(chol_hetero_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HX_SMEM_BYTES));
        GSUM_CUDA(c, cudaFuncSetAttribute(chol_hetero_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HX_SMEM_BYTES));
        c->hx_ready = 1;
    }
    HeteroMaps maps;
    const uint64_t rowsA = (uint64_t)(P.bstride / P.ld), rowsW = (uint64_t)(P.wstride / P.ld);
    GSUM_TRY(ht_make_map(c, &maps.A, P.A, (uint64_t)(batch - 1) * rowsA + (uint64_t)P.T * GSUM_TILE, (uint64_t)P.ld, (uint64_t)P.ld, GSUM_TILE));
    if (P.Trows > P.T) {
        const uint64_t wr = (uint64_t)(batch - 1) * rowsW + (uint64_t)(P.Trows - P.T) * GSUM_TILE;
        GSUM_TRY(ht_make_map(c, &maps.W, P.W, wr, (uint64_t)P.ld, (uint64_t)P.ld, GSUM_TILE));
        GSUM_TRY(ht_make_map(c, &maps.W8, P.W, wr, (uint64_t)P.ld, (uint64_t)P.ld, 8));
    } else { maps.W = maps.A; maps.W8 = maps.A; }
    GSUM_TRY(ht_make_map(c, &maps.M, dM, (uint64_t)batch * P.T * GSUM_TILE, GSUM_TILE, GSUM_TILE, GSUM_TILE));
    void *args[] = {&D, &maps};
    const void *kfn = D.stats ? (const void *)chol_hetero_tma_kernel<true> : (const void *)chol_hetero_tma_kernel<false>;
    GSUM_CUDA(c, cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(HX_THREADS), args, HX_SMEM_BYTES, c->stream));
    }
    if (!solve_only && P.logdet_part) {
        ht_logdet_kernel<<<dim3(P.T, batch), 32, 0, c->stream>>>(P);
        c->launches += 1;
    }
    df_check_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, c->stream>>>(c->df_ctl + 1, P.info, batch, c->df_ctl + 2);
    c->launches += 2;
    GSUM_CUDA(c, cudaPeekAtLastError());
    if (D.stats) {
        std::vector<long long> h(HT_NSTAT * grid);
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(h.data(), dbg_stats, sizeof(long long) * HT_NSTAT * grid, cudaMemcpyDeviceToHost);
        if (D.trace) {
            std::vector<long long> tr(3 * 64 * 8);
            cudaMemcpy(tr.data(), dbg_trace, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost);
            for (int q = 0; q < 3; q++) for (int n = 0; n < 64; n++) {
                const long long *e = tr.data() + (q * 64 + n) * 8;
                if (e[0]) fprintf(stderr, "[trace] %d %d %lld %lld %lld %lld %lld %lld %lld %lld\n", q, n, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]);
            }
        }
        double f[6] = {0, 0, 0, 0, 0, 0}, a[HT_NSTAT] = {0};
        const int ngrp = HX_NG;
        for (int g = 0; g < nf; g++) for (int wk = 0; wk < nwk; wk++) for (int q = 0; q < 6; q++) f[q] += (double)h[HT_NSTAT * g + wk * 6 + q];
        for (int g = nf; g < grid; g++) for (int grp = 0; grp < ngrp; grp++) for (int q = 0; q < 12; q++) a[q] += (double)h[HT_NSTAT * g + grp * 12 + q];
        if (chain) {
            double ch[10] = {0};
            for (int g = 0; g < nf; g++) for (int q = 0; q < 10; q++) ch[q] += (double)h[HT_NSTAT * g + q];
            const double cols = ch[6] > 0 ? ch[6] : 1;
            fprintf(stderr, "[ht] chain workers %d: cycles/worker %.0f | per column: wait_pre %.0f potrf %.0f invert %.0f solve %.0f wait_helper %.0f update %.0f (sum %.0f) | invert compute %.0f, own update %.0f\n",
                    nf, ch[0] / nf, ch[1] / cols, ch[2] / cols, ch[3] / cols, ch[4] / cols, ch[7] / cols, ch[5] / cols, (ch[1] + ch[2] + ch[3] + ch[4] + ch[5] + ch[7]) / cols, ch[8] / cols, ch[9] / cols);
        } else
        if (nf) fprintf(stderr, "[ht] factor CTAs %d: cycles/worker %.0f | wait_S %.1f%% | busy %.1f%% (%.0f cycles per diagonal tile: load %.0f, potrf %.0f; %.1f tiles per worker)\n",
                        nf, f[0] / (nwk * nf), 100 * f[1] / f[0], 100 * f[2] / f[0], f[2] / (f[3] + 1e-9), f[4] / (f[3] + 1e-9), f[5] / (f[3] + 1e-9), f[3] / (nwk * nf));
        if (ng) fprintf(stderr, "[ht] GEMM CTAs %d x %d groups: cycles/group %.0f, tasks/group %.1f | producer: wait_queue %.1f%% wait_flag %.1f%% wait_ring %.1f%% wait_Mkk %.1f%% | math: wait_queue %.1f%% wait_operands %.1f%% wait_Mkk %.1f%% trsm %.1f%% fence+flag %.1f%%\n",
                        ng, ngrp, a[5] / (ngrp * ng), a[11] / (ngrp * ng), 100 * a[1] / a[0], 100 * a[2] / a[0], 100 * a[3] / a[0], 100 * a[4] / a[0],
                        100 * a[6] / a[5], 100 * a[7] / a[5], 100 * a[8] / a[5], 100 * a[9] / a[5], 100 * a[10] / a[5]);
    }
    return 0;
}
static int factor_run(gsum_ctx *c, const BorderedBatch &P, int batch) { return hetero_run(c, P, batch, false); }
static int solve_run(gsum_ctx *c, const BorderedBatch &P, int batch) {
    if (P.Trows - P.T <= 0) return 0;
    return hetero_run(c, P, batch, true);
}

// ---- K1 ---------------------------------------------------------------------------------------------
extern "C" int gsum_kernel_matrix(gsum_ctx *c, const double *X1, int64_t n1, const double *X2, int64_t n2, int32_t d,
                                  const double *ls, int32_t ls_dim, double constant, double noise, double *out,
                                  int32_t mem_kind) {
    GSUM_RANGE("gsum_kernel_matrix");
    if (!c || !X1 || !ls || !out || n1 <= 0 || d <= 0 || d > COV_MAXD || (ls_dim != 1 && ls_dim != d))
        return gsum_fail(c, -1, "gsum_kernel_matrix: bad argument (d must be 1..%d, ls_dim 1 or d)", COV_MAXD);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const bool sym = (X2 == nullptr);
    if (sym) n2 = n1;
    const void *dX1, *dX2 = nullptr, *dls;
    GSUM_TRY(dev_in(c, WS_X, X1, sizeof(double) * n1 * d, mem_kind, &dX1));
    if (!sym) GSUM_TRY(dev_in(c, WS_IO0, X2, sizeof(double) * n2 * d, mem_kind, &dX2));
    GSUM_TRY(dev_in(c, WS_LS, ls, sizeof(double) * ls_dim, mem_kind, &dls));
    void *xs;
    GSUM_TRY(gsum_ws(c, WS_XS, sizeof(double) * (n1 + n2) * d, &xs));
    double *xs1 = (double *)xs, *xs2 = xs1 + n1 * d;
    scale_coords(c, (const double *)dX1, (const double *)dls, xs1, n1, d, ls_dim, 1);
    if (!sym) scale_coords(c, (const double *)dX2, (const double *)dls, xs2, n2, d, ls_dim, 1);
    void *dout;
    GSUM_TRY(dev_out(c, WS_MAT, out, sizeof(double) * n1 * n2, mem_kind, &dout));
    CrossArgs P;
    P.XS1 = xs1; P.n1 = n1; P.XS2 = sym ? xs1 : xs2; P.n2 = n2; P.d = d;
    P.constant = constant; P.diag_add = noise; P.sym_diag = sym ? 1 : 0;
    P.out = (double *)dout; P.ldo = n2; P.rows_out = n1; P.cols_out = n2;
    dim3 grid((unsigned)((n2 + 63) / 64), (unsigned)((n1 + 63) / 64));
    cov_cross_kernel<<<grid, 256, 0, c->stream>>>(P);
    LAUNCHED(c, 1);
    GSUM_TRY(dev_out_finish(c, out, dout, sizeof(double) * n1 * n2, mem_kind));
    return finish(c, mem_kind);
}

// ---- K2 ---------------------------------------------------------------------------------------------
// contiguous (batch, n, n)  <->  bordered padded layout
__global__ void pad_in_kernel(const double *__restrict__ src, int64_t n, double *__restrict__ dst, int64_t ld, int64_t bstride,
                              int64_t rows) {
    const int64_t b = blockIdx.z, r = blockIdx.y;
    const double *s = src + b * n * n + r * n;
    double *dr = dst + b * bstride + r * ld;
    for (int64_t cidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cidx < ld; cidx += (int64_t)gridDim.x * blockDim.x)
        dr[cidx] = (r < n && cidx < n) ? s[cidx] : (r == cidx ? 1.0 : 0.0);
}
__global__ void pad_out_lower_kernel(const double *__restrict__ src, int64_t ld, int64_t bstride, double *__restrict__ dst, int64_t n) {
    const int64_t b = blockIdx.z, r = blockIdx.y;
    const double *s = src + b * bstride + r * ld;
    double *dr = dst + b * n * n + r * n;
    for (int64_t cidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cidx < n; cidx += (int64_t)gridDim.x * blockDim.x)
        dr[cidx] = cidx <= r ? s[cidx] : 0.0;
}
__global__ void fill_kernel(double *p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void add_diag_kernel(double *F, int64_t ld, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) F[i * ld + i] = __dadd_rn(F[i * ld + i], v);
}
// contiguous lower factor (n,n) -> padded (np x np): strict upper part forced to zero, identity padding
__global__ void pad_lower_kernel(const double *__restrict__ L, int64_t n, double *__restrict__ F, int64_t np) {
    const int64_t r = blockIdx.y;
    for (int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cc < np; cc += (int64_t)gridDim.x * blockDim.x)
        F[r * np + cc] = (r < n && cc < n) ? (cc <= r ? L[r * n + cc] : 0.0) : (r == cc ? 1.0 : 0.0);
}
__global__ void logdet_sum_kernel(const double *__restrict__ part, int T, int64_t batch, double *__restrict__ out) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double s = 0.0;
    for (int k = 0; k < T; k++) s += part[b * T + k];
    out[b] = s;
}

// Factor `batch` padded matrices already resident in workspace WS_MAT (bordered layout, Trows tile rows).
static int factor_bordered(gsum_ctx *c, double *dA, int64_t n, int Trows, int64_t batch, int **dinfo_out, double **dpart_out) {
    // c->prof_border_rows: right-hand sides riding along (set by the caller; 0 for a plain factorisation)
    const int T = (int)(gsum_pad64(n) / GSUM_TILE);
    void *dinfo, *dpart;
    GSUM_TRY(gsum_ws(c, WS_INFO, sizeof(int) * batch, &dinfo));
    GSUM_TRY(gsum_ws(c, WS_LOGDET, sizeof(double) * batch * T, &dpart));
    GSUM_CUDA(c, cudaMemsetAsync(dinfo, 0, sizeof(int) * batch, c->stream));
    BorderedBatch P;
    P.A = dA; P.ld = (int64_t)T * GSUM_TILE; P.bstride = (int64_t)Trows * GSUM_TILE * P.ld;
    P.W = dA + (int64_t)T * GSUM_TILE * P.ld; P.wstride = P.bstride;
    P.T = T; P.Trows = Trows; P.info = (int *)dinfo; P.logdet_part = (double *)dpart; P.n = (int)n;
    P.border_used = (int)c->prof_border_rows;
    const bool prof = c->prof_enabled && c->prof_count < 256;
    if (prof) {
        if (!c->prof_ev[2 * c->prof_count]) { cudaEventCreate(&c->prof_ev[2 * c->prof_count]); cudaEventCreate(&c->prof_ev[2 * c->prof_count + 1]); }
        cudaEventRecord(c->prof_ev[2 * c->prof_count], c->stream);
    }
    GSUM_TRY(factor_run(c, P, (int)batch));
    if (prof) {
        cudaEventRecord(c->prof_ev[2 * c->prof_count + 1], c->stream);
        c->prof_count++;
        // algorithmic work: N^3/3 per factorisation + N^2 per forward-solved border row, true (unpadded) sizes
        c->prof_flops += (double)batch * ((double)n * n * n / 3.0 + (double)n * n * (double)c->prof_border_rows);
    }
    *dinfo_out = (int *)dinfo; *dpart_out = (double *)dpart;
    return 0;
}

extern "C" int gsum_cholesky(gsum_ctx *c, double *A, int64_t n, int64_t batch, int32_t *info, double *logdet,
                             int32_t mem_kind) {
    GSUM_RANGE("gsum_cholesky");
    if (!c || !A || n <= 0 || batch <= 0) return gsum_fail(c, -1, "gsum_cholesky: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int64_t np = gsum_pad64(n);
    const int T = (int)(np / GSUM_TILE);
    const void *dsrc;
    GSUM_TRY(dev_in(c, WS_IO0, A, sizeof(double) * batch * n * n, mem_kind, &dsrc));
    void *dmat;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * batch * np * np, &dmat));
    dim3 g1((unsigned)((np + 255) / 256), (unsigned)np, (unsigned)batch);
    pad_in_kernel<<<g1, 256, 0, c->stream>>>((const double *)dsrc, n, (double *)dmat, np, np * np, np);
    LAUNCHED(c, 1);
    int *dinfo; double *dpart;
    GSUM_TRY(factor_bordered(c, (double *)dmat, n, T, batch, &dinfo, &dpart));
    dim3 g2((unsigned)((n + 255) / 256), (unsigned)n, (unsigned)batch);
    double *ddst = (double *)dsrc;       // in place: the caller's device buffer, or our staging copy
    pad_out_lower_kernel<<<g2, 256, 0, c->stream>>>((const double *)dmat, np, np * np, ddst, n);
    LAUNCHED(c, 1);
    GSUM_TRY(dev_out_finish(c, A, ddst, sizeof(double) * batch * n * n, mem_kind));
    if (info) {
        if (mem_kind == GSUM_MEM_DEVICE) GSUM_CUDA(c, cudaMemcpyAsync(info, dinfo, sizeof(int) * batch, cudaMemcpyDeviceToDevice, c->stream));
        else GSUM_CUDA(c, cudaMemcpyAsync(info, dinfo, sizeof(int) * batch, cudaMemcpyDeviceToHost, c->stream));
    }
    if (logdet) {
        void *dl;
        GSUM_TRY(dev_out(c, WS_LL, logdet, sizeof(double) * batch, mem_kind, &dl));
        logdet_sum_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, c->stream>>>(dpart, T, batch, (double *)dl);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, logdet, dl, sizeof(double) * batch, mem_kind));
    }
    return finish(c, mem_kind);
}

// ---- K1-K4 fused: the (Q, l) grid ---------------------------------------------------------------------
template <int R>
static void launch_gram(gsum_ctx *c, const double *A, int64_t ld, int64_t bstride, int T, int64_t n, double *G, int64_t nq_rows,
                        int64_t batch) {
    dim3 grid((unsigned)nq_rows, (unsigned)batch);
    gram_rows_kernel<R><<<grid, 256, 0, c->stream>>>(A, ld, bstride, T, n, G, nq_rows);
}

extern "C" int gsum_lml_grid(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *dy, int32_t n_c,
                             const double *ref, const int32_t *orders, const double *ls, int64_t n_ls, int32_t ls_dim,
                             const double *Q, int64_t n_q, int32_t q_x_dependent, const double *detf, double constant,
                             double noise, double nugget, double center0, double disp0, double df0, double scale0,
                             int32_t student, double *ll, double *logdet, int32_t *status, int32_t mem_kind) {
    GSUM_RANGE("gsum_lml_grid");
    if (!c || !X || !dy || !ref || !orders || !ls || !Q || !ll)
        return gsum_fail(c, -1, "gsum_lml_grid: null argument");
    if (n <= 0 || n_ls <= 0 || n_q <= 0 || d <= 0 || d > COV_MAXD || (ls_dim != 1 && ls_dim != d) || n_c < 1 || n_c + 1 > LML_MAXR)
        return gsum_fail(c, -1, "gsum_lml_grid: bad shape (d<=%d, n_c<=%d, ls_dim in {1,d})", COV_MAXD, LML_MAXR - 1);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int R = n_c + 1;
    const int64_t np = gsum_pad64(n);
    const int T = (int)(np / GSUM_TILE);
    const int64_t nq_rows = q_x_dependent ? n_q : 1;            // RHS blocks per matrix
    const int64_t r_rhs = 1 + nq_rows * n_c;                    // border rows in use
    const int64_t rp = gsum_pad64(r_rhs);
    const int Trows = T + (int)(rp / GSUM_TILE);

    const void *dX, *ddy, *dref, *dord, *dls, *dQ, *ddetf;
    GSUM_TRY(dev_in(c, WS_X, X, sizeof(double) * n * d, mem_kind, &dX));
    GSUM_TRY(dev_in(c, WS_DY, dy, sizeof(double) * n * n_c, mem_kind, &ddy));
    GSUM_TRY(dev_in(c, WS_REF, ref, sizeof(double) * n, mem_kind, &dref));
    GSUM_TRY(dev_in(c, WS_ORD, orders, sizeof(int32_t) * n_c, mem_kind, &dord));
    GSUM_TRY(dev_in(c, WS_LS, ls, sizeof(double) * n_ls * ls_dim, mem_kind, &dls));
    GSUM_TRY(dev_in(c, WS_Q, Q, sizeof(double) * n_q * (q_x_dependent ? n : 1), mem_kind, &dQ));
    GSUM_TRY(dev_in(c, WS_DETF, detf, sizeof(double) * n_q, mem_kind, &ddetf));

    const bool small = c->use_smalln && !q_x_dependent && n <= SN_MAXN && R <= SN_MAXR && d <= SN_MAXD;
    // RHS rows (shared by every length scale): basis + coefficients, transposed
    void *drhs = nullptr;
    if (!small) {
        GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * r_rhs * n, &drhs));
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)r_rhs);
        stage_rhs_kernel<<<grid, 256, 0, c->stream>>>((double *)drhs, n, (const double *)ddy, (const double *)dref,
                                                      q_x_dependent ? (const double *)dQ : nullptr, (const int32_t *)dord, n,
                                                      n_c, nq_rows);
        LAUNCHED(c, 1);
    }
    void *dll;
    GSUM_TRY(dev_out(c, WS_LL, ll, sizeof(double) * n_q * n_ls, mem_kind, &dll));
    void *dlogdet = nullptr;
    if (logdet) GSUM_TRY(dev_out(c, WS_MISC0, logdet, sizeof(double) * n_ls, mem_kind, &dlogdet));

    // ---- small N, scalar Q: one CTA per length scale, everything in shared memory (smalln.cuh) ---------------------------
    if (small) {
        void *dgram, *dld, *dinfo;
        GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * n_ls * R * R, &dgram));
        GSUM_TRY(gsum_ws(c, WS_LOGDET, sizeof(double) * n_ls, &dld));
        GSUM_TRY(gsum_ws(c, WS_MISC1, sizeof(int) * n_ls, &dinfo));
        SmallNArgs SA;
        SA.X = (const double *)dX; SA.ls = (const double *)dls; SA.dy = (const double *)ddy; SA.ref = (const double *)dref;
        SA.n = n; SA.d = d; SA.ls_dim = ls_dim; SA.n_c = n_c; SA.constant = constant; SA.noise = noise; SA.nugget = nugget;
        SA.G = (double *)dgram; SA.logdet = (double *)dld; SA.info = (int *)dinfo;
        const size_t smem = smalln_smem_bytes(n);
        if (!c->sn_ready) {
            GSUM_CUDA(c, cudaFuncSetAttribute(smalln_lml_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smalln_smem_bytes(SN_MAXN)));
            c->sn_ready = 1;
        }
        smalln_lml_kernel<<<(unsigned)n_ls, SN_THREADS, smem, c->stream>>>(SA);
        LmlCellArgs LA;
        LA.G = (const double *)dgram; LA.logdet_part = (const double *)dld; LA.info = (const int *)dinfo; LA.T = 1; LA.R = R;
        LA.n = n; LA.n_l = n_ls; LA.n_q = n_q; LA.separable = 1;
        LA.Q = (const double *)dQ; LA.orders = (const int32_t *)dord; LA.detf = (const double *)ddetf;
        LA.center0 = center0; LA.disp0 = disp0; LA.df0 = df0; LA.scale0 = scale0; LA.student = student;
        LA.ll = (double *)dll; LA.logdet_out = (double *)dlogdet; LA.post = nullptr;
        lml_cell_chunk_kernel<<<(unsigned)((n_ls * n_q + 127) / 128), 128, 0, c->stream>>>(LA, 0, n_ls);
        LAUNCHED(c, 2);
        GSUM_TRY(dev_out_finish(c, ll, dll, sizeof(double) * n_q * n_ls, mem_kind));
        GSUM_TRY(dev_out_finish(c, logdet, dlogdet, sizeof(double) * n_ls, mem_kind));
        if (status) {
            GSUM_CUDA(c, cudaMemcpyAsync(status, dinfo, sizeof(int) * n_ls,
                                         mem_kind == GSUM_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        }
        return finish(c, mem_kind);
    }

    // length scales are processed in chunks sized to a fixed HBM budget (whole grid at once for the BASELINE configs)
    const int64_t per_mat = (int64_t)Trows * GSUM_TILE * np;                     // doubles
    const int64_t budget = (int64_t)24 << 30;                                     // bytes
    int64_t chunk = budget / (per_mat * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > n_ls) chunk = n_ls;
    void *dmat, *dxs, *dgram, *dinfo_all;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * chunk * per_mat, &dmat));
    GSUM_TRY(gsum_ws(c, WS_XS, sizeof(double) * chunk * n * d, &dxs));
    GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * chunk * nq_rows * R * R, &dgram));
    GSUM_TRY(gsum_ws(c, WS_MISC1, sizeof(int) * n_ls, &dinfo_all));

    for (int64_t l0 = 0; l0 < n_ls; l0 += chunk) {
        const int64_t nb = (n_ls - l0 < chunk) ? (n_ls - l0) : chunk;
        scale_coords(c, (const double *)dX, (const double *)dls + l0 * ls_dim, (double *)dxs, n, d, ls_dim, nb);
        CovArgs CA;
        CA.XS = (const double *)dxs; CA.n = n; CA.d = d; CA.constant = constant; CA.noise = noise; CA.nugget = nugget;
        CA.A = (double *)dmat; CA.ld = np; CA.bstride = per_mat; CA.T = T;
        CA.tiles_per_cta = cov_tiles_per_cta(T * (T + 1) / 2, nb, c->sm_count);
        dim3 gcov((unsigned)((T * (T + 1) / 2 + CA.tiles_per_cta - 1) / CA.tiles_per_cta), (unsigned)nb);
        cov_sym_kernel<<<gcov, 256, 0, c->stream>>>(CA);
        // a last border tile row that runs as thin (8-row) tasks is only ever touched in its first 8 rows
        int64_t fill_rows = rp;
        if (c->use_thin && r_rhs - (rp - GSUM_TILE) <= 8) fill_rows = rp - GSUM_TILE + 8;
        dim3 gb((unsigned)fill_rows, (unsigned)((np + 255) / 256), (unsigned)nb);
        border_fill_kernel<<<gb, 256, 0, c->stream>>>((double *)dmat, np, per_mat, T, (int)fill_rows, (const double *)drhs, (int)r_rhs, n, n);
        LAUNCHED(c, 2);
        int *dinfo; double *dpart;
        c->prof_border_rows = r_rhs;
        GSUM_TRY(factor_bordered(c, (double *)dmat, n, Trows, nb, &dinfo, &dpart));
        c->prof_border_rows = 0;
        switch (R) {
#define GRAM_CASE(RR) case RR: launch_gram<RR>(c, (const double *)dmat, np, per_mat, T, n, (double *)dgram, nq_rows, nb); break;
            GRAM_CASE(2) GRAM_CASE(3) GRAM_CASE(4) GRAM_CASE(5) GRAM_CASE(6) GRAM_CASE(7) GRAM_CASE(8)
#undef GRAM_CASE
            default: {
                dim3 grid((unsigned)nq_rows, (unsigned)nb);
                gram_rows_generic_kernel<<<grid, 256, 0, c->stream>>>((const double *)dmat, np, per_mat, T, n, (double *)dgram, nq_rows, R);
            }
        }
        LAUNCHED(c, 1);
        LmlCellArgs LA;
        LA.G = (const double *)dgram; LA.logdet_part = dpart; LA.info = dinfo; LA.T = T; LA.R = R;
        LA.n = n; LA.n_l = nb; LA.n_q = n_q; LA.separable = q_x_dependent ? 0 : 1;
        LA.Q = (const double *)dQ; LA.orders = (const int32_t *)dord; LA.detf = (const double *)ddetf;
        LA.center0 = center0; LA.disp0 = disp0; LA.df0 = df0; LA.scale0 = scale0; LA.student = student;
        LA.ll = nullptr; LA.logdet_out = dlogdet ? (double *)dlogdet + l0 : nullptr; LA.post = nullptr;
        // cells of this chunk go to columns [l0, l0+nb) of the (n_q, n_ls) grid
        LA.ll = (double *)dll;
        lml_cell_chunk_kernel<<<(unsigned)((nb * n_q + 127) / 128), 128, 0, c->stream>>>(LA, l0, n_ls);
        LAUNCHED(c, 1);
        GSUM_CUDA(c, cudaMemcpyAsync((int *)dinfo_all + l0, dinfo, sizeof(int) * nb, cudaMemcpyDeviceToDevice, c->stream));
    }
    GSUM_TRY(dev_out_finish(c, ll, dll, sizeof(double) * n_q * n_ls, mem_kind));
    GSUM_TRY(dev_out_finish(c, logdet, dlogdet, sizeof(double) * n_ls, mem_kind));
    if (status) {
        GSUM_CUDA(c, cudaMemcpyAsync(status, dinfo_all, sizeof(int) * n_ls,
                                     mem_kind == GSUM_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    }
    return finish(c, mem_kind);
}


extern "C" int gsum_grid_normalize(gsum_ctx *c, const double *ll, int64_t count, double *post, double *lse, int32_t mem_kind) {
    GSUM_RANGE("gsum_grid_normalize");
    if (!c || !ll || count <= 0) return gsum_fail(c, -1, "gsum_grid_normalize: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const void *dll; void *dpost, *dlse;
    GSUM_TRY(dev_in(c, WS_IO0, ll, sizeof(double) * count, mem_kind, &dll));
    GSUM_TRY(dev_out(c, WS_IO1, post, sizeof(double) * count, mem_kind, &dpost));
    GSUM_TRY(dev_out(c, WS_IO2, lse, sizeof(double), mem_kind, &dlse));
    if (count >= 8192)      // one 8-CTA cluster (DSMEM exchange); tiny grids stay on the single-CTA kernel
        grid_normalize_cluster_kernel<<<GN_CLUSTER, 1024, 0, c->stream>>>((const double *)dll, count, (double *)dpost, (double *)dlse);
    else
        grid_normalize_kernel<<<1, 1024, 0, c->stream>>>((const double *)dll, count, (double *)dpost, (double *)dlse);
    LAUNCHED(c, 1);
    GSUM_TRY(dev_out_finish(c, post, dpost, sizeof(double) * count, mem_kind));
    GSUM_TRY(dev_out_finish(c, lse, dlse, sizeof(double), mem_kind));
    return finish(c, mem_kind);
}

// ---- the collective of the sharded grid through the C ABI (SURVEY.md 8b lower face; north_star: ONE all-gather + device logsumexp) ----
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already carries, e.g. torch's, else the system one), so the
// library has no link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>                   // types and enums only; no symbol of it is linked
struct GsumNccl {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static GsumNccl *gsum_nccl(gsum_ctx *c) {
    static GsumNccl N = {};
    if (N.lib) return &N;
    const char *names[] = {getenv("GSUM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) if (nm && !h) h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { gsum_fail(c, -110, "gsum_comm: libnccl.so.2 not found (%s); set GSUM_B200_NCCL_LIB", dlerror()); return nullptr; }
    N.GetUniqueId = (decltype(N.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))dlsym(h, "ncclCommInitRank");
    N.CommDestroy = (decltype(N.CommDestroy))dlsym(h, "ncclCommDestroy");
    N.AllGather = (decltype(N.AllGather))dlsym(h, "ncclAllGather");
    N.AllReduce = (decltype(N.AllReduce))dlsym(h, "ncclAllReduce");
    N.GetErrorString = (decltype(N.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!N.GetUniqueId || !N.CommInitRank || !N.CommDestroy || !N.AllGather || !N.AllReduce || !N.GetErrorString) {
        gsum_fail(c, -110, "gsum_comm: the NCCL library lacks an expected entry point");
        return nullptr;
    }
    N.lib = h;
    return &N;
}
#define GSUM_NCCL(ctx, N, call)                                                                                      \
    do {                                                                                                             \
        ncclResult_t _r = (call);                                                                                    \
        if (_r != ncclSuccess) return gsum_fail((ctx), -111, "NCCL error at %s:%d (%s)", __FILE__, __LINE__, (N)->GetErrorString(_r)); \
    } while (0)

extern "C" int gsum_comm_unique_id(gsum_ctx *c, void *id_out) {
    if (!c || !id_out) return gsum_fail(c, -1, "gsum_comm_unique_id: bad argument");
    GsumNccl *N = gsum_nccl(c);
    if (!N) return -110;
    ncclUniqueId id;
    GSUM_NCCL(c, N, N->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));                    // GSUM_COMM_ID_BYTES = 128
    return 0;
}
extern "C" int gsum_comm_init(gsum_ctx *c, int32_t nranks, int32_t rank, const void *nccl_unique_id) {
    GSUM_RANGE("gsum_comm_init");
    if (!c || nranks < 1 || rank < 0 || rank >= nranks || !nccl_unique_id) return gsum_fail(c, -1, "gsum_comm_init: bad argument");
    if (c->comm) return gsum_fail(c, -1, "gsum_comm_init: this context already has a communicator");
    GsumNccl *N = gsum_nccl(c);
    if (!N) return -110;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    ncclComm_t comm;
    GSUM_NCCL(c, N, N->CommInitRank(&comm, nranks, id, rank));
    c->comm = comm; c->comm_nranks = nranks; c->comm_rank = rank;
    return 0;
}
extern "C" int gsum_comm_destroy(gsum_ctx *c) {
    if (!c) return -1;
    if (!c->comm) return 0;
    GsumNccl *N = gsum_nccl(c);
    if (!N) return -110;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    N->CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr; c->comm_nranks = 0; c->comm_rank = 0;
    return 0;
}
// full[q][r + j * P] = recv[r][q][j]: undo the round-robin deal of the length scales (rank r owns l = r, r + P, ...)
__global__ void grid_unshard_kernel(const double *__restrict__ recv, int64_t n_q, int64_t per, int nranks, int64_t n_ls, double *__restrict__ full) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_q * n_ls) return;
    const int64_t q = idx / n_ls, l = idx % n_ls;
    const int64_t r = l % nranks, j = l / nranks;
    full[idx] = recv[(r * n_q + q) * per + j];
}
extern "C" int gsum_grid_allgather(gsum_ctx *c, const double *block, int64_t n_q, int64_t per, int64_t n_ls, double *ll_full, double *post,
                                   double *lse, int32_t mem_kind) {
    GSUM_RANGE("gsum_grid_allgather");
    if (!c || !block || !ll_full || n_q <= 0 || per <= 0 || n_ls <= 0) return gsum_fail(c, -1, "gsum_grid_allgather: bad argument");
    if (!c->comm) return gsum_fail(c, -1, "gsum_grid_allgather: no communicator (gsum_comm_init)");
    const int P = c->comm_nranks;
    if (per != (n_ls + P - 1) / P) return gsum_fail(c, -1, "gsum_grid_allgather: per must be ceil(n_ls / nranks)");
    GsumNccl *N = gsum_nccl(c);
    if (!N) return -110;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const void *dblock; void *drecv, *dfull, *dpost = nullptr, *dlse = nullptr;
    GSUM_TRY(dev_in(c, WS_IO0, block, sizeof(double) * n_q * per, mem_kind, &dblock));
    GSUM_TRY(gsum_ws(c, WS_IO1, sizeof(double) * P * n_q * per, &drecv));
    GSUM_TRY(dev_out(c, WS_IO2, ll_full, sizeof(double) * n_q * n_ls, mem_kind, &dfull));
    GSUM_NCCL(c, N, N->AllGather(dblock, drecv, (size_t)(n_q * per), ncclDouble, (ncclComm_t)c->comm, c->stream));   // the single collective
    grid_unshard_kernel<<<(unsigned)((n_q * n_ls + 255) / 256), 256, 0, c->stream>>>((const double *)drecv, n_q, per, P, n_ls, (double *)dfull);
    LAUNCHED(c, 1);
    if (post || lse) {
        const int64_t count = n_q * n_ls;
        GSUM_TRY(dev_out(c, WS_IO3, post, sizeof(double) * count, mem_kind, &dpost));
        if (!dpost) GSUM_TRY(gsum_ws(c, WS_IO3, sizeof(double) * count, &dpost));
        GSUM_TRY(dev_out(c, WS_MISC0, lse, sizeof(double), mem_kind, &dlse));
        if (!dlse) GSUM_TRY(gsum_ws(c, WS_MISC0, sizeof(double), &dlse));
        if (count >= 8192)
            grid_normalize_cluster_kernel<<<GN_CLUSTER, 1024, 0, c->stream>>>((const double *)dfull, count, (double *)dpost, (double *)dlse);
        else
            grid_normalize_kernel<<<1, 1024, 0, c->stream>>>((const double *)dfull, count, (double *)dpost, (double *)dlse);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, post, dpost, sizeof(double) * count, mem_kind));
        GSUM_TRY(dev_out_finish(c, lse, dlse, sizeof(double), mem_kind));
    }
    GSUM_TRY(dev_out_finish(c, ll_full, dfull, sizeof(double) * n_q * n_ls, mem_kind));
    return finish(c, mem_kind);
}
// Sum of int64 counts over the ranks (coverage counts of sharded posterior draws, gsum/diagnostics.py:161-171): in place.
extern "C" int gsum_comm_allreduce_counts(gsum_ctx *c, int64_t *counts, int64_t n, int32_t mem_kind) {
    GSUM_RANGE("gsum_comm_allreduce_counts");
    if (!c || !counts || n <= 0) return gsum_fail(c, -1, "gsum_comm_allreduce_counts: bad argument");
    if (!c->comm) return gsum_fail(c, -1, "gsum_comm_allreduce_counts: no communicator (gsum_comm_init)");
    GsumNccl *N = gsum_nccl(c);
    if (!N) return -110;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const void *din;
    GSUM_TRY(dev_in(c, WS_IO0, counts, sizeof(int64_t) * n, mem_kind, &din));
    GSUM_NCCL(c, N, N->AllReduce(din, (void *)din, (size_t)n, ncclInt64, ncclSum, (ncclComm_t)c->comm, c->stream));
    GSUM_TRY(dev_out_finish(c, counts, din, sizeof(int64_t) * n, mem_kind));
    return finish(c, mem_kind);
}

// ---- shared plumbing for solves against an existing factor ------------------------------------------------
static void launch_transpose_in(gsum_ctx *c, const double *src, int64_t n, int64_t m, const double *sub, const int32_t *perm,
                                int flip, double sign, double *dst, int64_t dst_ld, int64_t rows_pad) {
    dim3 grid((unsigned)((rows_pad + 31) / 32), (unsigned)((dst_ld + 31) / 32));
    transpose_in_kernel<<<grid, 256, 0, c->stream>>>(src, n, m, sub, perm, flip, sign, dst, dst_ld, rows_pad);
    LAUNCHED(c, 1);
}
static void launch_transpose_out(gsum_ctx *c, const double *src, int64_t src_ld, int64_t n, int64_t m, int flip, double scale,
                                 const double *add, double *dst) {
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
    transpose_out_kernel<<<grid, 256, 0, c->stream>>>(src, src_ld, n, m, flip, scale, add, dst);
    LAUNCHED(c, 1);
}
static BorderedBatch solve_desc(double *F, double *W, int64_t np, int T, int64_t rows_pad, int64_t rows_used = 0) {
    BorderedBatch P;
    P.border_used = (int)rows_used;
    P.A = F; P.ld = np; P.bstride = np * np; P.W = W; P.wstride = rows_pad * np;
    P.T = T; P.Trows = T + (int)(rows_pad / GSUM_TILE); P.info = nullptr; P.logdet_part = nullptr; P.n = (int)np;
    return P;
}
// contiguous factor L (n,n) -> padded (np x np) with identity padding
static int pad_factor(gsum_ctx *c, const double *dL, int64_t n, double *dF) {
    const int64_t np = gsum_pad64(n);
    dim3 g((unsigned)((np + 255) / 256), (unsigned)np, 1);
    pad_in_kernel<<<g, 256, 0, c->stream>>>(dL, n, dF, np, np * np, np);
    LAUNCHED(c, 1);
    return 0;
}
// Lt[i][j] = L[n-1-j][n-1-i]: the lower-triangular matrix whose forward substitution is L^T's back substitution
__global__ void flip_transpose_kernel(const double *__restrict__ L, int64_t n, double *__restrict__ F, int64_t np) {
    const int64_t i = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < np; j += (int64_t)gridDim.x * blockDim.x)
        F[i * np + j] = (i < n && j < n) ? ((j <= i) ? L[(n - 1 - j) * n + (n - 1 - i)] : 0.0) : (i == j ? 1.0 : 0.0);
}
__global__ void flip_rows_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t ld, int64_t n) {
    const int64_t r = blockIdx.x;
    for (int64_t x = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; x < ld; x += (int64_t)gridDim.y * blockDim.x)
        dst[r * ld + x] = x < n ? src[r * ld + (n - 1 - x)] : 0.0;
}

// ---- K3 -------------------------------------------------------------------------------------------------
extern "C" int gsum_cho_solve(gsum_ctx *c, const double *L, int64_t n, double *B, int64_t nrhs, int32_t forward_only,
                              int32_t mem_kind) {
    GSUM_RANGE("gsum_cho_solve");
    if (!c || !L || !B || n <= 0 || nrhs <= 0) return gsum_fail(c, -1, "gsum_cho_solve: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int fk = factor_kind(mem_kind);
    const int64_t np = gsum_pad64(n), rp = gsum_pad64(nrhs);
    const int T = (int)(np / GSUM_TILE);
    const void *dL, *dB;
    GSUM_TRY(dev_in(c, WS_IO0, L, sizeof(double) * n * n, fk, &dL));
    GSUM_TRY(dev_in(c, WS_IO1, B, sizeof(double) * n * nrhs, mem_kind, &dB));
    void *dF, *dW;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dF));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rp * np, &dW));
    pad_factor(c, (const double *)dL, n, (double *)dF);
    launch_transpose_in(c, (const double *)dB, n, nrhs, nullptr, nullptr, 0, 1.0, (double *)dW, np, rp);
    GSUM_TRY(solve_run(c, solve_desc((double *)dF, (double *)dW, np, T, rp, nrhs), 1));
    double *dBout = (double *)dB;     // in place (caller's device buffer or our staging copy)
    if (forward_only) {
        launch_transpose_out(c, (const double *)dW, np, n, nrhs, 0, 1.0, nullptr, dBout);
    } else {
        void *dW2;
        GSUM_TRY(gsum_ws(c, WS_MISC2, sizeof(double) * rp * np, &dW2));
        dim3 g1((unsigned)((np + 255) / 256), (unsigned)np);
        flip_transpose_kernel<<<g1, 256, 0, c->stream>>>((const double *)dL, n, (double *)dF, np);
        dim3 g2((unsigned)rp, (unsigned)((np + 255) / 256));
        flip_rows_kernel<<<g2, 256, 0, c->stream>>>((const double *)dW, (double *)dW2, np, n);
        LAUNCHED(c, 2);
        GSUM_TRY(solve_run(c, solve_desc((double *)dF, (double *)dW2, np, T, rp, nrhs), 1));
        launch_transpose_out(c, (const double *)dW2, np, n, nrhs, 1, 1.0, nullptr, dBout);
    }
    GSUM_TRY(dev_out_finish(c, B, dBout, sizeof(double) * n * nrhs, mem_kind));
    return finish(c, mem_kind);
}

// ---- gradient terms of the likelihood (gsum/models.py:957-1056) ------------------------------------------------
extern "C" int gsum_eigh(gsum_ctx *c, const double *A, int64_t n, double *w, double *V, int32_t *sweeps_out, int32_t mem_kind);
extern "C" int gsum_eig_solve(gsum_ctx *c, const double *w, const double *V, int64_t n, const double *Y, int64_t nrhs,
                              const double *mean, double *X, int32_t mode, int32_t mem_kind);

static int lml_grad_terms_impl(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *RHS, int32_t r,
                               const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                               double *G, double *H, double *tr, double *logdet, int32_t *info, int32_t mem_kind, int use_eig) {
    if (!c || !X || !RHS || !ls || !G || !H || !tr || !logdet || !info)
        return gsum_fail(c, -1, "gsum_lml_grad_terms: null argument");
    if (n <= 0 || d <= 0 || d > COV_MAXD || r < 1 || r > GRAD_MAXR || (ls_dim != 1 && ls_dim != d))
        return gsum_fail(c, -1, "gsum_lml_grad_terms: bad shape (d<=%d, r<=%d, ls_dim in {1,d})", COV_MAXD, GRAD_MAXR);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int P = ls_dim + 2;
    const void *dX, *dRHS, *dls;
    GSUM_TRY(dev_in(c, WS_G0, X, sizeof(double) * n * d, mem_kind, &dX));
    GSUM_TRY(dev_in(c, WS_G1, RHS, sizeof(double) * n * r, mem_kind, &dRHS));
    GSUM_TRY(dev_in(c, WS_G2, ls, sizeof(double) * ls_dim, mem_kind, &dls));
    void *dR, *dB, *dsm;
    const int64_t ldz = r + n;
    GSUM_TRY(gsum_ws(c, WS_G3, sizeof(double) * n * n, &dR));
    GSUM_TRY(gsum_ws(c, WS_G4, sizeof(double) * n * ldz, &dB));
    // small outputs and per-row partials (one set per chunk of 128 columns):
    // G (r*r) | H (P*r*r) | tr (P) | logdet (1) | Y (nchunk*P*n*GRAD_MAXR) | trow (nchunk*P*n) | info
    const int nchunk = (int)((n + 127) / 128);
    const size_t nsmall = (size_t)r * r + (size_t)P * r * r + P + 1, nrows = (size_t)nchunk * P * n;
    GSUM_TRY(gsum_ws(c, WS_G5, sizeof(double) * (nsmall + nrows * GRAD_MAXR + nrows) + 64, &dsm));
    double *dG = (double *)dsm, *dH = dG + r * r, *dtr = dH + (size_t)P * r * r, *dld = dtr + P, *dY = dld + 1,
           *dtrow = dY + nrows * GRAD_MAXR;
    int32_t *dinfo = (int32_t *)(dtrow + nrows);
    // R = c * rbf + noise I (diagonal exactly c + noise), then + nugget, as gsum/models.py:960-963
    GSUM_TRY(gsum_kernel_matrix(c, (const double *)dX, n, nullptr, 0, d, (const double *)dls, ls_dim, constant, noise, (double *)dR, GSUM_MEM_DEVICE));
    add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((double *)dR, n, n, nugget);
    LAUNCHED(c, 1);
    grad_stage_kernel<<<dim3((unsigned)((ldz + 255) / 256), (unsigned)n), 256, 0, c->stream>>>((const double *)dRHS, n, r, (double *)dB);
    LAUNCHED(c, 1);
    if (!use_eig) {
        GSUM_TRY(gsum_cholesky(c, (double *)dR, n, 1, dinfo, dld, GSUM_MEM_DEVICE));
        GSUM_TRY(gsum_cho_solve(c, (const double *)dR, n, (double *)dB, ldz, 0, GSUM_MEM_DEVICE));   // [Z | R^{-1}]
    } else {
        // decomposition='eig' (gsum/models.py:973-974, 480-484, 1019): R = Q diag(eig) Q^T on the device,
        // [Z | R^{-1}] = Q diag(1/eig) Q^T [RHS | I], logdet R = sum log eig (NaN for a non-positive eigenvalue, as numpy's)
        void *dwv, *dVv;
        GSUM_TRY(gsum_ws(c, WS_DETF, sizeof(double) * n, &dwv));
        GSUM_TRY(gsum_ws(c, WS_MKK, sizeof(double) * n * n, &dVv));
        const int rc = gsum_eigh(c, (const double *)dR, n, (double *)dwv, (double *)dVv, nullptr, GSUM_MEM_DEVICE);
        if (rc != 0) return rc;
        GSUM_TRY(gsum_eig_solve(c, (const double *)dwv, (const double *)dVv, n, (const double *)dB, ldz, nullptr, (double *)dB, 0, GSUM_MEM_DEVICE));
        std::vector<double> hwv(n);
        GSUM_CUDA(c, cudaMemcpyAsync(hwv.data(), dwv, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        double ld_host = 0.0;
        for (int64_t i = 0; i < n; i++) ld_host += log(hwv[i]);
        GSUM_CUDA(c, cudaMemcpy(dld, &ld_host, sizeof(double), cudaMemcpyHostToDevice));
        GSUM_CUDA(c, cudaMemsetAsync(dinfo, 0, sizeof(int32_t), c->stream));
    }
    // scaled coordinates (gsum_kernel_matrix left X / ls in WS_XS)
    const double *dXS = (const double *)c->ws[WS_XS];
    const size_t shbytes = sizeof(double) * (128 * COV_MAXD + 128 * GRAD_MAXR);
    grad_rows_kernel<<<dim3((unsigned)nchunk, (unsigned)P, (unsigned)nchunk), 128, shbytes, c->stream>>>(dXS, n, d, ls_dim, constant, noise,
                                                                                                      (const double *)dB, ldz, r, dY, dtrow);
    grad_reduce_kernel<<<dim3((unsigned)(P + 1), (unsigned)(r * r)), 256, 0, c->stream>>>((const double *)dB, ldz, (const double *)dRHS, dY, dtrow, n, r,
                                                                                          P, nchunk, dH, dtr, dG);
    LAUNCHED(c, 2);
    if (mem_kind == GSUM_MEM_DEVICE) {
        GSUM_CUDA(c, cudaMemcpyAsync(G, dG, sizeof(double) * r * r, cudaMemcpyDeviceToDevice, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(H, dH, sizeof(double) * P * r * r, cudaMemcpyDeviceToDevice, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(tr, dtr, sizeof(double) * P, cudaMemcpyDeviceToDevice, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(logdet, dld, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(info, dinfo, sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        GSUM_CUDA(c, cudaMemcpyAsync(G, dG, sizeof(double) * r * r, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(H, dH, sizeof(double) * P * r * r, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(tr, dtr, sizeof(double) * P, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(logdet, dld, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(info, dinfo, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    return finish(c, mem_kind);
}

extern "C" int gsum_lml_grad_terms(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *RHS, int32_t r,
                                   const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                                   double *G, double *H, double *tr, double *logdet, int32_t *info, int32_t mem_kind) {
    GSUM_RANGE("gsum_lml_grad_terms");
    return lml_grad_terms_impl(c, X, n, d, RHS, r, ls, ls_dim, constant, noise, nugget, G, H, tr, logdet, info, mem_kind, 0);
}
extern "C" int gsum_lml_grad_terms_eig(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *RHS, int32_t r,
                                       const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                                       double *G, double *H, double *tr, double *logdet, int32_t *info, int32_t mem_kind) {
    GSUM_RANGE("gsum_lml_grad_terms_eig");
    return lml_grad_terms_impl(c, X, n, d, RHS, r, ls, ls_dim, constant, noise, nugget, G, H, tr, logdet, info, mem_kind, 1);
}

// ---- fit --------------------------------------------------------------------------------------------------
struct gsum_fit {
    gsum_ctx *ctx;
    int64_t n, np; int d, n_c, T, ls_dim;
    double ls[COV_MAXD];
    double constant, noise, nugget;
    double post[7];          // center, disp, df, scale, cov_factor, lml, logdet
    double *dX, *dXS, *dy, *dL, *dls;   // device: X (n,d), X/ls, y (n,n_c), padded factor (np,np), ls
};

extern "C" int gsum_fit_destroy(gsum_fit *f) {
    if (!f) return 0;
    cudaSetDevice(f->ctx->device);
    cudaStream_t st = f->ctx->stream;
    if (f->dX) cudaFreeAsync(f->dX, st);
    if (f->dXS) cudaFreeAsync(f->dXS, st);
    if (f->dy) cudaFreeAsync(f->dy, st);
    if (f->dL) cudaFreeAsync(f->dL, st);
    if (f->dls) cudaFreeAsync(f->dls, st);
    delete f;
    return 0;
}

extern "C" int gsum_fit_create(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *y, int32_t n_c,
                               const double *ls, int32_t ls_dim, double constant, double noise, double nugget,
                               double center0, double disp0, double df0, double scale0, int32_t student, double *out7,
                               double *L_out, int32_t mem_kind, gsum_fit **fit) {
    GSUM_RANGE("gsum_fit_create");
    if (!c || !X || !y || !ls || !fit || n <= 0 || d <= 0 || d > COV_MAXD || (ls_dim != 1 && ls_dim != d) || n_c < 1 || n_c + 1 > LML_MAXR)
        return gsum_fail(c, -1, "gsum_fit_create: bad argument");
    *fit = nullptr;
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int64_t np = gsum_pad64(n);
    const int T = (int)(np / GSUM_TILE), R = n_c + 1;
    gsum_fit *f = new gsum_fit();
    memset(f, 0, sizeof(*f));
    f->ctx = c; f->n = n; f->np = np; f->d = d; f->n_c = n_c; f->T = T; f->ls_dim = ls_dim;
    f->constant = constant; f->noise = noise; f->nugget = nugget;
#define FIT_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { gsum_fit_destroy(f); return gsum_fail(c, -100, "CUDA error %s in gsum_fit_create", cudaGetErrorName(_e)); } } while (0)
    FIT_CUDA(cudaMallocAsync((void **)&f->dX, sizeof(double) * n * d, c->stream));
    FIT_CUDA(cudaMallocAsync((void **)&f->dXS, sizeof(double) * n * d, c->stream));
    FIT_CUDA(cudaMallocAsync((void **)&f->dy, sizeof(double) * n * n_c, c->stream));
    FIT_CUDA(cudaMallocAsync((void **)&f->dls, sizeof(double) * ls_dim, c->stream));
    FIT_CUDA(cudaMallocAsync((void **)&f->dL, sizeof(double) * (np + GSUM_TILE) * np, c->stream));      // factor + one border tile row (basis, y)
    const cudaMemcpyKind kin = mem_kind == GSUM_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    FIT_CUDA(cudaMemcpyAsync(f->dX, X, sizeof(double) * n * d, kin, c->stream));
    FIT_CUDA(cudaMemcpyAsync(f->dy, y, sizeof(double) * n * n_c, kin, c->stream));
    FIT_CUDA(cudaMemcpyAsync(f->dls, ls, sizeof(double) * ls_dim, kin, c->stream));
    if (mem_kind == GSUM_MEM_HOST) memcpy(f->ls, ls, sizeof(double) * ls_dim);
    else FIT_CUDA(cudaMemcpyAsync(f->ls, ls, sizeof(double) * ls_dim, cudaMemcpyDeviceToHost, c->stream));
    scale_coords(c, f->dX, f->dls, f->dXS, n, d, ls_dim, 1);
    CovArgs CA;
    CA.XS = f->dXS; CA.n = n; CA.d = d; CA.constant = constant; CA.noise = noise; CA.nugget = nugget;
    CA.A = f->dL; CA.ld = np; CA.bstride = (np + GSUM_TILE) * np; CA.T = T;
    CA.tiles_per_cta = cov_tiles_per_cta(T * (T + 1) / 2, 1, c->sm_count);
    cov_sym_kernel<<<dim3((unsigned)((T * (T + 1) / 2 + CA.tiles_per_cta - 1) / CA.tiles_per_cta), 1), 256, 0, c->stream>>>(CA);
    // border rows: basis (ones) then the n_c curves, transposed — same staging as the grid path with ref = 1, Q absent
    void *dones, *dord, *drhs, *dgram, *dll, *dpost;
    GSUM_TRY(gsum_ws(c, WS_REF, sizeof(double) * n, &dones));
    GSUM_TRY(gsum_ws(c, WS_ORD, sizeof(int32_t) * n_c, &dord));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * R * n, &drhs));
    GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * R * R, &dgram));
    GSUM_TRY(gsum_ws(c, WS_LL, sizeof(double) * 2, &dll));
    GSUM_TRY(gsum_ws(c, WS_MISC0, sizeof(double) * 8, &dpost));
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((double *)dones, n, 1.0);
    FIT_CUDA(cudaMemsetAsync(dord, 0, sizeof(int32_t) * n_c, c->stream));
    stage_rhs_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)R), 256, 0, c->stream>>>((double *)drhs, n, f->dy, (const double *)dones, nullptr,
                                                                                             (const int32_t *)dord, n, n_c, 1);
    border_fill_kernel<<<dim3(GSUM_TILE, (unsigned)((np + 255) / 256), 1), 256, 0, c->stream>>>(f->dL, np, (np + GSUM_TILE) * np, T, GSUM_TILE,
                                                                                                (const double *)drhs, R, n, n);
    LAUNCHED(c, 4);
    int *dinfo; double *dpart;
    c->prof_border_rows = R;
    GSUM_TRY(factor_bordered(c, f->dL, n, T + 1, 1, &dinfo, &dpart));
    c->prof_border_rows = 0;
    switch (R) {
#define GRAM_CASE(RR) case RR: launch_gram<RR>(c, f->dL, np, (np + GSUM_TILE) * np, T, n, (double *)dgram, 1, 1); break;
        GRAM_CASE(2) GRAM_CASE(3) GRAM_CASE(4) GRAM_CASE(5) GRAM_CASE(6) GRAM_CASE(7) GRAM_CASE(8)
#undef GRAM_CASE
        default: gram_rows_generic_kernel<<<dim3(1, 1), 256, 0, c->stream>>>(f->dL, np, (np + GSUM_TILE) * np, T, n, (double *)dgram, 1, R);
    }
    LmlCellArgs LA;
    LA.G = (const double *)dgram; LA.logdet_part = dpart; LA.info = dinfo; LA.T = T; LA.R = R; LA.n = n; LA.n_l = 1; LA.n_q = 1;
    LA.separable = 0; LA.Q = nullptr; LA.orders = (const int32_t *)dord; LA.detf = nullptr;
    LA.center0 = center0; LA.disp0 = disp0; LA.df0 = df0; LA.scale0 = scale0; LA.student = student;
    LA.ll = (double *)dll; LA.logdet_out = (double *)dll + 1; LA.post = (double *)dpost;
    lml_cell_chunk_kernel<<<1, 32, 0, c->stream>>>(LA, 0, 1);
    LAUNCHED(c, 2);
    int hinfo = 0; double hll[2], hpost[5];
    FIT_CUDA(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FIT_CUDA(cudaMemcpyAsync(hll, dll, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
    FIT_CUDA(cudaMemcpyAsync(hpost, dpost, sizeof(double) * 5, cudaMemcpyDeviceToHost, c->stream));
    FIT_CUDA(cudaStreamSynchronize(c->stream));
    if (hinfo != 0) {
        gsum_fit_destroy(f);
        return gsum_fail(c, hinfo, "gsum_fit_create: correlation matrix is not positive definite (leading minor %d)", hinfo);
    }
    f->post[0] = hpost[0]; f->post[1] = hpost[1]; f->post[2] = hpost[2]; f->post[3] = sqrt(hpost[3]); f->post[4] = hpost[4];
    f->post[5] = hll[0]; f->post[6] = hll[1];
    if (out7) {
        if (mem_kind == GSUM_MEM_HOST) memcpy(out7, f->post, sizeof(f->post));
        else FIT_CUDA(cudaMemcpyAsync(out7, f->post, sizeof(f->post), cudaMemcpyHostToDevice, c->stream));
    }
    if (L_out) {
        void *dLo;
        GSUM_TRY(dev_out(c, WS_IO0, L_out, sizeof(double) * n * n, mem_kind, &dLo));
        pad_out_lower_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n, 1), 256, 0, c->stream>>>(f->dL, np, 0, (double *)dLo, n);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, L_out, dLo, sizeof(double) * n * n, mem_kind));
    }
    // the factor's strict upper tiles are never read by the tile kernels, but draws / exports want exact zeros there
    *fit = f;
    return finish(c, GSUM_MEM_HOST);
#undef FIT_CUDA
}

// ---- predict / process covariance ---------------------------------------------------------------------------
static GeoSum make_geosum(const double *q, double start, double end, const int32_t *excluded, int n_excluded) {
    GeoSum g; memset(&g, 0, sizeof(g));
    g.enabled = q != nullptr; g.start = start; g.end = end;
    g.n_excl = n_excluded > 8 ? 8 : (n_excluded < 0 ? 0 : n_excluded);
    for (int i = 0; i < g.n_excl; i++) g.excl[i] = excluded[i];
    return g;
}
static void launch_cross(gsum_ctx *c, const double *xs1, int64_t n1, const double *xs2, int64_t n2, int d, double constant, double diag_add,
                         int sym_diag, double *out, int64_t ldo, int64_t rows_out, int64_t cols_out) {
    CrossArgs P;
    P.XS1 = xs1; P.n1 = n1; P.XS2 = xs2; P.n2 = n2; P.d = d; P.constant = constant; P.diag_add = diag_add; P.sym_diag = sym_diag;
    P.out = out; P.ldo = ldo; P.rows_out = rows_out; P.cols_out = cols_out;
    dim3 grid((unsigned)((cols_out + 63) / 64), (unsigned)((rows_out + 63) / 64));
    cov_cross_kernel<<<grid, 256, 0, c->stream>>>(P);
    LAUNCHED(c, 1);
}
static void launch_scale_cov(gsum_ctx *c, double *K, int64_t ld, int64_t rows, int64_t cols, const double *sr, const double *sc,
                             const double *qr, const double *qc, const GeoSum &g, double factor, double kadd = 0.0) {
    dim3 grid((unsigned)rows, (unsigned)((cols + 255) / 256));
    scale_cov_kernel<<<grid, 256, 0, c->stream>>>(K, ld, rows, cols, sr, sc, qr, qc, g, factor, kadd);
    LAUNCHED(c, 1);
}
__global__ void pad_identity_kernel(double *F, int64_t np, int64_t n) {
    const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) F[i * np + i] = 1.0;
}
// var_out[j] = factor * (diag[j] - sq[j] + add)
__global__ void finish_var_kernel(const double *__restrict__ sq, int64_t m, double diag_const, const double *__restrict__ sc,
                                  const double *__restrict__ q, GeoSum g, double kfactor, double factor, double add, double *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double dj = diag_const;
    if (sc || q) {                         // truncation: ((s s) gs(q q)) * (kfactor * c)
        const double s = sc ? sc[j] : 1.0;
        const double gs = (g.enabled && q) ? geo_sum(g, q[j] * q[j]) : 1.0;
        dj = ((s * s) * gs) * (kfactor * diag_const);
    } else dj = kfactor * diag_const;
    out[j] = factor * ((dj - sq[j]) + add);
}
// out (m x m) <- factor * sym(C lower) (+ factor * add on the diagonal)
__global__ void cov_out_kernel(const double *__restrict__ C, int64_t ldc, int64_t m, double factor, double add, double *__restrict__ out) {
    const int64_t r = blockIdx.x;
    for (int64_t cc = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; cc < m; cc += (int64_t)gridDim.y * blockDim.x) {
        double v = r >= cc ? C[r * ldc + cc] : C[cc * ldc + r];
        if (r == cc) v += add;
        out[r * m + cc] = factor * v;
    }
}

extern "C" int gsum_predict(gsum_ctx *c, gsum_fit *f, const gsum_predict_args *a, int32_t mem_kind) {
    GSUM_RANGE("gsum_predict");
    if (!c || !f || !a || !a->Xnew || a->m <= 0) return gsum_fail(c, -1, "gsum_predict: bad argument");
    if (a->want != GSUM_PREDICT_MEAN && a->want != GSUM_PREDICT_VAR && a->want != GSUM_PREDICT_COV)
        return gsum_fail(c, -1, "gsum_predict: bad `want`");
    if ((a->want != GSUM_PREDICT_MEAN) && !a->var_out) return gsum_fail(c, -1, "gsum_predict: var_out is NULL");
    if (a->yc && a->n_y < 1) return gsum_fail(c, -1, "gsum_predict: n_y < 1");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int d = f->d;
    const int64_t m = a->m, mp = gsum_pad64(m);
    const bool own_cond = a->Xc != nullptr;
    const int64_t n = own_cond ? a->n_cond : f->n;
    if (n <= 0) return gsum_fail(c, -1, "gsum_predict: n_cond <= 0");
    const int64_t np = gsum_pad64(n);
    const int T = (int)(np / GSUM_TILE);
    const int trunc = a->truncation != 0;
    const int n_y = a->yc ? a->n_y : f->n_c;
    const int want_basis = a->cond_basis_out != nullptr;
    const int64_t yrows = n_y + (want_basis ? 1 : 0);
    const int64_t yp = gsum_pad64(yrows);
    const double cov_factor = f->post[4];
    const GeoSum g = make_geosum(a->q_old, a->gs_start, a->gs_end, a->excluded, a->n_excluded);

    // inputs
    const void *dXn, *dXc = nullptr, *dyc, *dmo, *dmn, *dbo, *dbn, *dso, *dsn, *dqo, *dqn;
    GSUM_TRY(dev_in(c, WS_X, a->Xnew, sizeof(double) * m * d, mem_kind, &dXn));
    if (own_cond) GSUM_TRY(dev_in(c, WS_IO0, a->Xc, sizeof(double) * n * d, mem_kind, &dXc));
    GSUM_TRY(dev_in(c, WS_DY, a->yc, sizeof(double) * n * n_y, mem_kind, &dyc));
    GSUM_TRY(dev_in(c, WS_REF, a->mean_old, sizeof(double) * n, mem_kind, &dmo));
    GSUM_TRY(dev_in(c, WS_Q, a->mean_new, sizeof(double) * m, mem_kind, &dmn));
    GSUM_TRY(dev_in(c, WS_DETF, a->basis_old, sizeof(double) * n, mem_kind, &dbo));
    GSUM_TRY(dev_in(c, WS_ORD, a->basis_new, sizeof(double) * m, mem_kind, &dbn));
    GSUM_TRY(dev_in(c, WS_IO1, a->sc_old, sizeof(double) * n, mem_kind, &dso));
    GSUM_TRY(dev_in(c, WS_IO2, a->sc_new, sizeof(double) * m, mem_kind, &dsn));
    GSUM_TRY(dev_in(c, WS_IO3, a->q_old, sizeof(double) * n, mem_kind, &dqo));
    GSUM_TRY(dev_in(c, WS_MISC0, a->q_new, sizeof(double) * m, mem_kind, &dqn));
    if (!a->yc) dyc = f->dy;
    if (want_basis && (!a->basis_old || !a->basis_new)) return gsum_fail(c, -1, "gsum_predict: cond_basis_out needs basis_old and basis_new");

    // scaled coordinates
    void *dxs;
    GSUM_TRY(gsum_ws(c, WS_XS, sizeof(double) * (m + n) * d, &dxs));
    double *xsn = (double *)dxs, *xso = xsn + m * d;
    scale_coords(c, (const double *)dXn, f->dls, xsn, m, d, f->ls_dim, 1);
    const double *xs_old = f->dXS;
    if (own_cond) { scale_coords(c, (const double *)dXc, f->dls, xso, n, d, f->ls_dim, 1); xs_old = xso; }

    // conditioning factor
    double *F;
    if (!own_cond && !trunc) F = f->dL;
    else {
        void *dF;
        GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dF));
        F = (double *)dF;
        if (trunc) {
            launch_cross(c, xs_old, n, xs_old, n, d, f->constant, 0.0, 0, F, np, np, np);
            launch_scale_cov(c, F, np, n, n, (const double *)dso, (const double *)dso, (const double *)dqo, (const double *)dqo, g, cov_factor, a->kernel_add);
        } else {
            // kernel_(Xc) + nugget I  (gsum/models.py:807): diagonal = (c + noise) + nugget
            launch_cross(c, xs_old, n, xs_old, n, d, f->constant, f->noise, 1, F, np, np, np);
            add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(F, np, n, f->nugget);
            LAUNCHED(c, 1);
        }
        pad_identity_kernel<<<(unsigned)((np - n + 255) / 256 + 1), 256, 0, c->stream>>>(F, np, n);
        LAUNCHED(c, 1);
        int *dinfo; double *dpart;
        GSUM_TRY(factor_bordered(c, F, n, T, 1, &dinfo, &dpart));
        int hinfo = 0;
        GSUM_CUDA(c, cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hinfo != 0) return gsum_fail(c, hinfo, "gsum_predict: conditioning covariance is not positive definite (leading minor %d)", hinfo);
    }

    // border rows: K_no (m rows) | (yc - mean_old)^T | basis_old^T
    const int64_t rows_pad = mp + yp;
    void *dW;
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rows_pad * np, &dW));
    double *W = (double *)dW, *Wy = W + mp * np;
    launch_cross(c, xsn, m, xs_old, n, d, f->constant, 0.0, 0, W, np, mp, np);
    if (trunc) launch_scale_cov(c, W, np, m, n, (const double *)dsn, (const double *)dso, (const double *)dqn, (const double *)dqo, g, cov_factor, a->kernel_add);
    launch_transpose_in(c, (const double *)dyc, n, n_y, (const double *)dmo, nullptr, 0, 1.0, Wy, np, want_basis ? n_y : yp);
    if (want_basis) launch_transpose_in(c, (const double *)dbo, n, 1, nullptr, nullptr, 0, 1.0, Wy + (int64_t)n_y * np, np, yp - n_y);
    GSUM_TRY(solve_run(c, solve_desc(F, W, np, T, rows_pad), 1));

    // mean
    if (a->mean_out) {
        void *dmean;
        GSUM_TRY(dev_out(c, WS_LL, a->mean_out, sizeof(double) * m * n_y, mem_kind, &dmean));
        rows_dot_kernel<<<(unsigned)((m + 7) / 8), 256, 0, c->stream>>>(W, np, m, n, Wy, n_y, (const double *)dmn, 1.0, (double *)dmean, n_y);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, a->mean_out, dmean, sizeof(double) * m * n_y, mem_kind));
    }
    if (want_basis) {
        void *dcb;
        GSUM_TRY(dev_out(c, WS_MISC1, a->cond_basis_out, sizeof(double) * m, mem_kind, &dcb));
        rows_dot_kernel<<<(unsigned)((m + 7) / 8), 256, 0, c->stream>>>(W, np, m, n, Wy + (int64_t)n_y * np, 1, (const double *)dbn, -1.0, (double *)dcb, 1);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, a->cond_basis_out, dcb, sizeof(double) * m, mem_kind));
    }
    const GeoSum gn = make_geosum(a->q_new, a->gs_start, a->gs_end, a->excluded, a->n_excluded);
    const double out_factor = trunc ? 1.0 : cov_factor;
    const double add = (a->pred_noise && !trunc) ? f->nugget : 0.0;
    if (a->want == GSUM_PREDICT_VAR) {
        void *dsq, *dvar;
        GSUM_TRY(gsum_ws(c, WS_MISC2, sizeof(double) * m, &dsq));
        GSUM_TRY(dev_out(c, WS_MISC3, a->var_out, sizeof(double) * m, mem_kind, &dvar));
        rows_sqnorm_kernel<<<(unsigned)((m + 7) / 8), 256, 0, c->stream>>>(W, np, m, n, (double *)dsq);
        const double diag_const = trunc ? (f->constant + a->kernel_add) : (f->constant + f->noise);
        finish_var_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>((const double *)dsq, m, diag_const, trunc ? (const double *)dsn : nullptr,
                                                                              trunc ? (const double *)dqn : nullptr, gn, trunc ? cov_factor : 1.0,
                                                                              out_factor, add, (double *)dvar);
        LAUNCHED(c, 2);
        GSUM_TRY(dev_out_finish(c, a->var_out, dvar, sizeof(double) * m, mem_kind));
    } else if (a->want == GSUM_PREDICT_COV) {
        void *dC, *dcov;
        GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * mp * mp, &dC));
        GSUM_TRY(dev_out(c, WS_MISC3, a->var_out, sizeof(double) * m * m, mem_kind, &dcov));
        if (trunc) {
            launch_cross(c, xsn, m, xsn, m, d, f->constant, 0.0, 0, (double *)dC, mp, mp, mp);
            launch_scale_cov(c, (double *)dC, mp, m, m, (const double *)dsn, (const double *)dsn, (const double *)dqn, (const double *)dqn, gn, cov_factor, a->kernel_add);
        } else {
            launch_cross(c, xsn, m, xsn, m, d, f->constant, f->noise, 1, (double *)dC, mp, mp, mp);
        }
        SchurArgs S;
        S.W = W; S.ld = np; S.bstride = 0; S.T = T; S.C = (double *)dC; S.ldc = mp; S.cstride = 0; S.lower_only = 1;
        GSUM_TRY(schur_run(c, S, (int)(mp / GSUM_TILE), 1));
        cov_out_kernel<<<dim3((unsigned)m, (unsigned)((m + 255) / 256)), 256, 0, c->stream>>>((const double *)dC, mp, m, out_factor, add, (double *)dcov);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, a->var_out, dcov, sizeof(double) * m * m, mem_kind));
    }
    return finish(c, mem_kind);
}

extern "C" int gsum_process_cov(gsum_ctx *c, int32_t d, const double *ls, int32_t ls_dim, double constant, double noise,
                                const double *X1, int64_t n1, const double *X2, int64_t n2, const double *sc1, const double *sc2, const double *q1, const double *q2, double gs_start,
                                double gs_end, const int32_t *excluded, int32_t n_excluded, double factor, double kernel_add,
                                double *out, int32_t mem_kind) {
    GSUM_RANGE("gsum_process_cov");
    if (!c || !ls || !X1 || !out || n1 <= 0 || d <= 0 || d > COV_MAXD || (ls_dim != 1 && ls_dim != d))
        return gsum_fail(c, -1, "gsum_process_cov: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const bool sym = X2 == nullptr;
    const void *dls;
    GSUM_TRY(dev_in(c, WS_LS, ls, sizeof(double) * ls_dim, mem_kind, &dls));
    if (sym) { n2 = n1; sc2 = sc1; q2 = q1; }
    const void *dX1, *dX2 = nullptr, *ds1, *ds2, *dq1, *dq2;
    GSUM_TRY(dev_in(c, WS_X, X1, sizeof(double) * n1 * d, mem_kind, &dX1));
    if (!sym) GSUM_TRY(dev_in(c, WS_IO0, X2, sizeof(double) * n2 * d, mem_kind, &dX2));
    GSUM_TRY(dev_in(c, WS_IO1, sc1, sizeof(double) * n1, mem_kind, &ds1));
    GSUM_TRY(dev_in(c, WS_IO2, sc2, sizeof(double) * n2, mem_kind, &ds2));
    GSUM_TRY(dev_in(c, WS_IO3, q1, sizeof(double) * n1, mem_kind, &dq1));
    GSUM_TRY(dev_in(c, WS_MISC0, q2, sizeof(double) * n2, mem_kind, &dq2));
    void *dxs, *dout;
    GSUM_TRY(gsum_ws(c, WS_XS, sizeof(double) * (n1 + n2) * d, &dxs));
    double *xs1 = (double *)dxs, *xs2 = xs1 + n1 * d;
    scale_coords(c, (const double *)dX1, (const double *)dls, xs1, n1, d, ls_dim, 1);
    if (!sym) scale_coords(c, (const double *)dX2, (const double *)dls, xs2, n2, d, ls_dim, 1);
    GSUM_TRY(dev_out(c, WS_MAT, out, sizeof(double) * n1 * n2, mem_kind, &dout));
    launch_cross(c, xs1, n1, sym ? xs1 : xs2, n2, d, constant, noise, sym ? 1 : 0, (double *)dout, n2, n1, n2);
    const GeoSum g = make_geosum(q1, gs_start, gs_end, excluded, n_excluded);
    launch_scale_cov(c, (double *)dout, n2, n1, n2, (const double *)ds1, (const double *)ds2, (const double *)dq1, (const double *)dq2, g, factor, kernel_add);
    GSUM_TRY(dev_out_finish(c, out, dout, sizeof(double) * n1 * n2, mem_kind));
    return finish(c, mem_kind);
}

// ---- diagnostics ---------------------------------------------------------------------------------------------
// ---- SURVEY.md 8(f).4 -------------------------------------------------------------------------------------------------
static int pw_args(gsum_ctx *c, PointwiseArgs &P, const double *y, int64_t n, int n_o, const int32_t *orders, const int32_t *mask,
                   const int32_t *excluded, int n_ex, double df0, double scale0, int mem_kind, int *n_m_out) {
    const void *dy, *dord, *dmask, *dex = nullptr;
    GSUM_TRY(dev_in(c, WS_DY, y, sizeof(double) * n * n_o, mem_kind, &dy));
    GSUM_TRY(dev_in(c, WS_ORD, orders, sizeof(int) * n_o, mem_kind, &dord));
    GSUM_TRY(dev_in(c, WS_MISC0, mask, sizeof(int) * n_o, mem_kind, &dmask));
    if (n_ex > 0) GSUM_TRY(dev_in(c, WS_MISC1, excluded, sizeof(int) * n_ex, mem_kind, &dex));
    P.y = (const double *)dy; P.orders = (const int *)dord; P.mask = (const int *)dmask; P.excluded = (const int *)dex; P.n_ex = n_ex;
    P.n = n; P.n_o = n_o; P.df0 = df0; P.scale0 = scale0;
    int n_m = 0;
    if (mem_kind == GSUM_MEM_HOST) { for (int k = 0; k < n_o; k++) n_m += mask[k] ? 1 : 0; }
    else {
        int hm[PW_MAXO];
        GSUM_CUDA(c, cudaMemcpyAsync(hm, mask, sizeof(int) * n_o, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < n_o; k++) n_m += hm[k] ? 1 : 0;
    }
    *n_m_out = n_m;
    return 0;
}
extern "C" int gsum_pointwise_fit(gsum_ctx *c, const double *y, int64_t n, int32_t n_o, const int32_t *orders, const int32_t *mask,
                                  const int32_t *excluded, int32_t n_ex, const double *ratio, const double *ref, double df0,
                                  double scale0, double *coeffs, double *scale, double *trunc_scale, int32_t mem_kind) {
    GSUM_RANGE("gsum_pointwise_fit");
    if (!c || !y || !orders || !mask || !ratio || !ref || !coeffs || !scale || !trunc_scale || n <= 0 || n_o <= 0 || n_o > PW_MAXO || n_ex < 0 ||
        (n_ex > 0 && !excluded))
        return gsum_fail(c, -1, "gsum_pointwise_fit: bad argument (n_o <= %d)", PW_MAXO);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    PointwiseArgs P; int n_m;
    GSUM_TRY(pw_args(c, P, y, n, n_o, orders, mask, excluded, n_ex, df0, scale0, mem_kind, &n_m));
    if (n_m == 0) return gsum_fail(c, -1, "gsum_pointwise_fit: every order is excluded");
    const void *dq, *dr;
    GSUM_TRY(dev_in(c, WS_Q, ratio, sizeof(double) * n, mem_kind, &dq));
    GSUM_TRY(dev_in(c, WS_REF, ref, sizeof(double) * n, mem_kind, &dr));
    void *dc, *ds, *dt;
    GSUM_TRY(dev_out(c, WS_IO0, coeffs, sizeof(double) * n * n_m, mem_kind, &dc));
    GSUM_TRY(dev_out(c, WS_IO1, scale, sizeof(double) * n, mem_kind, &ds));
    GSUM_TRY(dev_out(c, WS_IO2, trunc_scale, sizeof(double) * n * n_m, mem_kind, &dt));
    pointwise_fit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(P, (const double *)dq, (const double *)dr, (double *)dc, (double *)ds,
                                                                             (double *)dt, n_m);
    LAUNCHED(c, 1);
    GSUM_TRY(dev_out_finish(c, coeffs, dc, sizeof(double) * n * n_m, mem_kind));
    GSUM_TRY(dev_out_finish(c, scale, ds, sizeof(double) * n, mem_kind));
    GSUM_TRY(dev_out_finish(c, trunc_scale, dt, sizeof(double) * n * n_m, mem_kind));
    return finish(c, mem_kind);
}
extern "C" int gsum_pointwise_loglike(gsum_ctx *c, const double *y, int64_t n, int32_t n_o, const int32_t *orders, const int32_t *mask,
                                      const double *ratios, int64_t n_r, int64_t n_rat, const double *ref, int64_t n_ref, double df0,
                                      double scale0, double *S1, double *S2, int32_t mem_kind) {
    GSUM_RANGE("gsum_pointwise_loglike");
    if (!c || !y || !orders || !mask || !ratios || !ref || !S1 || !S2 || n <= 0 || n_o <= 0 || n_o > PW_MAXO || n_r <= 0 ||
        (n_rat != 1 && n_rat != n) || (n_ref != 1 && n_ref != n))
        return gsum_fail(c, -1, "gsum_pointwise_loglike: bad argument (ratio / ref must have 1 or n entries)");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    PointwiseArgs P; int n_m;
    GSUM_TRY(pw_args(c, P, y, n, n_o, orders, mask, nullptr, 0, df0, scale0, mem_kind, &n_m));
    const void *dq, *dr, *dord_h = nullptr;
    (void)dord_h;
    GSUM_TRY(dev_in(c, WS_Q, ratios, sizeof(double) * n_r * n_rat, mem_kind, &dq));
    GSUM_TRY(dev_in(c, WS_REF, ref, sizeof(double) * n_ref, mem_kind, &dr));
    // sum of the kept orders (host copy of the two small arrays when they live on the device)
    int ho[PW_MAXO], hm[PW_MAXO];
    if (mem_kind == GSUM_MEM_HOST) { memcpy(ho, orders, sizeof(int) * n_o); memcpy(hm, mask, sizeof(int) * n_o); }
    else {
        GSUM_CUDA(c, cudaMemcpyAsync(ho, orders, sizeof(int) * n_o, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaMemcpyAsync(hm, mask, sizeof(int) * n_o, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    double so = 0.0;
    for (int k = 0; k < n_o; k++) if (hm[k]) so += (double)ho[k];
    void *d1, *d2;
    GSUM_TRY(dev_out(c, WS_IO0, S1, sizeof(double) * n_r, mem_kind, &d1));
    GSUM_TRY(dev_out(c, WS_IO1, S2, sizeof(double) * n_r, mem_kind, &d2));
    pointwise_loglike_kernel<<<(unsigned)n_r, 256, 0, c->stream>>>(P, (const double *)dq, (int)n_rat, (const double *)dr, (int)n_ref, n_m, so,
                                                                   (double *)d1, (double *)d2);
    LAUNCHED(c, 1);
    GSUM_TRY(dev_out_finish(c, S1, d1, sizeof(double) * n_r, mem_kind));
    GSUM_TRY(dev_out_finish(c, S2, d2, sizeof(double) * n_r, mem_kind));
    return finish(c, mem_kind);
}
extern "C" int gsum_variogram_bins(gsum_ctx *c, const double *X, int64_t n, int32_t d, const double *z, int32_t ncurves,
                                   const double *bounds, int32_t nbnd, int32_t *bin_grid, double *hij, int32_t *bin_idx, double *dij,
                                   int64_t *counts, double *hsum, double *dsum, int32_t mem_kind) {
    GSUM_RANGE("gsum_variogram_bins");
    if (!c || !X || !z || !bounds || !bin_grid || !hij || !bin_idx || !dij || !counts || !hsum || !dsum || n < 2 || d <= 0 || ncurves <= 0 || nbnd <= 0)
        return gsum_fail(c, -1, "gsum_variogram_bins: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int64_t np_ = n * (n - 1) / 2;
    const int nb = nbnd + 1;
    const void *dX, *dz, *db;
    GSUM_TRY(dev_in(c, WS_X, X, sizeof(double) * n * d, mem_kind, &dX));
    GSUM_TRY(dev_in(c, WS_DY, z, sizeof(double) * n * ncurves, mem_kind, &dz));
    GSUM_TRY(dev_in(c, WS_Q, bounds, sizeof(double) * nbnd, mem_kind, &db));
    void *dgrid, *dh, *dbi, *dd, *dcnt, *dhs, *dds;
    GSUM_TRY(dev_out(c, WS_MAT, bin_grid, sizeof(int) * n * n, mem_kind, &dgrid));
    GSUM_TRY(dev_out(c, WS_IO0, hij, sizeof(double) * np_, mem_kind, &dh));
    GSUM_TRY(dev_out(c, WS_IO1, bin_idx, sizeof(int) * np_, mem_kind, &dbi));
    GSUM_TRY(dev_out(c, WS_IO2, dij, sizeof(double) * np_ * ncurves, mem_kind, &dd));
    GSUM_TRY(dev_out(c, WS_MISC0, counts, sizeof(long long) * nb, mem_kind, &dcnt));
    GSUM_TRY(dev_out(c, WS_MISC1, hsum, sizeof(double) * nb, mem_kind, &dhs));
    GSUM_TRY(dev_out(c, WS_MISC2, dsum, sizeof(double) * nb * ncurves, mem_kind, &dds));
    variogram_grid_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, c->stream>>>((const double *)dX, n, d, (const double *)db, nbnd, (int *)dgrid);
    variogram_pairs_kernel<<<(unsigned)((np_ + 255) / 256), 256, 0, c->stream>>>((const double *)dX, (const double *)dz, n, d, ncurves, (const double *)db,
                                                                                 nbnd, (double *)dh, (int *)dbi, (double *)dd);
    variogram_binsum_kernel<<<nb, 256, 0, c->stream>>>((const double *)dh, (const int *)dbi, (const double *)dd, np_, ncurves, (long long *)dcnt,
                                                       (double *)dhs, (double *)dds);
    LAUNCHED(c, 3);
    GSUM_TRY(dev_out_finish(c, bin_grid, dgrid, sizeof(int) * n * n, mem_kind));
    GSUM_TRY(dev_out_finish(c, hij, dh, sizeof(double) * np_, mem_kind));
    GSUM_TRY(dev_out_finish(c, bin_idx, dbi, sizeof(int) * np_, mem_kind));
    GSUM_TRY(dev_out_finish(c, dij, dd, sizeof(double) * np_ * ncurves, mem_kind));
    GSUM_TRY(dev_out_finish(c, counts, dcnt, sizeof(long long) * nb, mem_kind));
    GSUM_TRY(dev_out_finish(c, hsum, dhs, sizeof(double) * nb, mem_kind));
    GSUM_TRY(dev_out_finish(c, dsum, dds, sizeof(double) * nb * ncurves, mem_kind));
    return finish(c, mem_kind);
}
extern "C" int gsum_variogram_cov(gsum_ctx *c, const int32_t *i1, const int32_t *j1, int64_t nb1, const int32_t *i2, const int32_t *j2,
                                  int64_t nb2, const int32_t *bin_grid, int64_t n, const double *gamma_tilde, int32_t nbins, int32_t ncurves,
                                  const double *tab, double var_factor, double corr_factor, int32_t same_is_one, double *out, int32_t mem_kind) {
    GSUM_RANGE("gsum_variogram_cov");
    if (!c || !i1 || !j1 || !i2 || !j2 || !bin_grid || !gamma_tilde || !tab || !out || nb1 <= 0 || nb2 <= 0 || n < 2 || nbins <= 0 ||
        ncurves <= 0 || ncurves > VG_MAXC)
        return gsum_fail(c, -1, "gsum_variogram_cov: bad argument (1 <= ncurves <= %d, non-empty bins)", VG_MAXC);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    VarioCovArgs P;
    const void *p;
    GSUM_TRY(dev_in(c, WS_IO0, i1, sizeof(int) * nb1, mem_kind, &p)); P.i1 = (const int *)p;
    GSUM_TRY(dev_in(c, WS_IO1, j1, sizeof(int) * nb1, mem_kind, &p)); P.j1 = (const int *)p;
    GSUM_TRY(dev_in(c, WS_IO2, i2, sizeof(int) * nb2, mem_kind, &p)); P.i2 = (const int *)p;
    GSUM_TRY(dev_in(c, WS_IO3, j2, sizeof(int) * nb2, mem_kind, &p)); P.j2 = (const int *)p;
    GSUM_TRY(dev_in(c, WS_MAT, bin_grid, sizeof(int) * n * n, mem_kind, &p)); P.bin_grid = (const int *)p;
    GSUM_TRY(dev_in(c, WS_MISC0, gamma_tilde, sizeof(double) * nbins * ncurves, mem_kind, &p)); P.gamma_tilde = (const double *)p;
    GSUM_TRY(dev_in(c, WS_MISC1, tab, sizeof(double) * (3 * VG_NT + 2), mem_kind, &p)); P.tab = (const double *)p;
    P.nb1 = nb1; P.nb2 = nb2; P.n = n; P.ncurves = ncurves; P.var_factor = var_factor; P.corr_factor = corr_factor; P.same_is_one = same_is_one ? 1 : 0;
    const int64_t nblocks = (nb1 + 15) / 16;
    void *dpart, *dout;
    GSUM_TRY(gsum_ws(c, WS_MISC2, sizeof(double) * nblocks * ncurves, &dpart));
    GSUM_TRY(dev_out(c, WS_MISC3, out, sizeof(double) * ncurves, mem_kind, &dout));
    variogram_cov_kernel<<<(unsigned)nblocks, 256, 0, c->stream>>>(P, (double *)dpart);
    variogram_cov_reduce_kernel<<<ncurves, 256, 0, c->stream>>>((const double *)dpart, nblocks, ncurves, (double)nb1 * (double)nb2, (double *)dout);
    LAUNCHED(c, 2);
    GSUM_TRY(dev_out_finish(c, out, dout, sizeof(double) * ncurves, mem_kind));
    return finish(c, mem_kind);
}

extern "C" int gsum_quadratic_forms(gsum_ctx *c, const double *A, int64_t n, const double *mean, const double *Y, int64_t n_curves,
                                    double *q, int32_t mem_kind) {
    GSUM_RANGE("gsum_quadratic_forms");
    if (!c || !A || !mean || !Y || !q || n <= 0 || n_curves <= 0) return gsum_fail(c, -1, "gsum_quadratic_forms: bad argument");
    if ((size_t)n * QF_C * sizeof(double) > 200 * 1024) return gsum_fail(c, -1, "gsum_quadratic_forms: n <= %d", 200 * 1024 / (QF_C * 8));
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const void *dA, *dm, *dY;
    GSUM_TRY(dev_in(c, WS_MAT, A, sizeof(double) * n * n, mem_kind, &dA));
    GSUM_TRY(dev_in(c, WS_IO0, mean, sizeof(double) * n, mem_kind, &dm));
    GSUM_TRY(dev_in(c, WS_IO1, Y, sizeof(double) * n * n_curves, mem_kind, &dY));
    void *dq, *dpart;
    GSUM_TRY(dev_out(c, WS_IO2, q, sizeof(double) * n_curves, mem_kind, &dq));
    GSUM_TRY(gsum_ws(c, WS_MISC0, sizeof(double) * n * QF_C, &dpart));
    const size_t smem = sizeof(double) * (size_t)n * QF_C;
    GSUM_CUDA(c, cudaFuncSetAttribute(quadform_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)((n + QF_WARPS - 1) / QF_WARPS);
    if (grid > 2 * c->sm_count) grid = 2 * c->sm_count;
    for (int64_t c0 = 0; c0 < n_curves; c0 += QF_C) {
        quadform_rows_kernel<<<grid, 32 * QF_WARPS, smem, c->stream>>>((const double *)dA, n, (const double *)dm, (const double *)dY, n_curves, c0, (double *)dpart);
        quadform_reduce_kernel<<<QF_C, 256, 0, c->stream>>>((const double *)dpart, n, n_curves, c0, (double *)dq);
        LAUNCHED(c, 2);
    }
    GSUM_TRY(dev_out_finish(c, q, dq, sizeof(double) * n_curves, mem_kind));
    return finish(c, mem_kind);
}

extern "C" int gsum_cholesky_errors(gsum_ctx *c, const double *L, int64_t n, const double *mean, const double *Y,
                                    int64_t n_curves, double *E, double *md2, int32_t mem_kind) {
    GSUM_RANGE("gsum_cholesky_errors");
    if (!c || !L || !Y || n <= 0 || n_curves <= 0 || (!E && !md2)) return gsum_fail(c, -1, "gsum_cholesky_errors: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int fk = factor_kind(mem_kind);
    const int64_t np = gsum_pad64(n), rp = gsum_pad64(n_curves);
    const int T = (int)(np / GSUM_TILE);
    const void *dL, *dmean, *dY;
    GSUM_TRY(dev_in(c, WS_IO0, L, sizeof(double) * n * n, fk, &dL));
    GSUM_TRY(dev_in(c, WS_REF, mean, sizeof(double) * n, mem_kind, &dmean));
    GSUM_TRY(dev_in(c, WS_IO1, Y, sizeof(double) * n * n_curves, mem_kind, &dY));
    void *dF, *dW;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dF));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rp * np, &dW));
    pad_factor(c, (const double *)dL, n, (double *)dF);
    launch_transpose_in(c, (const double *)dY, n, n_curves, (const double *)dmean, nullptr, 0, 1.0, (double *)dW, np, rp);
    GSUM_TRY(solve_run(c, solve_desc((double *)dF, (double *)dW, np, T, rp, n_curves), 1));
    if (E) {
        void *dE;
        GSUM_TRY(dev_out(c, WS_IO2, E, sizeof(double) * n * n_curves, mem_kind, &dE));
        launch_transpose_out(c, (const double *)dW, np, n, n_curves, 0, 1.0, nullptr, (double *)dE);
        GSUM_TRY(dev_out_finish(c, E, dE, sizeof(double) * n * n_curves, mem_kind));
    }
    if (md2) {
        void *dm;
        GSUM_TRY(dev_out(c, WS_LL, md2, sizeof(double) * n_curves, mem_kind, &dm));
        rows_sqnorm_kernel<<<(unsigned)((n_curves + 7) / 8), 256, 0, c->stream>>>((const double *)dW, np, n_curves, n, (double *)dm);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, md2, dm, sizeof(double) * n_curves, mem_kind));
    }
    return finish(c, mem_kind);
}

extern "C" int gsum_pivoted_cholesky(gsum_ctx *c, const double *M, int64_t n, double *Lp, int32_t *piv, int32_t *rank,
                                     double *G_out, int32_t mem_kind) {
    GSUM_RANGE("gsum_pivoted_cholesky");
    if (!c || !M || n <= 0) return gsum_fail(c, -1, "gsum_pivoted_cholesky: bad argument");
    // the panel kernel is a cooperative launch of one CTA per 128 rows (all co-resident: 64.5 KiB + 4 n bytes of shared memory each)
    const size_t smem = sizeof(double) * PSTRF_NB * PSTRF_ROWS + sizeof(int) * n;
    const int ncta = (int)((n + PSTRF_ROWS - 1) / PSTRF_ROWS);
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_TRY(chol_set_attrs(c));
    GSUM_CUDA(c, cudaFuncSetAttribute(pstrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        int per_sm = 0, sms = 0;
        GSUM_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pstrf_panel_kernel, PSTRF_ROWS, smem));
        GSUM_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (smem > 220 * 1024 || ncta > per_sm * sms)
            return gsum_fail(c, -1, "gsum_pivoted_cholesky: n = %lld needs %d co-resident CTAs (limit %d)", (long long)n, ncta, per_sm * sms);
    }
    const int64_t np = gsum_pad64(n);
    const int T = (int)(np / GSUM_TILE);
    const void *dM;
    GSUM_TRY(dev_in(c, WS_IO0, M, sizeof(double) * n * n, mem_kind, &dM));
    void *dAf, *dLb, *dpiv, *dpos, *dst, *dPt, *dslots;
    GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * PSTRF_NB * np, &dPt));
    GSUM_TRY(gsum_ws(c, WS_MISC1, sizeof(int32_t) * n, &dpos));
    GSUM_TRY(gsum_ws(c, WS_MISC2, sizeof(PstrfSlot) * 2 * ncta, &dslots));
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dAf));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * np * np, &dLb));
    GSUM_TRY(gsum_ws(c, WS_INFO, sizeof(int32_t) * n, &dpiv));
    GSUM_TRY(gsum_ws(c, WS_MISC0, sizeof(PstrfState), &dst));
    pad_in_kernel<<<dim3((unsigned)((np + 255) / 256), (unsigned)np, 1), 256, 0, c->stream>>>((const double *)dM, n, (double *)dAf, np, np * np, np);
    GSUM_CUDA(c, cudaMemsetAsync(dLb, 0, sizeof(double) * np * np, c->stream));
    pstrf_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((int32_t *)dpiv, (int32_t *)dpos, (PstrfState *)dst, (int)n);
    LAUNCHED(c, 2);
    // n <= 4096: the panel as ONE thread-block cluster (DSMEM candidate exchange, hardware cluster barrier per column) when the device
    // can co-schedule it; otherwise the cooperative-grid kernel
    int cl = 0;
    size_t smem_c = 0;
    if (n <= (int64_t)PSTRF_CMAX * PSTRF_CROWS && !getenv("GSUM_B200_PSTRF_GRID")) {
        cl = 1;
        while (cl * PSTRF_CROWS < n) cl <<= 1;
        smem_c = sizeof(double) * PSTRF_NB * PSTRF_CROWS + sizeof(int) * n;
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(cl); cfg.blockDim = dim3(PSTRF_CROWS); cfg.dynamicSmemBytes = smem_c; cfg.stream = c->stream; cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaFuncSetAttribute(pstrf_panel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c) != cudaSuccess ||
            (cl > 8 && cudaFuncSetAttribute(pstrf_panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
            cudaOccupancyMaxActiveClusters(&nclusters, pstrf_panel_cluster_kernel, &cfg) != cudaSuccess || nclusters < 1) {
            cl = 0;
            cudaGetLastError();
        }
    }
    for (int k = 0; k < n; k += PSTRF_NB) {
        const double *aAf = (const double *)dAf; double *aLb = (double *)dLb, *aPt = (double *)dPt;
        int64_t ald = np; int an = (int)n, ak = k;
        int32_t *apiv = (int32_t *)dpiv, *apos = (int32_t *)dpos; PstrfState *ast = (PstrfState *)dst; PstrfSlot *asl = (PstrfSlot *)dslots;
        if (cl) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(cl); cfg.blockDim = dim3(PSTRF_CROWS); cfg.dynamicSmemBytes = smem_c; cfg.stream = c->stream; cfg.attrs = at; cfg.numAttrs = 1;
            GSUM_CUDA(c, cudaLaunchKernelEx(&cfg, pstrf_panel_cluster_kernel, aAf, aLb, aPt, ald, an, ak, apiv, apos, ast));
        } else {
            void *args[] = {&aAf, &aLb, &aPt, &ald, &an, &ak, &apiv, &apos, &ast, &asl};
            GSUM_CUDA(c, cudaLaunchCooperativeKernel((const void *)pstrf_panel_kernel, dim3(ncta), dim3(PSTRF_ROWS), args, smem, c->stream));
        }
        LAUNCHED(c, 1);
        if (k + PSTRF_NB < n) {
            // dsyrk: Af -= Lb[:, k:k+64] Lb[:, k:k+64]^T on the lower triangle of the (symmetric, physically indexed) matrix
            SchurArgs S;
            S.W = (const double *)dLb + k; S.ld = np; S.bstride = 0; S.T = 1; S.C = (double *)dAf; S.ldc = np; S.cstride = 0; S.lower_only = 1;
            GSUM_TRY(schur_run(c, S, T, 1));
        }
    }
    PstrfState hst;
    GSUM_CUDA(c, cudaMemcpyAsync(&hst, dst, sizeof(hst), cudaMemcpyDeviceToHost, c->stream));
    if (Lp) {
        void *dLp;
        GSUM_TRY(dev_out(c, WS_IO1, Lp, sizeof(double) * n * n, mem_kind, &dLp));
        pstrf_gather_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, c->stream>>>((const double *)dLb, np, (const int32_t *)dpiv, (int)n, (double *)dLp);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, Lp, dLp, sizeof(double) * n * n, mem_kind));
    }
    if (G_out) {
        void *dG;
        GSUM_TRY(dev_out(c, WS_IO2, G_out, sizeof(double) * n * n, mem_kind, &dG));
        GSUM_CUDA(c, cudaMemcpy2DAsync(dG, sizeof(double) * n, dLb, sizeof(double) * np, sizeof(double) * n, n, cudaMemcpyDeviceToDevice, c->stream));
        GSUM_TRY(dev_out_finish(c, G_out, dG, sizeof(double) * n * n, mem_kind));
    }
    if (piv) GSUM_CUDA(c, cudaMemcpyAsync(piv, dpiv, sizeof(int32_t) * n, mem_kind == GSUM_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (mem_kind == GSUM_MEM_HOST) flush_pending(c);
    if (rank) {
        if (mem_kind == GSUM_MEM_HOST) *rank = hst.rank;
        else GSUM_CUDA(c, cudaMemcpy(rank, &hst.rank, sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    GSUM_CUDA(c, cudaGetLastError());
    if (hst.info != 0) return gsum_fail(c, 1, "M is not positive-semidefinite (pivoted Cholesky stopped at rank %d of %lld)", hst.rank, (long long)n);
    return 0;
}

extern "C" int gsum_pc_errors(gsum_ctx *c, const double *Lp, const int32_t *piv, int64_t n, const double *mean,
                              const double *Y, int64_t n_curves, double *E, int32_t mem_kind) {
    GSUM_RANGE("gsum_pc_errors");
    if (!c || !Lp || !piv || !Y || !E || n <= 0 || n_curves <= 0) return gsum_fail(c, -1, "gsum_pc_errors: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int fk = factor_kind(mem_kind);
    const int64_t np = gsum_pad64(n), rp = gsum_pad64(n_curves);
    const int T = (int)(np / GSUM_TILE);
    const void *dL, *dpiv, *dmean, *dY;
    GSUM_TRY(dev_in(c, WS_IO0, Lp, sizeof(double) * n * n, fk, &dL));
    GSUM_TRY(dev_in(c, WS_INFO, piv, sizeof(int32_t) * n, fk, &dpiv));
    GSUM_TRY(dev_in(c, WS_REF, mean, sizeof(double) * n, mem_kind, &dmean));
    GSUM_TRY(dev_in(c, WS_IO1, Y, sizeof(double) * n * n_curves, mem_kind, &dY));
    void *dF, *dW, *dE;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dF));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rp * np, &dW));
    pad_factor(c, (const double *)dL, n, (double *)dF);
    // solve(G, r) with G = Lp[p_inv]  <=>  Lp e = r[piv]: gather rows in pivot order, forward substitute (gsum/diagnostics.py:103-104)
    launch_transpose_in(c, (const double *)dY, n, n_curves, (const double *)dmean, (const int32_t *)dpiv, 0, 1.0, (double *)dW, np, rp);
    GSUM_TRY(solve_run(c, solve_desc((double *)dF, (double *)dW, np, T, rp, n_curves), 1));
    GSUM_TRY(dev_out(c, WS_IO2, E, sizeof(double) * n * n_curves, mem_kind, &dE));
    launch_transpose_out(c, (const double *)dW, np, n, n_curves, 0, 1.0, nullptr, (double *)dE);
    GSUM_TRY(dev_out_finish(c, E, dE, sizeof(double) * n * n_curves, mem_kind));
    return finish(c, mem_kind);
}

static int coverage_rows(gsum_ctx *c, const double *dYt, int64_t ld, int64_t n_rows, int64_t n, const double *lower, const double *upper,
                         int n_alpha, double *coverage_out, int64_t *count_out, int mem_kind) {
    if (n_alpha < 1 || n_alpha > COVG_MAXA) return gsum_fail(c, -1, "coverage: n_alpha must be in 1..%d", COVG_MAXA);
    const void *dlo, *dup; void *dcov = nullptr, *dcnt = nullptr;
    GSUM_TRY(dev_in(c, WS_IO2, lower, sizeof(double) * n_alpha * n, mem_kind, &dlo));
    GSUM_TRY(dev_in(c, WS_IO3, upper, sizeof(double) * n_alpha * n, mem_kind, &dup));
    if (coverage_out) GSUM_TRY(dev_out(c, WS_LL, coverage_out, sizeof(double) * n_rows * n_alpha, mem_kind, &dcov));
    if (count_out) {
        GSUM_TRY(dev_out(c, WS_COUNTS, count_out, sizeof(int64_t) * n_alpha, mem_kind, &dcnt));
        GSUM_CUDA(c, cudaMemsetAsync(dcnt, 0, sizeof(int64_t) * n_alpha, c->stream));
    }
    void *dnested;
    GSUM_TRY(gsum_ws(c, WS_NESTED, sizeof(int), &dnested));
    GSUM_CUDA(c, cudaMemsetAsync(dnested, 0xff, sizeof(int), c->stream));           // nested until a violation is found
    if (n_alpha > 1) {
        const int64_t total = (int64_t)(n_alpha - 1) * n;
        coverage_nested_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1184), 256, 0, c->stream>>>(
            (const double *)dlo, (const double *)dup, n_alpha, (int)n, (int *)dnested);
        LAUNCHED(c, 1);
    }
    const size_t smem = sizeof(double) * 2 * n_alpha * 32 + sizeof(int) * COVG_ROWS * n_alpha;
    GSUM_CUDA(c, cudaFuncSetAttribute(coverage_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coverage_rows_kernel<<<(unsigned)((n_rows + COVG_ROWS - 1) / COVG_ROWS), COVG_WARPS * 32, smem, c->stream>>>(
        dYt, ld, n_rows, (int)n, (const double *)dlo, (const double *)dup, n_alpha, (double *)dcov, (unsigned long long *)dcnt,
        (const int *)dnested);
    LAUNCHED(c, 1);
    if (coverage_out) GSUM_TRY(dev_out_finish(c, coverage_out, dcov, sizeof(double) * n_rows * n_alpha, mem_kind));
    if (count_out) GSUM_TRY(dev_out_finish(c, count_out, dcnt, sizeof(int64_t) * n_alpha, mem_kind));
    return 0;
}

extern "C" int gsum_draws(gsum_ctx *c, const double *L, int64_t n, const double *mean, const double *Z, int64_t n_draws,
                          uint64_t seed, int64_t first_draw, const double *draw_scale, double *draws_out, const double *lower,
                          const double *upper, int32_t n_alpha, double *coverage_out, int64_t *count_out, int32_t mem_kind) {
    GSUM_RANGE("gsum_draws");
    if (!c || !L || n <= 0 || n_draws <= 0 || first_draw < 0) return gsum_fail(c, -1, "gsum_draws: bad argument");
    if ((coverage_out || count_out) && (!lower || !upper)) return gsum_fail(c, -1, "gsum_draws: coverage needs lower and upper");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    GSUM_TRY(chol_set_attrs(c));
    GSUM_CUDA(c, cudaFuncSetAttribute(draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
    const int fk = factor_kind(mem_kind);
    const int64_t np = gsum_pad64(n), rp = gsum_pad64(n_draws);
    const int T = (int)(np / GSUM_TILE);
    const void *dL, *dmean, *dZ;
    GSUM_TRY(dev_in(c, WS_IO0, L, sizeof(double) * n * n, fk, &dL));
    GSUM_TRY(dev_in(c, WS_REF, mean, sizeof(double) * n, mem_kind, &dmean));
    GSUM_TRY(dev_in(c, WS_IO1, Z, sizeof(double) * n * n_draws, mem_kind, &dZ));
    void *dF, *dZn, *dYt;
    GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * np * np, &dF));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rp * np, &dZn));
    GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * rp * np, &dYt));
    // factor with explicit zeros above the diagonal (the TRMM reads whole tiles)
    pad_lower_kernel<<<dim3((unsigned)((np + 255) / 256), (unsigned)np), 256, 0, c->stream>>>((const double *)dL, n, (double *)dF, np);
    LAUNCHED(c, 1);
    const void *dsc = nullptr;
    if (draw_scale) GSUM_TRY(dev_in(c, WS_SCALE, draw_scale, sizeof(double) * n_draws, mem_kind, &dsc));
    if (Z) {
        launch_transpose_in(c, (const double *)dZ, n, n_draws, nullptr, nullptr, 0, -1.0, (double *)dZn, np, rp);
        if (dsc) {
            scale_rows_kernel<<<dim3((unsigned)((np + 255) / 256), (unsigned)n_draws), 256, 0, c->stream>>>((double *)dZn, np, n_draws, (const double *)dsc);
            LAUNCHED(c, 1);
        }
    } else {
        normal_rows_kernel<<<dim3((unsigned)rp, (unsigned)((np / 2 + 255) / 256)), 256, 0, c->stream>>>((double *)dZn, np, rp, n_draws, (int)n, seed,
                                                                                                      first_draw, (const double *)dsc);
        LAUNCHED(c, 1);
    }
    DrawArgs D;
    D.Zn = (const double *)dZn; D.L = (const double *)dF; D.mean = (const double *)dmean; D.Yt = (double *)dYt; D.ld = np; D.n = (int)n;
    draws_kernel<<<dim3((unsigned)T, (unsigned)(rp / GSUM_TILE)), CHOL_THREADS, CHOL_SMEM_BYTES, c->stream>>>(D);
    LAUNCHED(c, 1);
    if (draws_out) {
        void *dD;
        GSUM_TRY(dev_out(c, WS_MISC2, draws_out, sizeof(double) * n * n_draws, mem_kind, &dD));
        launch_transpose_out(c, (const double *)dYt, np, n, n_draws, 0, 1.0, nullptr, (double *)dD);
        GSUM_TRY(dev_out_finish(c, draws_out, dD, sizeof(double) * n * n_draws, mem_kind));
    }
    if (coverage_out || count_out)
        GSUM_TRY(coverage_rows(c, (const double *)dYt, np, n_draws, n, lower, upper, n_alpha, coverage_out, count_out, mem_kind));
    return finish(c, mem_kind);
}

extern "C" int gsum_credible_interval(gsum_ctx *c, const double *Y, int64_t n, int64_t n_curves, const double *lower,
                                      const double *upper, int32_t n_alpha, double *coverage_out, int32_t mem_kind) {
    GSUM_RANGE("gsum_credible_interval");
    if (!c || !Y || !lower || !upper || !coverage_out || n <= 0 || n_curves <= 0) return gsum_fail(c, -1, "gsum_credible_interval: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int64_t np = gsum_pad64(n), rp = gsum_pad64(n_curves);
    const void *dY; void *dYt;
    GSUM_TRY(dev_in(c, WS_IO1, Y, sizeof(double) * n * n_curves, mem_kind, &dY));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * rp * np, &dYt));
    launch_transpose_in(c, (const double *)dY, n, n_curves, nullptr, nullptr, 0, 1.0, (double *)dYt, np, rp);
    GSUM_TRY(coverage_rows(c, (const double *)dYt, np, n_curves, n, lower, upper, n_alpha, coverage_out, nullptr, mem_kind));
    return finish(c, mem_kind);
}

// ---- decomposition='eig' route (eig.cuh) -------------------------------------------------------------------------
#include "eig.cuh"
#include <vector>
#include <algorithm>
#include <numeric>

static void launch_eig_gemm(gsum_ctx *c, const EigGemmArgs &g) {
    dim3 grid((unsigned)((g.N + 63) / 64), (unsigned)((g.M + 63) / 64));
    cudaFuncSetAttribute(eig_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EG_SMEM_BYTES);   // 72.7 KB: two stages
    eig_gemm_kernel<<<grid, 128, EG_SMEM_BYTES, c->stream>>>(g);
    LAUNCHED(c, 1);
}

extern "C" int gsum_eigh(gsum_ctx *c, const double *A, int64_t n, double *w, double *V, int32_t *sweeps_out, int32_t mem_kind) {
    GSUM_RANGE("gsum_eigh");
    if (!c || !A || !w || n <= 0 || n > (1 << 20)) return gsum_fail(c, -1, "gsum_eigh: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    mem_kind &= 1;
    const int ni = (int)n, np = ni + (ni & 1);
    const int64_t ld = n;
    const void *dA;
    void *dG, *dVt, *dw, *dflip, *dperm, *dcnt;
    GSUM_TRY(dev_in(c, WS_IO0, A, sizeof(double) * n * n, mem_kind, &dA));
    std::vector<double> hw(n), hs(n), hrq(n);
    const double eps = 2.220446049250313e-16;
    const double tol = sqrt((double)n) * eps;
    double tol_abs = 0.0;

    // factor mode (eig.cuh): pivoted Cholesky A = F F^T first; full rank -> iterate on F, no V
    bool factor_mode = false;
    double tol_gamma = 0.0;
    void *dGp;
    GSUM_TRY(gsum_ws(c, WS_IO3, sizeof(double) * n * n, &dGp));
    if (getenv("GSUM_B200_EIGH_FACTOR")) {                  // opt-in: faster, looser (see eig.cuh "factor mode")
        const int rc = gsum_pivoted_cholesky(c, (const double *)dA, n, nullptr, nullptr, nullptr, (double *)dGp, GSUM_MEM_DEVICE);
        if (rc <= -100) return rc;
        factor_mode = (rc == 0);
        if (!factor_mode) c->err[0] = 0;
    }
    GSUM_TRY(gsum_ws(c, WS_IO1, sizeof(double) * n * ld, &dG));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * n * ld, &dVt));
    GSUM_TRY(gsum_ws(c, WS_LL, sizeof(double) * n, &dw));
    GSUM_TRY(gsum_ws(c, WS_SCALE, sizeof(double) * 2 * n, &dflip));
    GSUM_TRY(gsum_ws(c, WS_ORD, sizeof(int32_t) * n, &dperm));
    GSUM_TRY(gsum_ws(c, WS_COUNTS, sizeof(unsigned int) * (JAC_MAX_SWEEPS + 1), &dcnt));
    GSUM_CUDA(c, cudaMemsetAsync(dcnt, 0, sizeof(unsigned int) * (JAC_MAX_SWEEPS + 1), c->stream));
    dim3 gt((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
    // G <- A, V <- I; |A|_F scales the optional absolute part of the rotation criterion (eig.cuh; GSUM_B200_EIGH_ABS = c
    // skips pairs with |gamma| <= c eps |A|_F min(|g_p|, |g_q|); default c = 0.1, c = 0: relative criterion only)
    jacobi_init_kernel<<<(unsigned)n, 256, 0, c->stream>>>((const double *)dA, (double *)dG, (double *)dVt, ni, ld);
    rows_sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>((const double *)dG, ld, n, n, (double *)dw);
    LAUNCHED(c, 2);
    GSUM_CUDA(c, cudaMemcpyAsync(hw.data(), dw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    double fro2 = 0.0;
    for (int64_t i = 0; i < n; i++) fro2 += hw[i];
    tol_abs = 0.1 * eps * sqrt(fro2);
    if (const char *e = getenv("GSUM_B200_EIGH_ABS")) tol_abs = atof(e) * eps * sqrt(fro2);
    if (factor_mode) {
        tol_gamma = getenv("GSUM_B200_EIGH_ABS") ? tol_abs : 0.01 * eps * sqrt(fro2);      // c = 0.01: best measured trade (eig.cuh)
        if (getenv("GSUM_B200_EIGH_STRICT")) tol_gamma = 0.0;                               // relative criterion only
        tol_abs = 0.0;
        jacobi_init_factor_kernel<<<gt, 256, 0, c->stream>>>((const double *)dGp, n, (double *)dG, ni, ld);
        LAUNCHED(c, 1);
    }
    double *dVt_it = factor_mode ? nullptr : (double *)dVt;
    // One sweep = n - 1 dependent launches of a few microseconds each: launch-bound, so the sweep is captured once as a
    // CUDA graph (counter reset + the rounds) and replayed until a sweep makes no rotation.
    const bool trace = getenv("GSUM_B200_EIGH_TRACE") != nullptr;      // rotations per sweep on stderr
    int sweeps = 0;
    bool converged = (n == 1);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    // register-resident round for n <= 1024 (G and V^T sit in L2 and the round is latency-bound: 99 -> 89 ms at
    // N = 1024); beyond, its unconditional fetch of the V^T rows costs more than the second pass it saves
    // (measured: N = 2048 385 -> 1548 ms), so the two-pass kernel stays
    void (*round_fn)(double *, double *, int, int64_t, int, int, double, double, double, unsigned int *, const int32_t *) = jacobi_round_kernel;
    if (!getenv("GSUM_B200_EIGH_TWOPASS")) {
        if (ni <= 1 * JAC_THREADS) round_fn = jacobi_round_reg_kernel<1>;
        else if (ni <= 2 * JAC_THREADS) round_fn = jacobi_round_reg_kernel<2>;
        else if (ni <= 4 * JAC_THREADS) round_fn = jacobi_round_reg_kernel<4>;
    }
    // GSUM_B200_EIGH_ORDER=modulus: modulus ordering on positions sorted by decreasing row norm (eig.cuh, jacobi_select);
    // the ranking is refreshed on the host before every sweep (one row-norm kernel, n doubles down, n ints up)
    // Default since round 2: the whole GPU suite is green under it and it needs 12-13 instead of 19-20 sweeps at the same residual
    // (N = 1024 / 2048: 62 / 224 ms against 87 / 384 ms; LAPACK on the host 65 / 284 ms; profiles/r02_eig_probe.txt).
    // GSUM_B200_EIGH_ORDER=roundrobin selects the tournament order of round 1.
    const bool modulus = !(getenv("GSUM_B200_EIGH_ORDER") && !strcmp(getenv("GSUM_B200_EIGH_ORDER"), "roundrobin"));
    void *dorder = nullptr;
    std::vector<int32_t> horder(n);
    if (modulus) GSUM_TRY(gsum_ws(c, WS_Q, sizeof(int32_t) * n, &dorder));
    const int rounds_sweep = modulus ? ni : np - 1, ctas = modulus ? ni : np / 2;
    auto enqueue_sweep = [&]() {
        cudaMemsetAsync(dcnt, 0, sizeof(unsigned int), c->stream);
        for (int r = 0; r < rounds_sweep; r++)
            round_fn<<<ctas, JAC_THREADS, 0, c->stream>>>((double *)dG, dVt_it, ni, ld, np, r, tol, tol_abs, tol_gamma,
                                                         (unsigned int *)dcnt, (const int32_t *)dorder);
    };
    auto refresh_order = [&]() -> int {
        rows_sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>((const double *)dG, ld, n, n, (double *)dw);
        LAUNCHED(c, 1);
        GSUM_CUDA(c, cudaMemcpyAsync(hw.data(), dw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
        std::iota(horder.begin(), horder.end(), 0);
        std::stable_sort(horder.begin(), horder.end(), [&](int32_t a, int32_t b) { return hw[a] > hw[b]; });
        GSUM_CUDA(c, cudaMemcpyAsync(dorder, horder.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        return 0;
    };
    // a caller-provided stream may not be capturable (legacy default stream): plain launches then
    if (!converged && !getenv("GSUM_B200_EIGH_NOGRAPH") &&
        cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueue_sweep();
        if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            graph = nullptr; gexec = nullptr;
        }
    }
    cudaGetLastError();
    int rc_loop = 0;
    while (!converged && sweeps < JAC_MAX_SWEEPS) {
        unsigned int rot = 0;
        if (modulus) GSUM_TRY(refresh_order());
        if (gexec) { if (cudaGraphLaunch(gexec, c->stream) != cudaSuccess) { rc_loop = -100; break; } }
        else enqueue_sweep();
        if (cudaMemcpyAsync(&rot, dcnt, sizeof(rot), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { rc_loop = -100; break; }
        LAUNCHED(c, rounds_sweep);
        if (trace) fprintf(stderr, "[gsum_eigh] n=%d sweep %d: %u rotations\n", ni, sweeps, rot);
        sweeps++;
        converged = (rot == 0);
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    if (rc_loop) return gsum_fail(c, -100, "gsum_eigh: CUDA error in the sweep graph (%s)", cudaGetErrorString(cudaGetLastError()));
    if (sweeps_out) *sweeps_out = sweeps;
    if (factor_mode) {
        jacobi_finish_factor_kernel<<<(unsigned)n, JAC_THREADS, 0, c->stream>>>((const double *)dG, ni, ld, (double *)dw, (double *)dflip);
        dVt = dG;                                          // the gather below reads the eigenvectors from here
    } else {
        jacobi_finish_kernel<<<(unsigned)n, JAC_THREADS, 0, c->stream>>>((const double *)dG, (const double *)dVt, ni, ld, (double *)dw, (double *)dflip,
                                                                         (double *)dflip + n);
        GSUM_CUDA(c, cudaMemcpyAsync(hrq.data(), (double *)dflip + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    LAUNCHED(c, 1);
    std::vector<int32_t> perm(n);
    GSUM_CUDA(c, cudaMemcpyAsync(hw.data(), dw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    GSUM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < n; i++)
        if (!std::isfinite(hw[i])) return gsum_fail(c, 1, "gsum_eigh: non-finite eigenvalue (NaN or Inf in the input matrix)");
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return hw[a] < hw[b]; });
    for (int64_t i = 0; i < n; i++) hs[i] = hw[perm[i]];
    // v_j is an eigenvector iff its Rayleigh quotient reproduces |g_j| (see eig.cuh: +/- lambda pairs of an indefinite matrix)
    bool mixed = false;
    const double wmax = std::max(fabs(hs[0]), fabs(hs[n - 1]));
    for (int64_t i = 0; i < n && !factor_mode; i++)
        if (fabs(hw[i]) - fabs(hrq[i]) > 1e-6 * fabs(hw[i]) + 1e-10 * wmax) mixed = true;
    if (mem_kind == GSUM_MEM_DEVICE) GSUM_CUDA(c, cudaMemcpy(w, hs.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    else memcpy(w, hs.data(), sizeof(double) * n);
    if (V) {
        GSUM_CUDA(c, cudaMemcpy(dperm, perm.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
        void *dV;
        GSUM_TRY(dev_out(c, WS_IO2, V, sizeof(double) * n * n, mem_kind, &dV));
        if (factor_mode && n > 1) {
            // eigenvectors = normalised rows of G, orthogonal to ~1e-10 (eig.cuh): one Newton-Schulz step V <- V (3 I - V^T V) / 2
            void *dV0, *dS;
            GSUM_TRY(gsum_ws(c, WS_MAT, sizeof(double) * n * n, &dV0));
            GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * n * n, &dS));
            jacobi_gather_kernel<<<gt, 256, 0, c->stream>>>((const double *)dVt, ld, (const int32_t *)dperm, (const double *)dflip, ni, (double *)dV0);
            EigGemmArgs g1{};
            g1.A = (const double *)dV0; g1.lda = n; g1.transA = 1; g1.B = (const double *)dV0; g1.ldb = n;
            g1.C = (double *)dS; g1.ldc = n; g1.M = n; g1.N = n; g1.K = n;
            launch_eig_gemm(c, g1);
            newton_schulz_T_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, c->stream>>>((double *)dS, n);
            EigGemmArgs g2{};
            g2.A = (const double *)dV0; g2.lda = n; g2.transA = 0; g2.B = (const double *)dS; g2.ldb = n;
            g2.C = (double *)dV; g2.ldc = n; g2.M = n; g2.N = n; g2.K = n;
            launch_eig_gemm(c, g2);
            LAUNCHED(c, 2);
        } else {
            jacobi_gather_kernel<<<gt, 256, 0, c->stream>>>((const double *)dVt, ld, (const int32_t *)dperm,
                                                           (const double *)dflip, ni, (double *)dV);
            LAUNCHED(c, 1);
        }
        GSUM_TRY(dev_out_finish(c, V, dV, sizeof(double) * n * n, mem_kind));
    }
    GSUM_TRY(finish(c, mem_kind));
    if (!converged) {
        gsum_fail(c, 1, "gsum_eigh: Jacobi iteration did not converge in %d sweeps", JAC_MAX_SWEEPS);
        return 1;
    }
    if (mixed) {
        gsum_fail(c, 2, "gsum_eigh: indefinite matrix with eigenvalues of equal magnitude and opposite sign (not separable by one-sided Jacobi)");
        return 2;
    }
    return 0;
}


extern "C" int gsum_eig_solve(gsum_ctx *c, const double *w, const double *V, int64_t n, const double *Y, int64_t nrhs,
                              const double *mean, double *X, int32_t mode, int32_t mem_kind) {
    GSUM_RANGE("gsum_eig_solve");
    if (!c || !w || !V || !Y || !X || n <= 0 || nrhs <= 0 || (mode != 0 && mode != 1)) return gsum_fail(c, -1, "gsum_eig_solve: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int fk = factor_kind(mem_kind);
    const void *dw, *dV, *dY, *dmean;
    void *dT, *dX;
    GSUM_TRY(dev_in(c, WS_LOGDET, w, sizeof(double) * n, fk, &dw));
    GSUM_TRY(dev_in(c, WS_IO0, V, sizeof(double) * n * n, fk, &dV));
    GSUM_TRY(dev_in(c, WS_IO1, Y, sizeof(double) * n * nrhs, mem_kind, &dY));
    GSUM_TRY(dev_in(c, WS_REF, mean, sizeof(double) * n, mem_kind, &dmean));
    GSUM_TRY(dev_out(c, WS_IO2, X, sizeof(double) * n * nrhs, mem_kind, &dX));
    EigGemmArgs g1{};
    g1.A = (const double *)dV; g1.lda = n; g1.transA = 1;                  // T = diag(f(w)) V^T (Y - mean)
    g1.B = (const double *)dY; g1.ldb = nrhs; g1.bsub = (const double *)dmean;
    g1.M = n; g1.N = nrhs; g1.K = n;
    g1.rs = (const double *)dw; g1.row_mode = mode == 0 ? 1 : 2;
    if (mode == 1) {
        g1.C = (double *)dX; g1.ldc = nrhs;
        launch_eig_gemm(c, g1);
    } else {
        GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * n * nrhs, &dT));
        g1.C = (double *)dT; g1.ldc = nrhs;
        launch_eig_gemm(c, g1);
        EigGemmArgs g2{};
        g2.A = (const double *)dV; g2.lda = n; g2.transA = 0;              // X = V T
        g2.B = (const double *)dT; g2.ldb = nrhs; g2.bsub = nullptr;
        g2.C = (double *)dX; g2.ldc = nrhs;
        g2.M = n; g2.N = nrhs; g2.K = n;
        g2.rs = nullptr; g2.row_mode = 0;
        launch_eig_gemm(c, g2);
    }
    GSUM_TRY(dev_out_finish(c, X, dX, sizeof(double) * n * nrhs, mem_kind));
    return finish(c, mem_kind);
}

extern "C" int gsum_eig_conditional(gsum_ctx *c, const double *w, const double *V, int64_t n, const double *R_on, int64_t m,
                                    const double *D, int64_t k, double *lin_out, double *var_out, double *cov_out, int32_t mem_kind) {
    GSUM_RANGE("gsum_eig_conditional");
    if (!c || !w || !V || !R_on || n <= 0 || m <= 0 || (lin_out && (!D || k <= 0)) || (!lin_out && !var_out && !cov_out))
        return gsum_fail(c, -1, "gsum_eig_conditional: bad argument");
    GSUM_CUDA(c, cudaSetDevice(c->device));
    const int fk = factor_kind(mem_kind);
    const void *dw, *dV, *dR, *dD = nullptr;
    void *dU;
    GSUM_TRY(dev_in(c, WS_LOGDET, w, sizeof(double) * n, fk, &dw));
    GSUM_TRY(dev_in(c, WS_IO0, V, sizeof(double) * n * n, fk, &dV));
    GSUM_TRY(dev_in(c, WS_IO1, R_on, sizeof(double) * n * m, mem_kind, &dR));
    GSUM_TRY(gsum_ws(c, WS_RHS, sizeof(double) * n * m, &dU));
    EigGemmArgs g{};
    g.A = (const double *)dV; g.lda = n; g.transA = 1;                     // U = V^T R_on
    g.B = (const double *)dR; g.ldb = m; g.C = (double *)dU; g.ldc = m;
    g.M = n; g.N = m; g.K = n;
    launch_eig_gemm(c, g);
    if (lin_out) {
        void *dUD, *dlin;
        GSUM_TRY(dev_in(c, WS_IO3, D, sizeof(double) * n * k, mem_kind, &dD));
        GSUM_TRY(gsum_ws(c, WS_GRAM, sizeof(double) * n * k, &dUD));
        GSUM_TRY(dev_out(c, WS_LL, lin_out, sizeof(double) * m * k, mem_kind, &dlin));
        EigGemmArgs h = g;                                                 // UD = V^T D
        h.B = (const double *)dD; h.ldb = k; h.C = (double *)dUD; h.ldc = k; h.N = k;
        launch_eig_gemm(c, h);
        EigGemmArgs l{};                                                   // lin = U^T diag(1/w) UD
        l.A = (const double *)dU; l.lda = m; l.transA = 1;
        l.B = (const double *)dUD; l.ldb = k; l.bdiv = (const double *)dw;
        l.C = (double *)dlin; l.ldc = k; l.M = m; l.N = k; l.K = n;
        launch_eig_gemm(c, l);
        GSUM_TRY(dev_out_finish(c, lin_out, dlin, sizeof(double) * m * k, mem_kind));
    }
    if (var_out) {
        void *dvar;
        GSUM_TRY(dev_out(c, WS_MISC0, var_out, sizeof(double) * m, mem_kind, &dvar));
        eig_colquad_kernel<<<(unsigned)((m + 31) / 32), 1024, 0, c->stream>>>((const double *)dU, n, m, (const double *)dw, (double *)dvar);
        LAUNCHED(c, 1);
        GSUM_TRY(dev_out_finish(c, var_out, dvar, sizeof(double) * m, mem_kind));
    }
    if (cov_out) {
        void *dcov;
        GSUM_TRY(dev_out(c, WS_IO2, cov_out, sizeof(double) * m * m, mem_kind, &dcov));
        EigGemmArgs q{};                                                   // cov = U^T diag(1/w) U
        q.A = (const double *)dU; q.lda = m; q.transA = 1;
        q.B = (const double *)dU; q.ldb = m; q.bdiv = (const double *)dw;
        q.C = (double *)dcov; q.ldc = m; q.M = m; q.N = m; q.K = n;
        launch_eig_gemm(c, q);
        GSUM_TRY(dev_out_finish(c, cov_out, dcov, sizeof(double) * m * m, mem_kind));
    }
    return finish(c, mem_kind);
}
