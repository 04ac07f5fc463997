// The factorisation kernel: the heterogeneous schedule of hetero.cuh with the operand rings of the GEMM CTAs filled by the
// TMA unit.
//
// Why TMA: beside DMMA-streaming warps a producer warp gets few issue slots (measured in round 1: one cp.async warp per
// group left the math warps waiting 16 % of the time; two per group is what fits in 384 threads and caps the CTA at two
// math groups — that variant ran the C4 launch in 2.04 ms against 1.91 ms here).
// A tiled tensor-map copy (cp.async.bulk.tensor.2d, SASS UTMALDG) moves a 64 x 16 FP64 box (8 KiB) per instruction, so
// ONE elected lane per group feeds its ring and the CTA can run THREE math groups = three DMMA warps per sub-partition.
//
// Shared-memory layout: a stage is four 8 KiB boxes of 64 rows x 128 B written with CU_TENSOR_MAP_SWIZZLE_128B (16-byte
// chunk index XOR row % 8; dense rows, no padding).  An operand stage (K depth 32) holds  A[k 0..15] A[k 16..31]
// B[k 0..15] B[k 16..31];  a tile stage (the C tile, or M_kk) holds the four 16-column slices of the 64 x 64 tile.
// The DMMA fragment loads stay bank-conflict free by PERMUTING THE CONTRACTION INDEX (the same permutation for A and B,
// so the product is unchanged): k-step a of a box contracts k' in {2a, 2a+1, 2a+8, 2a+9}; lane (g, t) reads chunk
// a ^ 4(t >> 1) of row g — the four rows of a half-warp then cover all eight chunks, i.e. all 32 banks exactly once.
#pragma once
#include <cuda.h>
#include "hetero.cuh"
#include "chain.cuh"

#define HX_NG 3                         // math groups per GEMM CTA (4 warps each, one per sub-partition)
#define HX_NST 2                        // ring stages per group
#define HX_THREADS 512                  // warps 0-11 math (3 groups), 12-14 producers (one elected lane each), 15 idle
#define HX_BOX_DOUBLES 1024             // 64 rows x 16 doubles
#define HX_STAGE_DOUBLES (4 * HX_BOX_DOUBLES)
#define HX_THINX_DOUBLES 1024            // per group: two 8 x 64 exchange buffers of the thin tasks (see hx_stage_mma_thin)
#define HX_SMEM_BYTES ((HX_NG * HX_NST * HX_STAGE_DOUBLES + HX_NG * HX_THINX_DOUBLES) * 8 + 1024)      // + alignment slack (swizzle atoms: 1024 B)
#define HX_MATH_REGS 160                // 384 * 160 + 128 * 32 = 65536

struct HeteroMaps {
    CUtensorMap A;      // factor rows:  (batch * bstride / ld) x ld, box 16 x 64
    CUtensorMap W;      // border rows:  (batch * wstride / ld) x ld, box 16 x 64
    CUtensorMap W8;     // border rows, box 16 x 8 (thin tasks)
    CUtensorMap M;      // (batch * T * 64) x 64, box 16 x 64
};

__device__ __forceinline__ void hx_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hx_ring_advance(RingState &r) {
    if (++r.stage == HX_NST) { r.stage = 0; r.phase ^= 1u; }
}

// One ring stage on a warp's two m tiles (rows row0 + g, row1 + g; 8-aligned) against n tiles 0..7 of the B boxes.
// oa[a] = 2 (a ^ g ^ 4(t >> 1)) + (t & 1): this lane's offset inside a 128-byte row for k-step a.
__device__ __forceinline__ void hx_stage_mma(Acc &acc, const double *As, const double *Bs, int row0, int row1, int g, const int (&oa)[4]) {
    const double *a0p = As + (row0 + g) * 16, *a1p = As + (row1 + g) * 16, *bp = Bs + g * 16;
#pragma unroll
    for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int off = kb * HX_BOX_DOUBLES + oa[a];
            double b[8];
            const double a0 = a0p[off], a1 = a1p[off];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) b[nt] = bp[nt * 128 + off];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                dmma884(acc[0][nt][0], acc[0][nt][1], a0, b[nt]);
                dmma884(acc[1][nt][0], acc[1][nt][1], a1, b[nt]);
            }
        }
}

// The same stage on a DIAGONAL task (SYRK half: only the 36 blocks on or below the diagonal).  Warp wg owns the block rows
// wg (N0 = wg + 1 blocks) and 7 - wg (9 - N0 blocks): nine blocks per warp, the same load on every sub-partition.  N0 is a
// template parameter so that the code is straight-line: a DMMA under a run-time predicate costs a WARPSYNC.ALL + NOP pair
// each (SASS), and beside two groups streaming unconditional DMMAs such a warp falls far behind its share of the pipe
// (measured: 15-19k cycles per K = 64 step for 9/16 of the work of a full step, which takes 11.5k; profiles/r02_notes.md).
template <int N0>
__device__ __forceinline__ void hx_stage_mma_diag(Acc &acc, const double *As, int g, const int (&oa)[4]) {
    constexpr int N1 = 9 - N0, ROW0 = (N0 - 1) * 8, ROW1 = (8 - N0) * 8;
    const double *a0p = As + (ROW0 + g) * 16, *a1p = As + (ROW1 + g) * 16, *bp = As + g * 16;
#pragma unroll
    for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int off = kb * HX_BOX_DOUBLES + oa[a];
            double b[N1];
            const double a0 = a0p[off], a1 = a1p[off];
#pragma unroll
            for (int nt = 0; nt < N1; nt++) b[nt] = bp[nt * 128 + off];
#pragma unroll
            for (int nt = 0; nt < N1; nt++) {
                if (nt < N0) dmma884(acc[0][nt][0], acc[0][nt][1], a0, b[nt]);
                dmma884(acc[1][nt][0], acc[1][nt][1], a1, b[nt]);
            }
        }
}

// The same stage on a THIN task (a border tile row with <= 8 right-hand sides: an 8 x 64 row block).  The 64 columns are
// dealt over the group's four warps — n tiles 2 wg and 2 wg + 1, accumulated in acc[0][0] and acc[0][1] — so the task costs
// a quarter of a full-height DMMA stream on every sub-partition instead of one warp's worth on sub-partition 0 alone (which
// shares that sub-partition with two streaming groups: measured 7-10k cycles per K = 64 step for 1/8 of the work of a
// full step).  After the main loop the row block is gathered in warp 0 through shared memory for the warp-local
// triangular solve; every element is still accumulated by one warp over k in the same order, so the result is unchanged.
__device__ __forceinline__ void hx_stage_mma_thin(Acc &acc, const double *As, const double *Bs, int wg, int g, const int (&oa)[4]) {
    const double *ap = As + g * 16, *bp = Bs + (16 * wg + g) * 16;
#pragma unroll
    for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int off = kb * HX_BOX_DOUBLES + oa[a];
            const double a0 = ap[off], b0 = bp[off], b1 = bp[128 + off];
            dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        }
}

// X = S * L_kk^{-T} on this warp's MT row blocks of 8 x 64, held as C fragments  T[mt][nt][e] <-> row 8 mt + g, column
// 8 nt + 2t + e, with M_kk in the swizzled tile layout.  Right-looking over 8-column blocks:
//   X_cb = S_cb * Dinv_cb^T  (two DMMAs),   S_j -= X_cb * L[j, cb]^T  for the later blocks j (two DMMAs each, independent).
// The C -> A fragment re-layouts are quad shuffles; the MT row blocks are independent chains that interleave.  om[p][h] = this lane's offset inside a
// 128-byte row for column  8 (2 p' + p) + 4 h + t  (p = cb & 1 selects the half of the 16-column box).
template <int MT>
__device__ __forceinline__ void hx_trsm_dinv(Acc &T, const double *Ms, int g, int t, const int (&om)[2][2]) {
    const unsigned FULLMASK = 0xffffffffu;
    const int s0 = t >> 1, s1 = 2 + (t >> 1);
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int cb = 0; cb < 8; cb++) {
        const double *box = Ms + (cb >> 1) * HX_BOX_DOUBLES;
        const double b0 = box[(cb * 8 + g) * 16 + om[cb & 1][0]], b1 = box[(cb * 8 + g) * 16 + om[cb & 1][1]];
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
            const double l0 = box[(j * 8 + g) * 16 + om[cb & 1][0]], l1 = box[(j * 8 + g) * 16 + om[cb & 1][1]];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                dmma884(T[mt][j][0], T[mt][j][1], a0[mt], l0);
                dmma884(T[mt][j][0], T[mt][j][1], a1[mt], l1);
            }
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(HX_THREADS, 1) chol_hetero_tma_kernel(HeteroArgs D, const __grid_constant__ HeteroMaps maps) {
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[HX_NG][HX_NST], empty_bar[HX_NG][HX_NST], tq_full[HX_NG][HT_QD], tq_empty[HX_NG][HT_QD];
    __shared__ int4 tq[HX_NG][HT_QD];
    __shared__ int done_cnt[HX_NG][HT_QD];
    __shared__ int helper_done;
    __shared__ int thin_cnt[HX_NG];
    const BorderedBatch &P = D.P;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool st_on = STATS && D.stats != nullptr;
    long long st[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long st_t0 = STATS ? clock64() : 0;
#define HT_T0() const long long _t = st_on ? clock64() : 0
#define HT_ACC(q) do { if (st_on) st[q] += clock64() - _t; } while (0)
#define HT_TRACE(e) do { if (STATS && D.trace && (int)blockIdx.x == D.trace_cta && (w & 3) == 0 && lane == 0 && n < 64) D.trace[(q * 64 + n) * 8 + (e)] = clock64(); } while (0)
    // 1024-byte aligned base (128B-swizzle atoms)
    // (index arithmetic on the __shared__ array, not integer casts: the compiler must keep seeing shared-space pointers,
    // or every fragment load turns into a generic LD.E)
    double *smem = smem_raw + (((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) >> 3);

    if ((int)blockIdx.x < D.nfactor_ctas) {
        // ============================ factor CTA ================================================================
        // fewer than four workers (or chain mode): the idle warpgroup parks at 32 registers, as the producers of a GEMM CTA do
        // (four factor workers: every warpgroup keeps the 128 registers of the launch)
        if (D.chain || D.nworkers < 4) {
            if (w >= 12) { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); return; }
            asm volatile("setmaxnreg.inc.sync.aligned.u32 " HT_STR(HX_MATH_REGS) ";");
        }
        if (D.chain) {
            if (tid >= CH_THREADS + 32) return;
            if (tid >= CH_THREADS) { ht_chain_publisher(D, smem, (int)blockIdx.x); return; }
            if (tid >= 128) { ht_chain_helper(D, smem, (int)blockIdx.x); return; }
            ht_chain_worker(D, smem, (int)blockIdx.x, st_on ? st : nullptr);
            if (st_on && tid == 0) {
                long long *o = D.stats + (int64_t)blockIdx.x * HT_NSTAT;
                o[0] = clock64() - st_t0; o[1] = st[0]; o[2] = st[1]; o[3] = st[2]; o[4] = st[3]; o[5] = st[4]; o[6] = st[5]; o[7] = st[6]; o[8] = st[7]; o[9] = st[8];
            }
            return;
        }
        if (tid >= 128 * D.nworkers) return;
        ht_factor_worker(D, smem + (tid >> 7) * HT_WORKER_DOUBLES, st_on ? st : nullptr);
        if (st_on && (tid & 127) == 0) {
            long long *o = D.stats + (int64_t)blockIdx.x * HT_NSTAT + (tid >> 7) * 6;
            o[0] = clock64() - st_t0; o[1] = st[0]; o[2] = st[1]; o[3] = st[2]; o[4] = st[3]; o[5] = st[4];
        }
        return;
    }

    // ============================ GEMM CTA: three independent groups ============================================
    const int q = (w < 4 * HX_NG) ? (w >> 2) : (w - 4 * HX_NG);
    if (tid == 0) {
        for (int qq = 0; qq < HX_NG; qq++) {
            helper_done = 0;
            thin_cnt[qq] = 0;
            for (int s = 0; s < HT_QD; s++) done_cnt[qq][s] = 0;
            for (int s = 0; s < HX_NST; s++) { mbar_init(&full_bar[qq][s], 1); mbar_init(&empty_bar[qq][s], 4); }
            for (int s = 0; s < HT_QD; s++) { mbar_init(&tq_full[qq][s], 1); mbar_init(&tq_empty[qq][s], 4); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int *abort_flag = D.ctl + 1;

    if (w >= 4 * HX_NG) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (w >= 5 * HX_NG || lane != 0) return;             // one elected lane per group drives the TMA unit
        // ============================ producer lane ==========================================================
        double *ring_base = smem + q * (HX_NST * HX_STAGE_DOUBLES);
        uint64_t *fullb = full_bar[q], *emptyb = empty_bar[q], *tqf = tq_full[q], *tqe = tq_empty[q];
        int4 *tqs = tq[q];
        const int rowsA = (int)(P.bstride / P.ld), rowsW = (int)(P.wstride / P.ld);
        if (q == 0 && D.nf0 > 0) {                           // group 0's ring is the scratch of its column-0 POTRFs first
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile int *>(&helper_done) == 0) {
                __nanosleep(200);
                if (clock64() - t0 > DF_WATCHDOG_CYCLES) { atomicExch(abort_flag, 1); break; }
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        RingState ring = {0, 0u};
        for (int n = 0;; n++) {
            const int slot = n % HT_QD;
            int4 tk = make_int4(-1, 0, 0, 0);
            bool ok;
            { HT_T0(); ok = mbar_wait(&tqe[slot], (((unsigned)(n / HT_QD)) & 1u) ^ 1u, abort_flag); HT_ACC(0); }
            if (!ok) break;
            {
                const int tix = atomicAdd(D.ctl, 1);
                if (tix < D.ngtasks) tk = D.gtasks[tix];
                tqs[slot] = tk;
                mbar_arrive(&tqf[slot]);
            }
            if (tk.x < 0) break;
            const int i = tk.x, k = tk.y, b = tk.z;
            const bool diag = (i == k), thin = (tk.w & 1) != 0, pre = (tk.w & 2) != 0;
            const int nj = pre ? (diag ? k - 2 : k - 1) : k;    // chain mode: a pre tile leaves its last two terms to the chain CTA
            const bool border = (i >= P.T);
            const CUtensorMap *mapI = border ? (thin ? &maps.W8 : &maps.W) : &maps.A;
            const int rowI = border ? b * rowsW + (i - P.T) * GSUM_TILE : b * rowsA + i * GSUM_TILE;
            const int rowK = b * rowsA + k * GSUM_TILE;
            const int *frow_i = D.flags + ((int64_t)b * P.Trows + i) * P.T;
            const int *frow_k = D.flags + ((int64_t)b * P.Trows + k) * P.T;
            const unsigned abytes = thin ? 8 * 128 : 64 * 128;          // one A box
            // ---- stage 0 of the task: the C tile (original data, written before the launch) ----------------------
            {
                { HT_T0(); ok = mbar_wait(&emptyb[ring.stage], ring.phase ^ 1u, abort_flag); HT_ACC(2); }
                if (!ok) break;
                double *Cs = ring_base + ring.stage * HX_STAGE_DOUBLES;
                mbar_expect_tx(&fullb[ring.stage], 4 * abytes);
#pragma unroll
                for (int s = 0; s < 4; s++) hx_tma_2d(Cs + s * HX_BOX_DOUBLES, mapI, k * GSUM_TILE + 16 * s, rowI, &fullb[ring.stage]);
                hx_ring_advance(ring);
            }
            // ---- operand half-slabs ----------------------------------------------------------------------------
            // A finished tile (r, k-1) implies every (r, j < k-1): they were its operands.
            bool done_i = true, done_k = true;
            if (nj > 0) {
                done_i = ld_relaxed(frow_i + nj - 1) >= 1;
                done_k = diag ? done_i : (ld_relaxed(frow_k + nj - 1) >= 1);
            }
            for (int h = 0; h < 2 * nj; h++) {
                const int j = h >> 1;
                if ((h & 1) == 0 && !(done_i && done_k)) {
                    HT_T0();
                    ok = (done_i || flag_wait(frow_i + j, abort_flag)) && (diag || done_k || flag_wait(frow_k + j, abort_flag));
                    HT_ACC(1);
                    if (!ok) break;
                }
                { HT_T0(); ok = mbar_wait(&emptyb[ring.stage], ring.phase ^ 1u, abort_flag); HT_ACC(2); }
                if (!ok) break;
                // the tiles were written by other SMs through the generic proxy: order them before the async-proxy reads
                asm volatile("fence.proxy.async;" ::: "memory");
                double *As = ring_base + ring.stage * HX_STAGE_DOUBLES;
                const int col0 = j * GSUM_TILE + (h & 1) * GSUM_KH;
                mbar_expect_tx(&fullb[ring.stage], 2 * abytes + (diag ? 0u : 2u * 64 * 128));
                hx_tma_2d(As, mapI, col0, rowI, &fullb[ring.stage]);
                hx_tma_2d(As + HX_BOX_DOUBLES, mapI, col0 + 16, rowI, &fullb[ring.stage]);
                if (!diag) {
                    hx_tma_2d(As + 2 * HX_BOX_DOUBLES, &maps.A, col0, rowK, &fullb[ring.stage]);
                    hx_tma_2d(As + 3 * HX_BOX_DOUBLES, &maps.A, col0 + 16, rowK, &fullb[ring.stage]);
                }
                hx_ring_advance(ring);
            }
            if (!ok) break;
            // ---- last stage of a panel task: M_kk -----------------------------------------------------------------
            if (!diag && !pre) {
                { HT_T0(); ok = flag_wait_ge(frow_k + k, 2, abort_flag); HT_ACC(3); }
                if (!ok) break;
                { HT_T0(); ok = mbar_wait(&emptyb[ring.stage], ring.phase ^ 1u, abort_flag); HT_ACC(2); }
                if (!ok) break;
                asm volatile("fence.proxy.async;" ::: "memory");
                double *Ms = ring_base + ring.stage * HX_STAGE_DOUBLES;
                const int rowM = (b * P.T + k) * GSUM_TILE;
                mbar_expect_tx(&fullb[ring.stage], 4u * 64 * 128);
#pragma unroll
                for (int s = 0; s < 4; s++) hx_tma_2d(Ms + s * HX_BOX_DOUBLES, &maps.M, 16 * s, rowM, &fullb[ring.stage]);
                hx_ring_advance(ring);
            }
        }
        if (st_on) { long long *o = D.stats + (int64_t)blockIdx.x * HT_NSTAT + q * 12; o[0] = clock64() - st_t0; o[1] = st[0]; o[2] = st[1]; o[3] = st[2]; o[4] = st[3]; }
    } else {
        // ============================ math warps ============================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " HT_STR(HX_MATH_REGS) ";");
        double *ring_base = smem + q * (HX_NST * HX_STAGE_DOUBLES);
        uint64_t *fullb = full_bar[q], *emptyb = empty_bar[q], *tqf = tq_full[q], *tqe = tq_empty[q];
        int4 *tqs = tq[q];
        if (q == 0 && D.nf0 > 0) {
            // column-0 diagonal tiles: this group factors some of them before its first GEMM task (see ht_factor_worker)
            ht_factor_worker(D, ring_base, nullptr, true);
            if (tid == 0) { __threadfence_block(); *reinterpret_cast<volatile int *>(&helper_done) = 1; }
        }
        const int g = lane >> 2, t = lane & 3, wg = w & 3;
        int oa[4], om[2][2], oc[2];
        {
            const int x = g ^ ((t >> 1) << 2);
#pragma unroll
            for (int a = 0; a < 4; a++) oa[a] = ((a ^ x) << 1) | (t & 1);
#pragma unroll
            for (int p = 0; p < 2; p++) {
                oc[p] = ((4 * p + t) ^ g) << 1;
#pragma unroll
                for (int h = 0; h < 2; h++) om[p][h] = (((4 * p + 2 * h + (t >> 1)) ^ g) << 1) | (t & 1);
            }
        }
        RingState ring = {0, 0u};
        double *thin_x = smem + HX_NG * HX_NST * HX_STAGE_DOUBLES + q * HX_THINX_DOUBLES;
        int thin_n = 0;                                      // thin tasks this warp has seen (buffer parity, arrival target)
        for (int n = 0;; n++) {
            const int slot = n % HT_QD;
            bool alive;
            { HT_T0(); alive = mbar_wait(&tqf[slot], ((unsigned)(n / HT_QD)) & 1u, abort_flag); HT_ACC(0); }
            int4 tk = make_int4(-1, 0, 0, 0);
            if (alive) {
                tk = tqs[slot];
                __syncwarp();
                if (lane == 0) mbar_arrive(&tqe[slot]);
            }
            if (!alive || tk.x < 0) break;             // no CTA-level barrier anywhere in this role: a warp may leave alone
            const int i = tk.x, k = tk.y, b = tk.z;
            const bool diag = (i == k), thin = (tk.w & 1) != 0, pre = (tk.w & 2) != 0;
            const int nj = pre ? (diag ? k - 2 : k - 1) : k;
            double *Ab = P.A + (int64_t)b * P.bstride;
            double *Ri = (i < P.T) ? Ab + (int64_t)i * GSUM_TILE * P.ld
                                   : P.W + (int64_t)b * P.wstride + (int64_t)(i - P.T) * GSUM_TILE * P.ld;
            double *C = Ri + k * GSUM_TILE;
            HT_TRACE(0);
            if (STATS && D.trace && (int)blockIdx.x == D.trace_cta && (w & 3) == 0 && lane == 0 && n < 64) { D.trace[(q * 64 + n) * 8 + 6] = i * 1000 + k; D.trace[(q * 64 + n) * 8 + 7] = b; }
            const bool active = !thin || wg == 0;           // thin task (rows 0..7 in use): warp 0 of the group alone
            // rows of this warp's two m tiles: 16 wg, 16 wg + 8 — or, on a diagonal task, the block rows wg and 7 - wg
            const int row0 = diag ? wg * 8 : wg * 16, row1 = diag ? (7 - wg) * 8 : wg * 16 + 8;
            const int n0 = diag ? wg + 1 : 8, n1 = diag ? 8 - wg : 8;           // n tiles in use per m tile
            // ---- acc = -C ------------------------------------------------------------------------------------------
            Acc acc;
            {
                { HT_T0(); alive = mbar_wait(&fullb[ring.stage], ring.phase, abort_flag); HT_ACC(1); }
                if (!alive) break;
                HT_TRACE(1);
                const double *Cs = ring_base + ring.stage * HX_STAGE_DOUBLES;
                if (thin) {
#pragma unroll
                    for (int mt = 0; mt < 2; mt++)
#pragma unroll
                        for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const double2 v = *reinterpret_cast<const double2 *>(Cs + wg * HX_BOX_DOUBLES + g * 16 + oc[e]);
                        acc[0][e][0] = -v.x; acc[0][e][1] = -v.y;
                    }
                } else
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        if (nt < (mt ? n1 : n0)) {
                            const double2 v = *reinterpret_cast<const double2 *>(Cs + (nt >> 1) * HX_BOX_DOUBLES + ((mt ? row1 : row0) + g) * 16 + oc[nt & 1]);
                            acc[mt][nt][0] = -v.x; acc[mt][nt][1] = -v.y;
                        } else { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyb[ring.stage]);
                hx_ring_advance(ring);
            }
            // ---- main loop ---------------------------------------------------------------------------------------
            for (int h = 0; h < 2 * nj; h++) {
                { HT_T0(); alive = mbar_wait(&fullb[ring.stage], ring.phase, abort_flag); HT_ACC(1); }
                if (!alive) break;
                const double *As = ring_base + ring.stage * HX_STAGE_DOUBLES;
                if (thin) hx_stage_mma_thin(acc, As, As + 2 * HX_BOX_DOUBLES, wg, g, oa);
                else if (diag) {
                    if (wg == 0) hx_stage_mma_diag<1>(acc, As, g, oa);
                    else if (wg == 1) hx_stage_mma_diag<2>(acc, As, g, oa);
                    else if (wg == 2) hx_stage_mma_diag<3>(acc, As, g, oa);
                    else hx_stage_mma_diag<4>(acc, As, g, oa);
                }
                else hx_stage_mma(acc, As, As + 2 * HX_BOX_DOUBLES, row0, row1, g, oa);
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyb[ring.stage]);
                hx_ring_advance(ring);
            }
            if (!alive) break;
            HT_TRACE(2);
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = -acc[mt][nt][0]; acc[mt][nt][1] = -acc[mt][nt][1]; }
            if (thin) {
                // ---- gather the 8 x 64 row block in warp 0.  Two buffers (task parity): the other warps can run at most one
                // task ahead of warp 0 — a task takes at least two ring stages and warp 0's arrival frees each of them.
                double *xb = thin_x + (thin_n & 1) * (HX_THINX_DOUBLES / 2);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double2 v; v.x = acc[0][e][0]; v.y = acc[0][e][1];
                    *reinterpret_cast<double2 *>(xb + g * 64 + (2 * wg + e) * 8 + 2 * t) = v;
                }
                thin_n++;
                __syncwarp();
                if (wg != 0) {
                    if (lane == 0) asm volatile("red.release.cta.shared.add.s32 [%0], 1;" ::"r"(smem_u32(&thin_cnt[q])) : "memory");
                } else {
                    int ok = 1;
                    if (lane == 0) {
                        const int want = 3 * thin_n;
                        int v, spins = 0;
                        long long t0 = 0;
                        for (;;) {
                            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&thin_cnt[q])) : "memory");
                            if (v >= want) break;
                            if ((++spins & 255) == 0) {
                                if (ld_relaxed(abort_flag)) { ok = 0; break; }
                                if (t0 == 0) t0 = clock64();
                                else if (clock64() - t0 > DF_WATCHDOG_CYCLES) { atomicExch(abort_flag, 1); ok = 0; break; }
                            }
                        }
                    }
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (!ok) break;
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        const double2 v = *reinterpret_cast<const double2 *>(xb + g * 64 + nt * 8 + 2 * t);
                        acc[0][nt][0] = v.x; acc[0][nt][1] = v.y;
                    }
                }
            }
            if (diag) {
                // ---- S back in place; the factor CTAs take it from there ------------------------------------------
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++)
                        if (nt < (mt ? n1 : n0)) {
                            double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                            *reinterpret_cast<double2 *>(C + (int64_t)((mt ? row1 : row0) + g) * P.ld + nt * 8 + 2 * t) = v;
                        }
            } else if (pre) {
                // ---- chain mode: S' in place, the chain worker does the triangular solve -----------------------------
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                        *reinterpret_cast<double2 *>(C + (int64_t)(wg * 16 + mt * 8 + g) * P.ld + nt * 8 + 2 * t) = v;
                    }
            } else {
                // ---- the triangular solve, warp-local on the 16 x 64 row block ------------------------------------
                { HT_T0(); alive = mbar_wait(&fullb[ring.stage], ring.phase, abort_flag); HT_ACC(2); }
                if (!alive) break;
                HT_TRACE(3);
                if (active) {
                    HT_T0();
                    const double *Ms = ring_base + ring.stage * HX_STAGE_DOUBLES;
                    if (thin) hx_trsm_dinv<1>(acc, Ms, g, t, om); else hx_trsm_dinv<2>(acc, Ms, g, t, om);
                    HT_ACC(3);
#pragma unroll
                    for (int mt = 0; mt < 2; mt++)
#pragma unroll
                        for (int nt = 0; nt < 8; nt++)
                            if (!thin || mt == 0) {
                                double2 v; v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
                                *reinterpret_cast<double2 *>(C + (int64_t)(wg * 16 + mt * 8 + g) * P.ld + nt * 8 + 2 * t) = v;
                            }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyb[ring.stage]);
                hx_ring_advance(ring);
            }
            HT_TRACE(4);
            // ---- publish the tile: the last of the four warps to get here stores the flag (see hetero.cuh) ---------------
            { HT_T0();
            __syncwarp();
            if (lane == 0) {
                int old;
                asm volatile("atom.acq_rel.cta.shared.add.s32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done_cnt[q][slot])) : "memory");
                if (old == 3) {
                    done_cnt[q][slot] = 0;
                    st_release((pre && !diag) ? D.pre + (int64_t)b * P.T + k : D.flags + ((int64_t)b * P.Trows + i) * P.T + k, 1);
                }
            }
            HT_ACC(4); }
            HT_TRACE(5);
            st[5] += 1;
        }
        if (st_on && (tid & 127) == 0) { long long *o = D.stats + (int64_t)blockIdx.x * HT_NSTAT + q * 12; o[5] = clock64() - st_t0; o[6] = st[0]; o[7] = st[1]; o[8] = st[2]; o[9] = st[3]; o[10] = st[4]; o[11] = st[5]; }
    }
#undef HT_T0
#undef HT_ACC
#undef HT_TRACE
}
