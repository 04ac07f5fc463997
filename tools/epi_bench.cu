// Dev probe: the two epilogues (POTRF of a 64x64 tile, TRSM of a 64x64 tile against L_kk) timed in isolation on one
// CTA, alone and while four more warps (one per SM sub-partition) run the DMMA main loop — i.e. what a co-resident CTA
// does to them.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/epi_bench tools/epi_bench.cu
#include "../gsum_b200/csrc/chol.cuh"
#include <vector>
#include <cstdlib>

// mode 0: POTRF; 1: trsm_rows; 2: trsm_regs.   hog: 0 none, 1 = warps 4..7 stream stage_mma
__global__ void __launch_bounds__(256, 1) epi_kernel(int mode, int hog, const double *A, const double *L, int reps, long long *out, double *sink) {
    extern __shared__ __align__(16) double smem[];
    __shared__ volatile int done;
    double *S = smem, *Lk = smem + 4608, *ops = smem + 2 * 4608;
    const int tid = threadIdx.x, w = tid >> 5;
    for (int e = tid; e < 64 * 64; e += 256) { Lk[(e >> 6) * GSUM_LDS + (e & 63)] = L[e]; }
    for (int e = tid; e < 4608; e += 256) ops[e] = 1e-3 * (e % 7);
    if (tid == 0) done = 0;
    __syncthreads();
    if (w < 4) {
        long long tot = 0;
        double chk = 0;
        for (int r = 0; r < reps; r++) {
            Acc acc;
            const int lane = tid & 31, g = lane >> 2, t = lane & 3;
            for (int mt = 0; mt < 2; mt++) for (int nt = 0; nt < 8; nt++) {
                acc[mt][nt][0] = A[(w * 16 + mt * 8 + g) * 64 + nt * 8 + 2 * t]; acc[mt][nt][1] = A[(w * 16 + mt * 8 + g) * 64 + nt * 8 + 2 * t + 1];
            }
            if (mode == 0) {
                for (int e = tid; e < 64 * 64; e += 128) S[(e >> 6) * GSUM_LDS + (e & 63)] = A[e];
                double *dg = S + 64 * GSUM_LDS; int *sf = (int *)(dg + 128);
                if (tid == 0) *sf = 0;
                CONS_SYNC();
                const long long t0 = clock64();
                tile_potrf_blocked(S, dg, sf);
                tot += clock64() - t0;
                chk += S[(tid & 63) * GSUM_LDS + (tid & 31)];
            } else {
                double *rdiag = S, *Lp = S + 64, *scr = S + 64 + 512 + w * (16 * TRSM_SCR_LD);
                CONS_SYNC();
                const long long t0 = clock64();
                trsm_prepare(Lk, Lp, rdiag);
                CONS_SYNC();
                if (mode == 1) trsm_rows<2>(acc, Lk, Lp, rdiag, scr); else trsm_regs<2>(acc, Lk, rdiag);
                CONS_SYNC();
                tot += clock64() - t0;
                chk += acc[0][7][0] + acc[1][3][1];
            }
        }
        if (tid == 0) { out[0] = tot / reps; done = 1; }
        sink[tid] = chk;
    } else if (hog) {
        Acc acc;
        for (int mt = 0; mt < 2; mt++) for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = 0; acc[mt][nt][1] = 0; }
        long long n = 0;
        while (!done) { stage_mma<true>(acc, ops, ops + 64 * GSUM_LDH, 8); n++; }
        double sum = 0;
        for (int mt = 0; mt < 2; mt++) for (int nt = 0; nt < 8; nt++) sum += acc[mt][nt][0] + acc[mt][nt][1];
        sink[tid] = sum;
        if (tid == 128) out[1] = n;
    }
}

int main() {
    const int n = 64;
    std::vector<double> A(n * n), L(n * n, 0.0);
    // SPD tile: smooth kernel + nugget, and its Cholesky factor (host, plain loops)
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double d = (i - j) * 0.02; A[i * n + j] = exp(-0.5 * d * d) + (i == j ? 1e-4 : 0.0); }
    std::vector<double> W = A;
    for (int j = 0; j < n; j++) {
        double d = W[j * n + j];
        for (int m = 0; m < j; m++) d -= L[j * n + m] * L[j * n + m];
        L[j * n + j] = sqrt(d);
        for (int i = j + 1; i < n; i++) { double v = W[i * n + j]; for (int m = 0; m < j; m++) v -= L[i * n + m] * L[j * n + m]; L[i * n + j] = v / L[j * n + j]; }
    }
    double *dA, *dL, *sink; long long *out;
    cudaMalloc(&dA, 8 * n * n); cudaMalloc(&dL, 8 * n * n); cudaMalloc(&sink, 8 * 512); cudaMallocManaged(&out, 64);
    cudaMemcpy(dA, A.data(), 8 * n * n, cudaMemcpyHostToDevice); cudaMemcpy(dL, L.data(), 8 * n * n, cudaMemcpyHostToDevice);
    const int smem = 3 * 4608 * 8;
    cudaFuncSetAttribute(epi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char *names[] = {"POTRF 64x64", "TRSM rows (smem row-per-thread)", "TRSM regs (quad shuffles)"};
    for (int mode = 0; mode < 3; mode++)
        for (int hog = 0; hog < 2; hog++) {
            out[0] = out[1] = 0;
            epi_kernel<<<1, 256, smem>>>(mode, hog, dA, dL, 20, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%-34s %s: %8lld cycles   (MMA_PASSES %d, hog stages %lld)\n", names[mode], hog ? "under a DMMA main loop" : "alone                 ", out[0], MMA_PASSES, out[1]);
        }
    return 0;
}
