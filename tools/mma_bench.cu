// Dev probe: cycles per pipeline stage (K depth 32, 64x64 CTA tile) of the DMMA main loop, operands resident in smem.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DDMMA_GROUP=n] -o build/mma_bench tools/mma_bench.cu
#include "../gsum_b200/csrc/chol.cuh"
__global__ void __launch_bounds__(256, 1) k(int nwarps, int iters, long long *out, double *sink) {
    extern __shared__ __align__(16) double smem[];
    for (int e = threadIdx.x; e < 2 * 4608; e += blockDim.x) smem[e] = 1e-3 * (e % 11);
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (w >= nwarps) return;
    Acc acc;
    for (int mt = 0; mt < 2; mt++) for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = 0; acc[mt][nt][1] = 0; }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const double *As = smem + (i & 1) * 4608;
        stage_mma<true>(acc, As, As + 64 * GSUM_LDH, 8);
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[w] = t1 - t0;
    double sum = 0;
    for (int mt = 0; mt < 2; mt++) for (int nt = 0; nt < 8; nt++) sum += acc[mt][nt][0] + acc[mt][nt][1];
    sink[threadIdx.x] = sum;
}
int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 64 * 8); cudaMalloc(&sink, 8 * 256);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4608 * 8);
    for (int grid : {1, 148})
        for (int nw : {1, 4, 8})
            for (int iters : {800}) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                k<<<grid, 256, 2 * 4608 * 8>>>(nw, iters, out, sink);
                cudaEventRecord(e1);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("MMA_PASSES %d grid %3d  %d warps/CTA  iters %3d: %.0f cycles per stage per warp | kernel %.3f ms -> %.2f TFLOP/s\n", MMA_PASSES, grid, nw, iters,
                       (double)out[0] / iters, ms, (double)grid * nw * iters * 128 * 512 / (ms * 1e-3) * 1e-12);
            }
    return 0;
}
