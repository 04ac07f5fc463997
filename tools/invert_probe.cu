// Dev probe: cycles of the 8x8 block inverses (chain_invert_blocks) alone on an SM, and beside other warps' activity.
#include <cstdio>
#include <cuda_runtime.h>
#include "../gsum_b200/csrc/chain.cuh"
__global__ void probe(long long *out, double *sink, int mode) {
    __shared__ __align__(16) double S[GSUM_TILE * GSUM_LDS];
    __shared__ double rinv[64];
    __shared__ double Dv[8 * CH_DV_BLOCK];
    const int tid = threadIdx.x;
    for (int e = tid; e < GSUM_TILE * GSUM_LDS; e += blockDim.x) S[e] = 0.001 * ((e * 7) % 13);
    if (tid < 64) rinv[tid] = 1.0 / (1.0 + tid);
    __syncthreads();
    if (tid < 128) {
        long long t0 = 0, t1 = 0, cold = 0;
        for (int it = 0; it < 3; it++) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            t0 = clock64();
            chain_invert_blocks(S, rinv, Dv, false);
            t1 = clock64();
            if (it == 0) cold = t1 - t0;
        }
        if (tid == 0) { out[mode] = t1 - t0; out[4 + mode] = cold; }
        if (tid == 0) sink[0] = Dv[5];
    } else if (mode == 1) {
        // helper-like: one dependent DMMA chain per warp
        double c0 = 0, c1 = 0;
        for (int i = 0; i < 4000; i++) dmma884(c0, c1, S[(i & 63) * GSUM_LDS + (tid & 3)], S[(i & 31) * GSUM_LDS + 4 + (tid & 3)]);
        sink[tid] = c0 + c1;
    } else if (mode == 2) {
        for (int i = 0; i < 2000; i++) __nanosleep(100);
    }
}
int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 8 * 1024);
    for (int mode = 0; mode < 3; mode++) { probe<<<1, mode == 0 ? 128 : 256>>>(out, sink, mode); cudaDeviceSynchronize(); }
    printf("invert alone %lld cycles | beside 4 ILP-1 DMMA warps %lld | beside 4 sleeping warps %lld\n", out[0], out[1], out[2]);
    printf("first (cold instruction cache) pass: %lld | %lld | %lld\n", out[4], out[5], out[6]);
    return 0;
}
