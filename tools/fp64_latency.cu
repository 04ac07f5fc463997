// Dev probe: latency of dependent FP64 instruction chains on one SM, alone and while another warp streams DMMAs
// on the same / another SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode: 0 DFMA chain, 1 DMMA chain, 2 rsqrt chain, 3 sqrt+div chain, 4 shfl64 chain, 5 DMUL chain, 6 LDS chain
// hog_warp: -1 none; else that warp streams independent DMMAs (ILP 16) for the whole measurement
__global__ void probe(int mode, int hog_warp, int hog_warp2, int iters, long long *out, double *sink, double seed) {
    __shared__ volatile int go, done;
    __shared__ double sm[64];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { go = 0; done = 0; }
    if (threadIdx.x < 64) sm[threadIdx.x] = (threadIdx.x + 1) % 64;
    __syncthreads();
    if (w == 0) {
        double x = seed, y = 1.0 + seed * 1e-9, c0 = 0, c1 = 0;
        int idx = lane;
        // let the hog get going
        for (int i = 0; i < 2000; i++) asm volatile("nanosleep.u32 20;");
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (mode == 0) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
            else if (mode == 1) { dmma884(c0, c1, x, y); dmma884(c0, c1, x, y); dmma884(c0, c1, x, y); dmma884(c0, c1, x, y); }
            else if (mode == 2) { x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); }
            else if (mode == 3) { x = 1.0 / sqrt(x + 2.0); x = 1.0 / sqrt(x + 2.0); x = 1.0 / sqrt(x + 2.0); x = 1.0 / sqrt(x + 2.0); }
            else if (mode == 4) { x = __shfl_sync(0xffffffffu, x, 1, 4); x = __shfl_sync(0xffffffffu, x, 2, 4); x = __shfl_sync(0xffffffffu, x, 3, 4); x = __shfl_sync(0xffffffffu, x, 0, 4); }
            else if (mode == 5) { x = x * y; x = x * y; x = x * y; x = x * y; }
            else { idx = (int)sm[idx & 63]; idx = (int)sm[idx & 63]; idx = (int)sm[idx & 63]; idx = (int)sm[idx & 63]; }
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; done = 1; }
        sink[threadIdx.x] = x + c0 + c1 + idx;
    } else if (w == hog_warp || w == hog_warp2) {
        double acc[16][2];
        for (int i = 0; i < 16; i++) { acc[i][0] = 0; acc[i][1] = 0; }
        double a = seed, b = seed * 0.5;
        long long n = 0;
        while (!done) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) dmma884(acc[i][0], acc[i][1], a, b);
            n += 64;
        }
        double s = 0;
        for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
        sink[threadIdx.x] = s;
        if (lane == 0) out[1 + (w == hog_warp2 && hog_warp2 != hog_warp)] = n;
    }
}

int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 8 * 1024);
    const char *names[] = {"DFMA", "DMMA", "rsqrt", "1/sqrt", "shfl64", "DMUL", "LDS"};
    const int iters = 2000;
    for (int mode = 0; mode < 7; mode++) {
        for (int cfg = 0; cfg < 4; cfg++) {
            int hog = cfg == 0 ? -1 : (cfg == 1 ? 4 : (cfg == 2 ? 1 : 4)), hog2 = cfg == 3 ? 8 : hog;
            out[0] = out[1] = out[2] = 0;
            probe<<<1, 384>>>(mode, hog, hog2, iters, out, sink, 1.0000001);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            const char *cn[] = {"alone", "DMMA hog on same SMSP (warp 4)", "DMMA hog on other SMSP (warp 1)", "two DMMA hogs on same SMSP (warps 4, 8)"};
            printf("%-7s chain: %7.1f cycles/op   %-42s hog DMMAs/cycle %.4f\n", names[mode], (double)out[0] / (4.0 * iters), cn[cfg],
                   (double)(out[1] + out[2]) / (double)out[0]);
        }
    }
    return 0;
}
