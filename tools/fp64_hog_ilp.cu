// Dev probe: what a dependent FP64 chain (DFMA / rsqrt) suffers on one SM sub-partition beside a warp that issues DMMAs with a
// LIMITED number of independent accumulators (ILP 1, 2, 4, 8, 16).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_hog_ilp tools/fp64_hog_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void probe(int mode, int iters, long long *out, double *sink, double seed) {
    __shared__ volatile int done;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (w == 0) {
        double x = seed, y = 1.0 + seed * 1e-9;
        for (int i = 0; i < 2000; i++) asm volatile("nanosleep.u32 20;");
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (mode == 0) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
            else { x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); x = rsqrt(x + 2.0); }
        }
        const long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; done = 1; }
        sink[threadIdx.x] = x;
    } else if (w == 4) {           // same sub-partition as warp 0
        double acc[ILP][2];
        for (int i = 0; i < ILP; i++) { acc[i][0] = 0; acc[i][1] = 0; }
        double a = seed, b = seed * 0.5;
        long long n = 0;
        const long long t0 = clock64();
        while (!done) {
#pragma unroll
            for (int r = 0; r < 64 / ILP; r++)
#pragma unroll
                for (int i = 0; i < ILP; i++) dmma884(acc[i][0], acc[i][1], a, b);
            n += 64;
        }
        const long long t1 = clock64();
        double s = 0;
        for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
        sink[threadIdx.x] = s;
        if (lane == 0) { out[1] = n; out[2] = t1 - t0; }
    }
}
template <int ILP> void run(long long *out, double *sink) {
    const char *names[] = {"DFMA", "rsqrt+DADD"};
    for (int mode = 0; mode < 2; mode++) {
        out[0] = out[1] = out[2] = 0;
        probe<ILP><<<1, 256>>>(mode, 2000, out, sink, 1.0000001);
        cudaDeviceSynchronize();
        printf("hog ILP %2d: %-10s chain %7.1f cycles/op | hog: %.1f cycles per DMMA\n", ILP, names[mode], (double)out[0] / 8000.0, (double)out[2] / (double)out[1]);
    }
}
int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 64); cudaMalloc(&sink, 8 * 1024);
    run<1>(out, sink); run<2>(out, sink); run<3>(out, sink); run<4>(out, sink); run<8>(out, sink); run<16>(out, sink);
    return 0;
}
