// FP64 roofline denominators for B200 (sm_100a): register-resident DFMA and DMMA.8x8x4
// issue-rate microbenchmarks.  MEASURED_PEAKS.json carries no FP64 figure, so bench.py
// measures cuBLAS DGEMM live and this tool pins the pipe ceilings the kernels are judged by.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters) {
    double acc[ILP];
    double a = 1.0000001 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double *out, int iters) {
    double c[ILP][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    printf("device %s sms %d\n", p.name, sms);
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32, blocks = sms;
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters); });
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("DFMA ilp8 warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 256 * 8 * iters * (double)warps * blocks;
        printf("DMMA.8x8x4 ilp8 warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 256 * 2 * iters * (double)warps * blocks;
        printf("DMMA.8x8x4 ilp2 warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
    }
    return 0;
}
