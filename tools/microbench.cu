// Micro-benchmarks that size the FP64 roofline of the box: DMMA.8x8x4 issue rate, DFMA rate,
// as a function of resident warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = threadIdx.x * 1e-3, b = 1.0 + 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; CK(cudaMalloc(&out, 148 * 2048 * 8 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sms = p.multiProcessorCount;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks_per_sm = warps * 32 / threads;
        int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<16><<<sms * blocks_per_sm, threads>>>(out, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256 * 16 * (double)iters * warps * sms;
            if (rep) printf("DMMA.8x8x4 warps/SM=%2d acc=16: %.2f TFLOP/s (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<16><<<sms * blocks_per_sm, threads>>>(out, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
            if (rep) printf("DFMA       warps/SM=%2d acc=16: %.2f TFLOP/s (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
    }
    // dependent-chain latency of DMMA (1 warp, 1 accumulator)
    {
        int iters = 100000;
        cudaEventRecord(e0);
        dmma_kernel<1><<<1, 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA dependent chain: %.1f ns per DMMA\n", ms * 1e6 / iters);
        cudaEventRecord(e0);
        dmma_kernel<4><<<1, 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 1 warp x 4 independent: %.1f ns per DMMA\n", ms * 1e6 / iters / 4);
        cudaEventRecord(e0);
        dmma_kernel<16><<<1, 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 1 warp x 16 independent: %.1f ns per DMMA\n", ms * 1e6 / iters / 16);
    }
    return 0;
}
