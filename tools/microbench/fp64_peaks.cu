// Measures the fp64 issue-rate peaks of the GPU it runs on: DFMA (vector pipe) and DMMA m8n8k4 (tensor pipe),
// register-resident operands, no memory traffic.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a[8], b = 1.0000001, c = 0.9999999;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
    double c[16][2], a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* d; cudaMalloc(&d, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = p.multiProcessorCount;
        float ms;
        dfma_kernel<<<blocks, threads>>>(d, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dfma = 2.0 * 8 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
        dmma_kernel<<<blocks, threads>>>(d, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dmma = 2.0 * 256 * 16 * iters * (double)warps * blocks / (ms * 1e-3) / 1e12;
        printf("{\"warps_per_sm\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sms\": %d, \"clock_mhz\": %d}\n",
               warps, dfma, dmma, p.multiProcessorCount, p.clockRate / 1000);
    }
    return 0;
}
