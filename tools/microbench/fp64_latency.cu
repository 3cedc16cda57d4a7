// Dependent-issue latencies of the instructions on the Cholesky pivot chain (one warp, one SM), in SM cycles:
// DFMA, DMUL, MUFU.RSQ64H (rsqrt.approx.ftz.f64), 64-bit SHFL (2 x SHFL.IDX), LDS.64, and the issue interval of
// independent DFMAs for 1 / 2 / 4 warps per scheduler.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   -o tools/microbench/fp64_latency.bin tools/microbench/fp64_latency.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 2048;

__global__ void lat_kernel(double* out, long long* cyc, double seed)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    double a = seed + lane * 1e-9, b = 1.0 + 1e-12, c = 1e-13;
    sm[lane] = (double)((lane + 1) & 31);
    sm[32 + lane] = 0.0;
    __syncthreads();
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // rsqrt.approx chain (seed only)
    double r = fabs(a) + 1.5;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(r));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 64-bit shuffle chain
    double s = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // LDS.64 pointer chase
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx & 31];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // full Newton rsqrt (seed + 2 steps) chain
    double q = fabs(a) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 4; ++i) {
        double y;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
        double t = q * y, e = fma(-t, y, 1.0);
        y = fma(0.5 * y, e, y);
        t = q * y;
        e = fma(-t, y, 1.0);
        q = fma(0.5 * y, e, y) + 2.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    out[threadIdx.x] = a + r + s + idx + q;
}

// independent DFMAs: 8 accumulators per thread, W warps per scheduler (block = 128 W threads)
__global__ void thr_kernel(double* out, long long* cyc, double seed)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = seed + i + threadIdx.x * 1e-9;
    const double b = 1.0 + 1e-12, c = 1e-13;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < N / 8; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], b, c);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[threadIdx.x] = s;
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 4096 * 8);
    cudaMalloc(&cyc, 64 * 8);
    long long h[8];
    for (int rep = 0; rep < 2; ++rep) lat_kernel<<<1, 32>>>(out, cyc, 1.0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 6 * 8, cudaMemcpyDeviceToHost);
    printf("{\"dfma_latency\": %.2f, \"dmul_latency\": %.2f, \"rsqrt64h_latency\": %.2f, \"shfl64_latency\": %.2f, "
           "\"lds64_f2i_chase\": %.2f, \"newton_rsqrt_chain\": %.2f",
           (double)h[0] / N, (double)h[1] / N, (double)h[2] / N, (double)h[3] / N, (double)h[4] / N,
           (double)h[5] / (N / 4));
    for (int w = 1; w <= 4; w *= 2) {
        for (int rep = 0; rep < 2; ++rep) thr_kernel<<<1, 128 * w>>>(out, cyc, 1.0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf(", \"dfma_issue_interval_%dwarp_per_sched\": %.2f", w, (double)h[0] / N);
    }
    cudaError_t e = cudaGetLastError();
    printf(", \"cuda\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
