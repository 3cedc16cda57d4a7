// i8_umma_probe.cu — stand-alone probe for VERDICT item 8 (Ozaki split of the fp64 variance contraction onto the int8
// tensor pipe): is tcgen05.mma kind::i8 (s8 x s8 -> s32, TMEM accumulator) usable on this B200, are the descriptors
// right (checked against a CPU integer GEMM), and what is its issue rate?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o i8_umma_probe i8_umma_probe.cu -lcuda
//   ./i8_umma_probe            (prints one JSON line)
//
// Layout: A (128 x K) and B (N x K), int8, K contiguous ("K-major"); TMA box 128 bytes x 128 rows with 128B swizzle
// into 1024-byte aligned shared memory = the canonical K-major SWIZZLE_128B UMMA layout (8-row x 128-byte atoms,
// stride-byte-offset 1024); one MMA = 128 x N x 32; four MMAs per 128-byte k-block, advancing the descriptor start
// address by 32 bytes.  D (128 x N, s32) lives in TMEM: lane = row, column = column.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { printf("{\"error\": \"%s -> %s\"}\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int M = 128, N = 128, KB = 128;          // k-block: 128 int8 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 32;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major) bits [16,30), stride byte offset = 1024 >> 4 in bits [32,46),
// descriptor version 1 (sm_100) bits [46,48), layout type SWIZZLE_128B = 2 in bits [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: D = s32 (c_format 2, bits [4,6)), A and B signed
// 8-bit (format 1, bits [7,10) and [10,13)), both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// mode 0: C = A B^T once (correctness).  mode 1: `reps` passes over the same shared-memory operands (issue-rate probe;
// the accumulator wraps around, nothing is checked).
__global__ void __launch_bounds__(128, 1)
i8_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int K, int* C, int mode,
                int reps, long long* cycles)
{
    extern __shared__ unsigned char raw[];
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    const int nkb = K / KB;
    const uint32_t sA = base, sB = base + (uint32_t)nkb * M * KB;            // all k-blocks resident (K <= 512)
    const uint32_t bar_full = sB + (uint32_t)nkb * N * KB, bar_mma = bar_full + 8, tmem_slot = bar_full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {                      // TMEM: N columns of 32-bit accumulators (power of two >= 32), one warp allocates
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"((uint32_t)N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (tid == 0) {
        mbar_expect_tx(bar_full, (uint32_t)nkb * (M + N) * KB);
        for (int kb = 0; kb < nkb; ++kb) {
            tma_2d(sA + (uint32_t)kb * M * KB, &mapA, kb * KB, 0, bar_full);
            tma_2d(sB + (uint32_t)kb * N * KB, &mapB, kb * KB, 0, bar_full);
        }
        mbar_wait(bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = umma_idesc(M, N);
        const long long t0 = clock64();
        const int passes = mode == 0 ? 1 : reps;
        for (int p = 0; p < passes; ++p)
            for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
                for (int k = 0; k < KB / UMMA_K; ++k)
                    umma_i8(tmem, umma_desc(sA + (uint32_t)kb * M * KB + k * UMMA_K), umma_desc(sB + (uint32_t)kb * N * KB + k * UMMA_K),
                            idesc, (p | kb | k) != 0);
        umma_commit(bar_mma);
        mbar_wait(bar_mma, 0);
        if (cycles) cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
        // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (rows), 32 columns per tcgen05.ld
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) C[(long)row * N + c0 + j] = (int)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)N) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeFn fn, CUtensorMap* map, void* base, int rows, int K) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {(cuuint32_t)KB, 128u};
    cuuint32_t estr[2] = {1, 1};
    return (int)fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
    const int K = 512;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn fn = (EncodeFn)p;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    srand(7);
    for (auto& v : hA) v = (int8_t)(rand() % 255 - 127);
    for (auto& v : hB) v = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB;
    int* dC;
    long long* dcyc;
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    CKC(cudaMalloc(&dA, hA.size()));
    CKC(cudaMalloc(&dB, hB.size()));
    CKC(cudaMalloc(&dC, (size_t)M * N * 4));
    CKC(cudaMalloc(&dcyc, (size_t)nsm * 8));
    CKC(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CKC(cudaMemset(dC, 0xFF, (size_t)M * N * 4));
    CUtensorMap mapA, mapB;
    int r1 = make_map(fn, &mapA, dA, M, K), r2 = make_map(fn, &mapB, dB, N, K);
    if (r1 || r2) { printf("{\"error\": \"cuTensorMapEncodeTiled %d %d\"}\n", r1, r2); return 1; }
    const int smem = (K / KB) * (M + N) * KB + 1024 + 64;
    CKC(cudaFuncSetAttribute(i8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    i8_probe_kernel<<<1, 128, smem>>>(mapA, mapB, K, dC, 0, 1, nullptr);
    CKC(cudaGetLastError());
    CKC(cudaDeviceSynchronize());
    std::vector<int> hC((size_t)M * N);
    CKC(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0, first_bad = -1;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int acc = 0;
            for (int k = 0; k < K; ++k) acc += (int)hA[(size_t)i * K + k] * (int)hB[(size_t)j * K + k];
            if (acc != hC[(size_t)i * N + j]) { if (first_bad < 0) first_bad = (long)i * N + j; ++bad; }
        }
    // issue-rate probe: every SM runs `reps` passes of K = 512 (16 MMAs of 128 x 128 x 32 each)
    const int reps = 2000;
    cudaEvent_t e0, e1;
    CKC(cudaEventCreate(&e0));
    CKC(cudaEventCreate(&e1));
    i8_probe_kernel<<<nsm, 128, smem>>>(mapA, mapB, K, dC, 1, 10, dcyc);
    CKC(cudaDeviceSynchronize());
    CKC(cudaEventRecord(e0));
    i8_probe_kernel<<<nsm, 128, smem>>>(mapA, mapB, K, dC, 1, reps, dcyc);
    CKC(cudaEventRecord(e1));
    CKC(cudaDeviceSynchronize());
    float ms = 0;
    CKC(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(nsm);
    CKC(cudaMemcpy(cyc.data(), dcyc, (size_t)nsm * 8, cudaMemcpyDeviceToHost));
    const double ops = 2.0 * M * N * K * (double)reps * nsm;
    printf("{\"probe\": \"tcgen05.mma kind::i8 128x%dx32, s8 x s8 -> s32 in TMEM\", \"mismatches\": %ld, \"first_bad\": %ld, "
           "\"c00\": %d, \"c_last\": %d, \"sms\": %d, \"ms\": %.4f, \"int8_tops_all_sms\": %.1f, \"cycles_per_mma_sm0\": %.1f}\n",
           N, bad, first_bad, hC[0], hC[(size_t)M * N - 1], nsm, ms, ops / (ms * 1e-3) / 1e12,
           (double)cyc[0] / ((double)reps * (K / UMMA_K)));
    return bad ? 2 : 0;
}
