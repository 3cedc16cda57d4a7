// gpk_ozaki.cuh — the variance contraction on the int8 tensor pipe (option "ozaki" = 1).
//
// B200's fp64 ceiling is the DMMA / DFMA issue rate (37 TFLOP/s); its 5th-generation tensor cores have no fp64 kind,
// but tcgen05.mma kind::i8 (s8 x s8 -> s32, accumulators in TMEM) runs two orders of magnitude faster.  An
// error-free (Ozaki) split turns the fp64 product V = L^-1 K*^T into exact integer products:
//   P  = L^-1 :  P[i][k]  ~ 2^eP[i] sum_s Pq[s][i][k] 2^(-7 (s+1))     per-row exponent, S = 8 slices of 7 signed bits
//   K*        :  K*[c][k] ~ 2^eK    sum_t Kq[t][c][k] 2^(-7 (t+1))     one exponent (0 < k <= amp)
//   V[i][c] = 2^(eP[i] + eK) sum_lvl 2^(-7 (lvl + 2)) sum_{s + t = lvl} <Pq[s][i][:], Kq[t][c][:]>      (lvl < S)
// The 36 slice pairs with s + t < 8 carry 56 bits of each operand relative to its row maximum; the pairs of one level
// share one int32 accumulator ((lvl + 1) K 127^2 < 2^31 for K <= 16384), so a 128 x 64 tile keeps 8 accumulators of
// 64 columns = all 512 TMEM columns.  Per 64-byte k-block the CTA stages all 8 + 8 slice tiles (96 KB, TMA, 64B
// swizzle) once and issues 72 MMAs (128 x 64 x 32) on them: 26 bytes of operand traffic per 1000 MMA cycles.
// Measured accuracy (tools/ozaki_study.py, tools/microbench/ozaki_probe.cu): posterior variance within 3e-12 .. 3e-11
// (scaled as in the parity tests) of an 80-bit reference while max |L^-1| < 64; the handle falls back to the fp64
// DMMA kernel when the factor is worse conditioned than that (eP > OZ_MAX_EXP) or N > 16384.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocation, warps 2..5 = epilogue (TMEM ->
// fp64 with the level scales, least significant level first; row scale; column reduction sum V^2).
// Output: the per-row-block partial sums part_ssq [nb][ld] the DMMA kernel writes too.  The posterior MEAN does not go
// through the slices: mu - mean = K* alpha is one fp64 dot product of length N per candidate (gpk_rowdot_kernel on the
// fp64 K* that is built anyway, alpha = L^-T z once per fit), like george's own K* alpha; sum_i V_i z_i would put the
// slices' 5e-12 error in front of |z| ~ 1e2 and cost the 1e-10 tolerance on the mean (measured: 1.1e-10 .. 1.9e-10).
#pragma once
#include "gpk_gemm.cuh"

constexpr int OZ_S = 8;                       // slices per operand
constexpr int OZ_TM = 128, OZ_TN = 64;        // tile: 128 rows of L^-1 x 64 candidates
constexpr int OZ_KB = 64;                     // k-block: 64 int8 = one 64-byte swizzle row
constexpr int OZ_UK = 32;                     // K of one kind::i8 MMA
constexpr int OZ_NSTG = 2;
constexpr int OZ_A_SLICE = OZ_TM * OZ_KB, OZ_B_SLICE = OZ_TN * OZ_KB;
constexpr int OZ_STAGE = OZ_S * (OZ_A_SLICE + OZ_B_SLICE);               // 98304 bytes
constexpr int OZ_THREADS = 192;
constexpr int OZ_SMEM = OZ_NSTG * OZ_STAGE + 1024 + 256 + 4 * OZ_TN * 8;
constexpr int OZ_MAX_EXP = 7;                 // row exponents above this (|L^-1| >= 64): use the fp64 kernel

__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, stride byte
// offset 512 (8 rows x 64 bytes), descriptor version 1 (sm_100), layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::i8: D = s32 (bits [4,6) = 2), A / B signed 8-bit (bits [7,10), [10,13) = 1),
// both K-major, N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__host__ __device__ constexpr uint32_t oz_idesc(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void oz_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand split ------------------------------------------------------------------------------------------------
// per-row exponent e[r] with |A[r][:]| / 2^e[r] < 1/2; emax receives the maximum over the rows (atomicMax)
__global__ void gpk_oz_rowexp_kernel(const double* __restrict__ A, long ld, int cols, int* __restrict__ e, int* __restrict__ emax) {
    const long r = blockIdx.x;
    double m = 0.0;
    for (int c = threadIdx.x; c < cols; c += 256) m = fmax(m, fabs(A[r * ld + c]));
    __shared__ double sh[256];
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) {
        int ex = 0;
        if (sh[0] > 0.0) { frexp(sh[0], &ex); ex += 1; }
        e[r] = ex;
        atomicMax(emax, ex);
    }
}
// q[s][row][col] (slices slice_stride bytes apart) = the s-th 7-bit digit of A[row][col] / 2^e; e = erow[row] or (erow == NULL) e0.
// One thread per element; truncation towards zero keeps |q| <= 127 and the remainder's sign.
__global__ void gpk_oz_split_kernel(const double* __restrict__ A, long rows, long ld, const int* __restrict__ erow, int e0,
                                    int8_t* __restrict__ q, long slice_stride) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ld) return;
    const long r = idx / ld;
    double v = ldexp(A[idx], -(erow ? erow[r] : e0));
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) {
        v *= 128.0;
        const double t = trunc(v);
        v -= t;
        q[(long)s * slice_stride + idx] = (int8_t)(int)t;
    }
}

// ---- the contraction ----------------------------------------------------------------------------------------------
struct OzArgs {
    int nb, ncb;                        // row blocks of L^-1 (128 rows), candidate blocks of the chunk (64 candidates)
    int NP, rows;                       // L^-1 is NP x NP; the K* slices have `rows` rows each
    const int* eP; int eK;
    double* part_ssq; long ldpart;
};

__global__ void __launch_bounds__(OZ_THREADS, 1)
gpk_oz_vargemm_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapK, const OzArgs g)
{
    extern __shared__ unsigned char oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + OZ_NSTG * OZ_STAGE, bar_empty = bar_full + 8 * OZ_NSTG, bar_tmem = bar_empty + 8 * OZ_NSTG;
    const uint32_t tmem_slot = bar_tmem + 8;
    const uint32_t red = base + OZ_NSTG * OZ_STAGE + 256;            // [4 lane groups][64 columns] partial sums of V^2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ib = g.nb - 1 - (int)blockIdx.x / g.ncb, cb = (int)blockIdx.x % g.ncb;     // longest contractions first
    const int nkb = (ib + 1) * OZ_TM / OZ_KB;                                             // lower triangle only

    if (tid == 0) {
        for (int s = 0; s < OZ_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % OZ_NSTG;
                if (kb >= OZ_NSTG) oz_mbar_wait(bar_empty + 8 * s, (uint32_t)((kb / OZ_NSTG - 1) & 1));
                const uint32_t st = base + s * OZ_STAGE;
                mbar_arrive_expect_tx(bar_full + 8 * s, OZ_STAGE);
#pragma unroll
                for (int q = 0; q < OZ_S; ++q) {
                    tma_load_2d(st + q * OZ_A_SLICE, &mapP, kb * OZ_KB, q * g.NP + ib * OZ_TM, bar_full + 8 * s);
                    tma_load_2d(st + OZ_S * OZ_A_SLICE + q * OZ_B_SLICE, &mapK, kb * OZ_KB, q * g.rows + cb * OZ_TN, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = oz_idesc(OZ_TM, OZ_TN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % OZ_NSTG;
                oz_mbar_wait(bar_full + 8 * s, (uint32_t)((kb / OZ_NSTG) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * OZ_STAGE;
#pragma unroll
                for (int lvl = 0; lvl < OZ_S; ++lvl)
#pragma unroll
                    for (int a = 0; a <= lvl; ++a) {
                        const int b = lvl - a;
#pragma unroll
                        for (int k = 0; k < OZ_KB / OZ_UK; ++k)
                            oz_mma(tmem + (uint32_t)(lvl * OZ_TN), oz_desc(st + a * OZ_A_SLICE + k * OZ_UK),
                                   oz_desc(st + OZ_S * OZ_A_SLICE + b * OZ_B_SLICE + k * OZ_UK), idesc,
                                   (uint32_t)((kb | a | k) != 0));
                    }
                oz_commit(bar_empty + 8 * s);                        // the stage is free once these MMAs have read it
            }
            oz_commit(bar_tmem);                                     // all eight accumulators are final
        }
    } else {
        // epilogue: warps 2..5 own TMEM lanes 32 (warp % 4) .. + 31 = tile rows
        const int lg = warp & 3;
        const int row = ib * OZ_TM + lg * 32 + lane;
        const double rs = ldexp(1.0, g.eP[row] + g.eK);
        oz_mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {                       // 32 candidates at a time
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = OZ_S - 1; lvl >= 0; --lvl) {              // least significant level first
                uint32_t d[32];
                oz_tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(lvl * OZ_TN + half * 32), d);
                const double sc = ldexp(1.0, -7 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sc, v[j]);
            }
            // column sums over the warp's 32 rows by a transposed butterfly: lane l ends up with column l
            double q2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double x = v[j] * rs; q2[j] = x * x; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; ++j) {
                    const double keep2 = up ? q2[j + w] : q2[j], send2 = up ? q2[j] : q2[j + w];
                    q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, w);
                }
            }
            sts64(red + (uint32_t)(((lg * OZ_TN) + half * 32 + lane) * 8), q2[0]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;
        if (et < OZ_TN) {
            double s2 = 0.0;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) s2 += lds64(red + (uint32_t)((w4 * OZ_TN + et) * 8));
            g.part_ssq[(long)ib * g.ldpart + cb * OZ_TN + et] = s2;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}
