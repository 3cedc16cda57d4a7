// gpk_ozaki.cuh — the variance contraction on the int8 tensor pipe (option "ozaki" = 1).
//
// B200's fp64 ceiling is the DMMA / DFMA issue rate (37 TFLOP/s); its 5th-generation tensor cores have no fp64 kind,
// but tcgen05.mma kind::i8 (s8 x s8 -> s32, accumulators in TMEM) runs two orders of magnitude faster.  An
// error-free (Ozaki) split turns the fp64 product V = L^-1 K*^T into exact integer products:
//   P  = L^-1 :  P[i][k]  ~ 2^eP[i] sum_s Pq[s][i][k] 2^(-8 (s+1))     per-row exponent, S = 7 balanced base-256 digits
//   K*        :  K*[c][k] ~ 2^eK    sum_t Kq[t][c][k] 2^(-8 (t+1))     one exponent (0 < k <= amp)
//   V[i][c] = 2^(eP[i] + eK) sum_lvl 2^(-8 (lvl + 2)) sum_{s + t = lvl} <Pq[s][i][:], Kq[t][c][:]>      (lvl < S)
// Digits are BALANCED (-128 .. 127, oz_digits below), so every int8 carries 8 bits: 7 slices hold 56 bits of each
// operand relative to its row maximum and the triangle s + t < 7 has 28 slice pairs.  (The first version cut 7-bit
// truncated digits: 8 slices, 36 pairs, for a LARGER error — tools/ozaki_study.py prints both.)  The pairs of one
// level share one int32 accumulator ((lvl + 1) K 128^2 < 2^31 for K <= 16384), so a 128 x 64 tile keeps 7 accumulators
// of 64 columns = 448 of the 512 TMEM columns.  Per 64-byte k-block the CTA stages all 7 + 7 slice tiles (84 KB, TMA,
// 64B swizzle) once and issues 56 MMAs (128 x 64 x 32) on them.
// Measured accuracy (tools/ozaki_study.py, tools/microbench/ozaki_probe.cu): posterior variance within 4e-13 .. 4e-11
// (scaled as in the parity tests) of an 80-bit reference while max |L^-1| < 64; the handle falls back to the fp64
// DMMA kernel when the factor is worse conditioned than that (eP > OZ_MAX_EXP) or N > 16384.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocation, warps 2..5 = epilogue (TMEM ->
// fp64 with the level scales, least significant level first; row scale; column reduction sum V^2).
// Output: the per-row-block partial sums part_ssq [nb][ld] the DMMA kernel writes too.  The posterior MEAN does not go
// through the slices: mu - mean = K* alpha is one fp64 dot product of length N per candidate (gpk_rowdot_kernel on the
// fp64 K* that is built anyway, alpha = L^-T z once per fit), like george's own K* alpha; sum_i V_i z_i would put the
// slices' error in front of |z| ~ 1e2 and cost the 1e-10 tolerance on the mean (measured: 1.1e-10 .. 1.9e-10).
#pragma once
#include "gpk_gemm.cuh"

constexpr int OZ_S = 7;                       // slices (balanced base-256 digits) per operand
constexpr int OZ_PAIRS = OZ_S * (OZ_S + 1) / 2;   // slice pairs with s + t < OZ_S
constexpr int OZ_TM = 128, OZ_TN = 64;        // tile: 128 rows of L^-1 x 64 candidates
constexpr int OZ_KB = 64;                     // k-block: 64 int8 = one 64-byte swizzle row
constexpr int OZ_UK = 32;                     // K of one kind::i8 MMA
constexpr int OZ_NSTG = 2;
constexpr int OZ_A_SLICE = OZ_TM * OZ_KB, OZ_B_SLICE = OZ_TN * OZ_KB;
constexpr int OZ_STAGE = OZ_S * (OZ_A_SLICE + OZ_B_SLICE);               // 86016 bytes
constexpr int OZ_THREADS = 192;
constexpr int OZ_SMEM = OZ_NSTG * OZ_STAGE + 1024 + 256 + 4 * OZ_TN * 8;
constexpr int OZ_MAX_EXP = 7;                 // row exponents above this (|L^-1| >= 64): use the fp64 kernel

// Exponent e with |x| 2^-e inside the balanced digit interval [-128/255, 127/255) for every |x| <= amax: normally
// frexp's exponent + 1 (|x| 2^-e in [1/4, 1/2)), one more when the largest mantissa is within 0.004 of 1.
__host__ __device__ inline int oz_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int ex;
    const double m = frexp(amax, &ex);
    return ex + 1 + (m * 128.0 >= 127.49 ? 1 : 0);
}
// The OZ_S balanced base-256 digits of v (|v| < 127.49 / 256, i.e. already scaled by 2^-e), most significant first in
// the bytes 6 .. 0 of the result, each digit d as the int8 bit pattern: v ~ sum_s d_s 256^-(s+1), |error| <= 2^-57.
// Integer arithmetic: X = rint(v 2^56) (exact power-of-two scaling, one F2I), then with B = 0x80 in every byte the
// bytes b of X + B are the digits + 128 (sum (b_j - 128) 256^j = X, and 0 <= X + B < 2^56 because |X| < 0.997 2^55), so
// XOR 0x80 turns each byte into the two's-complement digit.  One conversion and three integer instructions replace
// seven rounds of scale / floor / clamp / subtract in fp64.
__device__ __forceinline__ unsigned long long oz_digits(double v) {
    const long long X = __double2ll_rn(v * 72057594037927936.0);             // 2^56
    return (unsigned long long)(X + 0x0080808080808080LL) ^ 0x0080808080808080ULL;
}
__device__ __forceinline__ int oz_digit_of(unsigned long long y, int s) { return (int)((y >> (8 * (OZ_S - 1 - s))) & 0xFFull); }

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

// The same load without the wait, and one wait for two loads: the drains of the pair kernel read two levels per round trip
// to TMEM instead of one (the wait costs a full TMEM latency and there is one epilogue warp per scheduler to hide it).
// The "+r" operands tie every destination register to the wait, so no consumer can be scheduled above it.
__device__ __forceinline__ void oz_tmem_ld32_issue(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(addr));
}
#define OZ_REGS32(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                     "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), \
                     "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), \
                     "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
__device__ __forceinline__ void oz_tmem_ld_wait(uint32_t (&a)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : OZ_REGS32(a) :: "memory");
}
__device__ __forceinline__ void oz_tmem_ld_wait(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : OZ_REGS32(a) :: "memory");
    asm volatile("" : OZ_REGS32(b) :: "memory");
}
// exact int32 -> fp64 without the quarter-rate I2F.F64: 2^52 + 2^31 + d as bit pattern, minus the constant
__device__ __forceinline__ double oz_i2d(uint32_t d) {
    return __hiloint2double(0x43300000, (int)(d ^ 0x80000000u)) - 4503601774854144.0;
}
// v[j] += scale(lvl) * accumulator(lvl)[j] for the levels hi, hi - 1, ..., lo (least significant first), two levels per wait;
// `col(lvl)` = TMEM column of level lvl's first column of this 32-column chunk
template <int HI, int LO, int LVL_COLS, int LVL_BASE>
__device__ __forceinline__ void oz_drain_levels(uint32_t lane_base, int c0, double (&v)[32]) {
    uint32_t da[32], db[32];
#pragma unroll
    for (int lvl = HI; lvl >= LO; lvl -= 2) {
        oz_tmem_ld32_issue(lane_base + (uint32_t)((lvl - LVL_BASE) * LVL_COLS + c0), da);
        if (lvl - 1 >= LO) {
            oz_tmem_ld32_issue(lane_base + (uint32_t)((lvl - 1 - LVL_BASE) * LVL_COLS + c0), db);
            oz_tmem_ld_wait(da, db);
        } else
            oz_tmem_ld_wait(da);
        const double sa = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fma(oz_i2d(da[j]), sa, v[j]);
        if (lvl - 1 >= LO) {
            const double sb = ldexp(1.0, -8 * (lvl + 1));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fma(oz_i2d(db[j]), sb, v[j]);
        }
    }
}

// ---- operand split ------------------------------------------------------------------------------------------------
// per-row exponent e[r] = oz_exponent(max |A[r][:]|); emax receives the maximum over the rows (atomicMax)
__global__ void gpk_oz_rowexp_kernel(const double* __restrict__ A, long ld, int cols, int* __restrict__ e, int* __restrict__ emax) {
    const long r = blockIdx.x;
    double m = 0.0;
    for (int c = threadIdx.x; c < cols; c += 256) m = fmax(m, fabs(A[r * ld + c]));
    __shared__ double sh[256];
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) {
        const int ex = oz_exponent(sh[0]);
        e[r] = ex;
        atomicMax(emax, ex);
    }
}
// q[s][row][col] (slices slice_stride bytes apart) = the s-th balanced base-256 digit of A[row][col] / 2^e; e = erow[row]
// or (erow == NULL) e0.  One thread per element.
__global__ void gpk_oz_split_kernel(const double* __restrict__ A, long rows, long ld, const int* __restrict__ erow, int e0,
                                    int8_t* __restrict__ q, long slice_stride) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ld) return;
    const long r = idx / ld;
    double v = ldexp(A[idx], -(erow ? erow[r] : e0));
#pragma unroll
    const unsigned long long y = oz_digits(v);
#pragma unroll
    for (int s = 0; s < OZ_S; ++s) q[(long)s * slice_stride + idx] = (int8_t)oz_digit_of(y, s);
}

// ---- the contraction ----------------------------------------------------------------------------------------------
// Tile order: candidate blocks are taken in groups of `group` blocks whose K* slices (group x TN x NP x 8 bytes, about
// 64 MB) stay L2-resident while the group walks all row blocks of L^-1, longest contraction first; without the grouping
// every row block re-reads the whole chunk's slices from HBM (measured: 9.0 GB per launch instead of ~1.1 GB).
__device__ __forceinline__ void oz_tile_of(int id, int nb, int ncb, int group, int& ib, int& cb) {
    const int full = ncb / group;
    int grp = id / (nb * group), gsz = group;
    if (grp >= full) { grp = full; gsz = ncb - full * group; }
    id -= grp * nb * group;
    ib = nb - 1 - id / gsz;
    cb = grp * group + id % gsz;
}

struct OzArgs {
    int nb, ncb;                        // row blocks of L^-1 (128 rows), candidate blocks of the chunk (64 candidates)
    int group;                          // candidate blocks per L2-resident group
    int NP, rows;                       // L^-1 is NP x NP; the K* slices have `rows` rows each
    const int* eP; int eK;
    double* part_ssq; long ldpart;
    long long* prof;                    // option "ozprof": per CTA 8 clock64() sums (gpk_oz_persist_kernel), else nullptr
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
    int ib, cb;
    oz_tile_of((int)blockIdx.x, g.nb, g.ncb, g.group, ib, cb);
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
            oz_commit(bar_tmem);                                     // all OZ_S accumulators are final
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
                const double sc = ldexp(1.0, -8 * (lvl + 2));
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

// ---------------------------------------------------------------------------------------
// CTA-pair variant (option "ozpair" = 1): tcgen05 cta_group::2.  A pair of CTAs on one TPC issues ONE MMA of M = 256
// (two row blocks of L^-1: 128 rows in each CTA's shared memory, 128 lanes in each CTA's TMEM) x 64 candidates; each CTA
// stages only HALF of the K* slice tiles (32 candidate rows) and the pair's tensor cores share them, so the operand
// fetch per CTA and MMA drops from 6 KB to 5 KB (40 clocks for 33 of arithmetic) and the L2 -> SM traffic per tile by
// 17 %; 3 stages of 70 KB fit.  The leader CTA (cluster rank 0) issues the MMAs and owns the "full" barriers: both
// CTAs' TMA loads (cp.async.bulk.tensor ... cta_group::2) complete on rank 0's barrier; "stage free" and "accumulators
// final" arrive in both CTAs through the multicast commit.  Both CTAs run the longer of the two contractions (the
// upper triangle of L^-1 is stored as zeros).  Needs an even number of row blocks.
// ---------------------------------------------------------------------------------------
constexpr int OZP_BH = OZ_TN / 2;                                   // K* rows staged per CTA
constexpr int OZP_BH_SLICE = OZP_BH * OZ_KB;                        // 2048
constexpr int OZP_STAGE = OZ_S * (OZ_A_SLICE + OZP_BH_SLICE);       // 71680
constexpr int OZP_NSTG = 3;
constexpr int OZP_SMEM = OZP_NSTG * OZP_STAGE + 1024 + 256 + 4 * OZ_TN * 8;
constexpr int OZP_SWAP = 0;                                         // which CTA of the pair stages which half (probe-checked)

__device__ __forceinline__ uint32_t oz_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t oz_map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void oz_tma_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void oz_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
gpk_oz_pair_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapKh, const OzArgs g)
{
    extern __shared__ unsigned char oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + OZP_NSTG * OZP_STAGE, bar_empty = bar_full + 8 * OZP_NSTG, bar_tmem = bar_empty + 8 * OZP_NSTG;
    const uint32_t tmem_slot = bar_tmem + 8;
    const uint32_t red = base + OZP_NSTG * OZP_STAGE + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)oz_cluster_rank();
    // Both CTAs of the pair pass the SAME shared-memory offset to the collective cta_group::2 allocation (a per-rank offset
    // was tried to silence compute-sanitizer and faults with "misaligned address"): racecheck reports the two CTAs'
    // writes of the identical TMEM base address through the pair as a hazard on this word (profiles/r02_sanitizer_racecheck.txt)
    int ibp, cb;
    oz_tile_of((int)blockIdx.x / 2, g.nb / 2, g.ncb, g.group, ibp, cb);
    const int ib = 2 * ibp + crank;
    const int nkb = (2 * ibp + 2) * OZ_TM / OZ_KB;

    if (tid == 0) {
        for (int s = 0; s < OZP_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    oz_cluster_sync();                                               // both CTAs' barriers and TMEM exist
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            const int crow = cb * OZ_TN + (crank ^ OZP_SWAP) * OZP_BH;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % OZP_NSTG;
                if (kb >= OZP_NSTG) oz_mbar_wait(bar_empty + 8 * s, (uint32_t)((kb / OZP_NSTG - 1) & 1));
                const uint32_t st = base + s * OZP_STAGE;
                const uint32_t lbar = oz_map_to_rank(bar_full + 8 * s, 0);
                if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * OZP_STAGE);       // both CTAs' bytes land here
#pragma unroll
                for (int q = 0; q < OZ_S; ++q) {
                    oz_tma_pair(st + q * OZ_A_SLICE, &mapP, kb * OZ_KB, q * g.NP + ib * OZ_TM, lbar);
                    oz_tma_pair(st + OZ_S * OZ_A_SLICE + q * OZP_BH_SLICE, &mapKh, kb * OZ_KB, q * g.rows + crow, lbar);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = oz_idesc(2 * OZ_TM, OZ_TN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % OZP_NSTG;
                oz_mbar_wait(bar_full + 8 * s, (uint32_t)((kb / OZP_NSTG) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * OZP_STAGE;
#pragma unroll
                for (int lvl = 0; lvl < OZ_S; ++lvl)
#pragma unroll
                    for (int a = 0; a <= lvl; ++a) {
                        const int b = lvl - a;
#pragma unroll
                        for (int k = 0; k < OZ_KB / OZ_UK; ++k)
                            oz_mma_pair(tmem + (uint32_t)(lvl * OZ_TN), oz_desc(st + a * OZ_A_SLICE + k * OZ_UK),
                                        oz_desc(st + OZ_S * OZ_A_SLICE + b * OZP_BH_SLICE + k * OZ_UK), idesc,
                                        (uint32_t)((kb | a | k) != 0));
                    }
                oz_commit_pair(bar_empty + 8 * s);                   // frees stage s in both CTAs
            }
            oz_commit_pair(bar_tmem);                                // accumulators final: both CTAs' epilogues go
        }
    } else {
        const int lg = warp & 3;
        const int row = ib * OZ_TM + lg * 32 + lane;
        const double rs = ldexp(1.0, g.eP[row] + g.eK);
        oz_mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = OZ_S - 1; lvl >= 0; --lvl) {
                uint32_t d[32];
                oz_tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(lvl * OZ_TN + half * 32), d);
                const double sc = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sc, v[j]);
            }
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
    oz_cluster_sync();                                               // the peer's shared memory and TMEM are no longer in use
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------
// Persistent variants (option "ozpersist" = 1, the default): one CTA (or CTA pair) per SM walks the tile list
// (tile t of unit u: t = u, u + units, ...; same L2-grouped, longest-first order), so
//   * all CTAs of the contraction are resident from the start and the block scheduler can place the covariance
//     builder of the NEXT chunk (low-priority side stream, FP64 ALU) into the SMs' spare registers / shared memory:
//     with one CTA per tile the builder only ran in the tail of the grid (measured: no overlap at all);
//   * barriers, TMEM allocation and the pipeline fill are paid once, and the TMA producer runs ahead into the next
//     tile while the epilogue warps drain the accumulators (the MMA issuer waits for "TMEM drained" per tile).
// PAIR: tcgen05 cta_group::2 as in gpk_oz_pair_kernel (M = 256, K* halves shared by the pair, 3 stages).
// ---------------------------------------------------------------------------------------
template <bool PAIR>
__global__ void __launch_bounds__(OZ_THREADS, 1)
gpk_oz_persist_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapK, const OzArgs g)
{
    constexpr int NSTG = PAIR ? OZP_NSTG : OZ_NSTG;
    constexpr int STAGE = PAIR ? OZP_STAGE : OZ_STAGE;
    constexpr int B_SLICE = PAIR ? OZP_BH_SLICE : OZ_B_SLICE;
    extern __shared__ unsigned char oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + NSTG * STAGE, bar_empty = bar_full + 8 * NSTG;
    const uint32_t bar_tfull = bar_empty + 8 * NSTG, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
    const uint32_t red = base + NSTG * STAGE + 256;                  // 2 x [4 lane groups][64 columns]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = PAIR ? (int)oz_cluster_rank() : 0;
    const int unit = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x, units = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int rows_per_tile = PAIR ? 2 : 1;
    const int nrow_tiles = g.nb / rows_per_tile, total = nrow_tiles * g.ncb;

    if (tid == 0) {
        for (int s = 0; s < NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, PAIR ? 2 : 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) oz_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            long long w_empty = 0;
            for (int t = unit; t < total; t += units) {
                int ibt, cb;
                oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
                const int ib = rows_per_tile * ibt + crank;
                const int nkb = (rows_per_tile * ibt + rows_per_tile) * OZ_TM / OZ_KB;
                const int crow = PAIR ? cb * OZ_TN + (crank ^ OZP_SWAP) * OZP_BH : cb * OZ_TN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NSTG;
                    if (it >= NSTG) {
                        const long long c0 = g.prof ? clock64() : 0;
                        oz_mbar_wait(bar_empty + 8 * s, (uint32_t)((it / NSTG - 1) & 1));
                        if (g.prof) w_empty += clock64() - c0;
                    }
                    const uint32_t st = base + s * STAGE;
                    if (PAIR) {
                        const uint32_t lbar = oz_map_to_rank(bar_full + 8 * s, 0);
                        if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * STAGE);
#pragma unroll
                        for (int q = 0; q < OZ_S; ++q) {
                            oz_tma_pair(st + q * OZ_A_SLICE, &mapP, kb * OZ_KB, q * g.NP + ib * OZ_TM, lbar);
                            oz_tma_pair(st + OZ_S * OZ_A_SLICE + q * B_SLICE, &mapK, kb * OZ_KB, q * g.rows + crow, lbar);
                        }
                    } else {
                        mbar_arrive_expect_tx(bar_full + 8 * s, STAGE);
#pragma unroll
                        for (int q = 0; q < OZ_S; ++q) {
                            tma_load_2d(st + q * OZ_A_SLICE, &mapP, kb * OZ_KB, q * g.NP + ib * OZ_TM, bar_full + 8 * s);
                            tma_load_2d(st + OZ_S * OZ_A_SLICE + q * B_SLICE, &mapK, kb * OZ_KB, q * g.rows + crow, bar_full + 8 * s);
                        }
                    }
                }
            }
            if (g.prof) g.prof[(long)blockIdx.x * 8 + 3] = w_empty;
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = oz_idesc(PAIR ? 2 * OZ_TM : OZ_TM, OZ_TN);
            int it = 0, tl = 0;
            long long w_full = 0, w_tempty = 0;
            const long long c_start = g.prof ? clock64() : 0;
            for (int t = unit; t < total; t += units, ++tl) {
                int ibt, cb;
                oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
                const int nkb = (rows_per_tile * ibt + rows_per_tile) * OZ_TM / OZ_KB;
                if (tl > 0) {                                        // the epilogue has drained the previous tile's accumulators
                    const long long c0 = g.prof ? clock64() : 0;
                    oz_mbar_wait(bar_tempty, (uint32_t)((tl - 1) & 1));
                    if (g.prof) w_tempty += clock64() - c0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NSTG;
                    {
                        const long long c0 = g.prof ? clock64() : 0;
                        oz_mbar_wait(bar_full + 8 * s, (uint32_t)((it / NSTG) & 1));
                        if (g.prof) w_full += clock64() - c0;
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * STAGE;
#pragma unroll
                    for (int lvl = 0; lvl < OZ_S; ++lvl)
#pragma unroll
                        for (int a = 0; a <= lvl; ++a) {
                            const int b = lvl - a;
#pragma unroll
                            for (int k = 0; k < OZ_KB / OZ_UK; ++k) {
                                const uint64_t da = oz_desc(st + a * OZ_A_SLICE + k * OZ_UK);
                                const uint64_t db = oz_desc(st + OZ_S * OZ_A_SLICE + b * B_SLICE + k * OZ_UK);
                                if (PAIR) oz_mma_pair(tmem + (uint32_t)(lvl * OZ_TN), da, db, idesc, (uint32_t)((kb | a | k) != 0));
                                else oz_mma(tmem + (uint32_t)(lvl * OZ_TN), da, db, idesc, (uint32_t)((kb | a | k) != 0));
                            }
                        }
                    if (PAIR) oz_commit_pair(bar_empty + 8 * s); else oz_commit(bar_empty + 8 * s);
                }
                if (PAIR) oz_commit_pair(bar_tfull); else oz_commit(bar_tfull);
            }
            if (g.prof) {
                g.prof[(long)blockIdx.x * 8 + 0] = clock64() - c_start;
                g.prof[(long)blockIdx.x * 8 + 1] = w_full;
                g.prof[(long)blockIdx.x * 8 + 2] = w_tempty;
                g.prof[(long)blockIdx.x * 8 + 6] = tl;
            }
        }
    } else {
        const int lg = warp & 3;
        int tl = 0;
        long long w_tfull = 0, w_drain = 0;
        for (int t = unit; t < total; t += units, ++tl) {
            int ibt, cb;
            oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
            const int ib = rows_per_tile * ibt + crank;
            const int row = ib * OZ_TM + lg * 32 + lane;
            const double rs = ldexp(1.0, g.eP[row] + g.eK);
            const uint32_t redt = red + (uint32_t)((tl & 1) * 4 * OZ_TN * 8);
            long long c1 = g.prof ? clock64() : 0;
            oz_mbar_wait(bar_tfull, (uint32_t)(tl & 1));
            if (g.prof) { const long long c2 = clock64(); w_tfull += c2 - c1; c1 = c2; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                double v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
                for (int lvl = OZ_S - 1; lvl >= 0; --lvl) {
                    uint32_t d[32];
                    oz_tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(lvl * OZ_TN + half * 32), d);
                    const double sc = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sc, v[j]);
                }
                if (half == 1) {                                     // TMEM is read out: hand it back before the reductions
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (tid == 64) {
                        if (PAIR) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(oz_map_to_rank(bar_tempty, 0)) : "memory");
                        else mbar_arrive(bar_tempty);
                        if (g.prof) w_drain += clock64() - c1;
                    }
                }
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
                sts64(redt + (uint32_t)(((lg * OZ_TN) + half * 32 + lane) * 8), q2[0]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int et = tid - 64;
            if (et < OZ_TN) {
                double s2 = 0.0;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) s2 += lds64(redt + (uint32_t)((w4 * OZ_TN + et) * 8));
                g.part_ssq[(long)ib * g.ldpart + cb * OZ_TN + et] = s2;
            }
        }
        if (g.prof && tid == 64) {
            g.prof[(long)blockIdx.x * 8 + 4] = w_tfull;
            g.prof[(long)blockIdx.x * 8 + 5] = w_drain;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) oz_cluster_sync();
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
    }
}
constexpr int OZ_PERSIST_SMEM = OZ_NSTG * OZ_STAGE + 1024 + 256 + 2 * 4 * OZ_TN * 8;
constexpr int OZP_PERSIST_SMEM = OZP_NSTG * OZP_STAGE + 1024 + 256 + 2 * 4 * OZ_TN * 8;

// ---------------------------------------------------------------------------------------
// Two-pass variant (option "oztile" = 128): 128 x 128 tiles, levels 0..3 in a first pass over the contraction, levels
// 4..6 in a second.  kind::i8 reads BOTH operands from shared memory at 128 B / clock / SM (measured: a 128 x 128 x 32
// MMA takes 65.9 cycles = 8 KB / 128 B), so the 128 x 64 MMAs of gpk_oz_vargemm_kernel are operand-fetch bound (6 KB ->
// 48 cycles for 33 cycles of arithmetic); 128 x 128 is balanced.  TMEM holds 4 accumulators of 128 columns per pass;
// the fp64 partial result of pass 1 waits in an L2-resident scratch tile (one per SM, indexed by %smid: 1 CTA / SM).
// 32-byte k-blocks (SWIZZLE_32B), 3 stages of 64 KB (pass 1 fills 4 + 4 slice tiles of a stage, pass 2 all 7 + 7).
// ---------------------------------------------------------------------------------------
constexpr int OZ2_KB = 32;
constexpr int OZ2_T = 128;
constexpr int OZ2_SL = OZ2_T * OZ2_KB;              // 4096 bytes per slice tile
constexpr int OZ2_STAGE = 16 * OZ2_SL;              // 65536
constexpr int OZ2_NSTG = 3;
constexpr int OZ2_SMEM = OZ2_NSTG * OZ2_STAGE + 1024 + 256 + 4 * OZ2_T * 8;
constexpr int OZ2_SCRATCH_SLOTS = 256;              // >= %nsmid

__device__ __forceinline__ uint64_t oz_desc32(uint32_t smem_addr) {       // K-major SWIZZLE_32B: 8 rows x 32 B atoms
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

struct Oz2Args {
    int nb, ncb;                        // row blocks of L^-1, candidate blocks of 128
    int group;
    int NP, rows;
    const int* eP; int eK;
    double* part_ssq; long ldpart;
    double* scratch;                    // [OZ2_SCRATCH_SLOTS][128][128]
    long long* prof;                    // option "ozprof" (gpk_oz_pair2_kernel): per CTA 8 clock64() sums, else nullptr
};

// ---------------------------------------------------------------------------------------
// CTA pair AND two passes (options "ozpair" = 1, "oztile" = 128): M = 256 x N = 128 per pair.  Measured (tools/microbench/
// ozaki_probe.cu, 16384 candidates x N = 4096): 2.60 ms against 3.35 ms for the 128 x 64 single-CTA tile.  Why: a CTA
// stages its own 128 rows of the L^-1 slices but only HALF (64 rows) of the K* slices for a 128 x 128 share of the
// output, 132 KB of L2 -> SM traffic per 64-byte k-block instead of 168 KB (one pass, N = 64) or 176 KB (two passes, one
// CTA), and the 256 x 128 x 32 MMA is not operand-fetch bound.  TMEM holds 128 columns per level, so the levels go in
// two passes over the contraction: first the three least significant (4..6, all 7 slices), drained to fp64 in an
// L2-resident scratch tile [column][row], then levels 0..3 (slices 0..3 only) on top, in the same least-significant-
// first order as the one-pass kernels (bit-identical results).  Tile walk as in gpk_oz_persist_kernel: with
// gridDim.x / 2 = number of pair tiles every pair does exactly one tile.
// ---------------------------------------------------------------------------------------
constexpr int OZQ_NT = 128;                                         // candidates per pair tile
constexpr int OZQ_BH_SLICE = (OZQ_NT / 2) * OZ_KB;                  // 4096: 64 K* rows per CTA and slice
constexpr int OZQ_STAGE = OZ_S * (OZ_A_SLICE + OZQ_BH_SLICE);       // 86016
constexpr int OZQ_NSTG = 2;
constexpr int OZQ_LOW = 4;                                          // levels OZQ_LOW .. S-1 in pass 0, 0 .. OZQ_LOW-1 in pass 1
constexpr int OZQ_SMEM = OZQ_NSTG * OZQ_STAGE + 1024 + 256 + 2 * 4 * OZQ_NT * 8;

__global__ void __launch_bounds__(OZ_THREADS, 1)
gpk_oz_pair2_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapKh, const Oz2Args g)
{
    extern __shared__ unsigned char oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + OZQ_NSTG * OZQ_STAGE, bar_empty = bar_full + 8 * OZQ_NSTG;
    const uint32_t bar_tfull = bar_empty + 8 * OZQ_NSTG, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
    const uint32_t red = base + OZQ_NSTG * OZQ_STAGE + 256;          // 2 x [4 lane groups][128 columns]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)oz_cluster_rank();
    const int unit = (int)blockIdx.x / 2, units = (int)gridDim.x / 2;
    const int nrow_tiles = g.nb / 2, total = nrow_tiles * g.ncb;

    if (tid == 0) {
        for (int s = 0; s < OZQ_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 2);                                   // one arrival per CTA of the pair (waited on by rank 0)
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    oz_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            long long w_empty = 0;
            for (int t = unit; t < total; t += units) {
                int ibt, cb;
                oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
                const int ib = 2 * ibt + crank;
                const int nkb = (2 * ibt + 2) * OZ_TM / OZ_KB;
                const int crow = cb * OZQ_NT + (crank ^ OZP_SWAP) * (OZQ_NT / 2);
                for (int pass = 0; pass < 2; ++pass) {
                    const int ns = pass == 0 ? OZ_S : OZQ_LOW;       // levels >= 4 touch every slice, levels < 4 only slices 0..3
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % OZQ_NSTG;
                        if (it >= OZQ_NSTG) {
                            const long long c0 = g.prof ? clock64() : 0;
                            oz_mbar_wait(bar_empty + 8 * s, (uint32_t)((it / OZQ_NSTG - 1) & 1));
                            if (g.prof) w_empty += clock64() - c0;
                        }
                        const uint32_t st = base + s * OZQ_STAGE;
                        const uint32_t lbar = oz_map_to_rank(bar_full + 8 * s, 0);
                        if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(2 * ns * (OZ_A_SLICE + OZQ_BH_SLICE)));
                        for (int q = 0; q < ns; ++q) {
                            oz_tma_pair(st + q * OZ_A_SLICE, &mapP, kb * OZ_KB, q * g.NP + ib * OZ_TM, lbar);
                            oz_tma_pair(st + OZ_S * OZ_A_SLICE + q * OZQ_BH_SLICE, &mapKh, kb * OZ_KB, q * g.rows + crow, lbar);
                        }
                    }
                }
            }
            if (g.prof) g.prof[(long)blockIdx.x * 8 + 3] = w_empty;
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = oz_idesc(2 * OZ_TM, OZQ_NT);
            int it = 0, n = 0;                                       // n = 2 * (tiles done) + pass
            long long w_full = 0, w_tempty = 0;
            const long long c_start = g.prof ? clock64() : 0;
            for (int t = unit; t < total; t += units) {
                int ibt, cb;
                oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
                const int nkb = (2 * ibt + 2) * OZ_TM / OZ_KB;
                for (int pass = 0; pass < 2; ++pass, ++n) {
                    if (n > 0) {                                     // both CTAs' epilogues have drained the previous accumulators
                        const long long c0 = g.prof ? clock64() : 0;
                        oz_mbar_wait(bar_tempty, (uint32_t)((n - 1) & 1));
                        if (g.prof) w_tempty += clock64() - c0;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % OZQ_NSTG;
                        {
                            const long long c0 = g.prof ? clock64() : 0;
                            oz_mbar_wait(bar_full + 8 * s, (uint32_t)((it / OZQ_NSTG) & 1));
                            if (g.prof) w_full += clock64() - c0;
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = base + s * OZQ_STAGE;
                        if (pass == 0) {
#pragma unroll
                            for (int lvl = OZQ_LOW; lvl < OZ_S; ++lvl)
#pragma unroll
                                for (int a = 0; a <= lvl; ++a)
#pragma unroll
                                    for (int k = 0; k < OZ_KB / OZ_UK; ++k)
                                        oz_mma_pair(tmem + (uint32_t)((lvl - OZQ_LOW) * OZQ_NT), oz_desc(st + a * OZ_A_SLICE + k * OZ_UK),
                                                    oz_desc(st + OZ_S * OZ_A_SLICE + (lvl - a) * OZQ_BH_SLICE + k * OZ_UK), idesc,
                                                    (uint32_t)((kb | a | k) != 0));
                        } else {
#pragma unroll
                            for (int lvl = 0; lvl < OZQ_LOW; ++lvl)
#pragma unroll
                                for (int a = 0; a <= lvl; ++a)
#pragma unroll
                                    for (int k = 0; k < OZ_KB / OZ_UK; ++k)
                                        oz_mma_pair(tmem + (uint32_t)(lvl * OZQ_NT), oz_desc(st + a * OZ_A_SLICE + k * OZ_UK),
                                                    oz_desc(st + OZ_S * OZ_A_SLICE + (lvl - a) * OZQ_BH_SLICE + k * OZ_UK), idesc,
                                                    (uint32_t)((kb | a | k) != 0));
                        }
                        oz_commit_pair(bar_empty + 8 * s);
                    }
                    oz_commit_pair(bar_tfull);
                }
            }
            if (g.prof) {
                g.prof[(long)blockIdx.x * 8 + 0] = clock64() - c_start;
                g.prof[(long)blockIdx.x * 8 + 1] = w_full;
                g.prof[(long)blockIdx.x * 8 + 2] = w_tempty;
                g.prof[(long)blockIdx.x * 8 + 6] = n / 2;
            }
        }
    } else {
        const int lg = warp & 3;
        const int rl = lg * 32 + lane;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        double* sc = g.scratch + (size_t)(smid % OZ2_SCRATCH_SLOTS) * OZ_TM * OZQ_NT;      // [column][row]: lanes = rows
        const uint32_t lane_base = tmem + ((uint32_t)(lg * 32) << 16);
        const uint32_t tempty_leader = oz_map_to_rank(bar_tempty, 0);
        int n = 0, tl = 0;
        for (int t = unit; t < total; t += units, ++tl) {
            int ibt, cb;
            oz_tile_of(t, nrow_tiles, g.ncb, g.group, ibt, cb);
            const int ib = 2 * ibt + crank;
            const int row = ib * OZ_TM + rl;
            const double rs = ldexp(1.0, g.eP[row] + g.eK);
            const uint32_t redt = red + (uint32_t)((tl & 1) * 4 * OZQ_NT * 8);
            // pass 0: levels S-1 .. OZQ_LOW -> fp64 partial in the scratch tile
            oz_mbar_wait(bar_tfull, (uint32_t)(n & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < OZQ_NT; c0 += 32) {
                double v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.0;
                oz_drain_levels<OZ_S - 1, OZQ_LOW, OZQ_NT, OZQ_LOW>(lane_base, c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sc[(size_t)(c0 + j) * OZ_TM + rl] = v[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tid == 64) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(tempty_leader) : "memory");
            ++n;
            // pass 1: levels OZQ_LOW-1 .. 0 on top, squares, column sums
            oz_mbar_wait(bar_tfull, (uint32_t)(n & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < OZQ_NT; c0 += 32) {
                double v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = sc[(size_t)(c0 + j) * OZ_TM + rl];
                oz_drain_levels<OZQ_LOW - 1, 0, OZQ_NT, 0>(lane_base, c0, v);
                if (c0 == OZQ_NT - 32) {                             // TMEM is read out: hand it back before the reductions
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (tid == 64) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(tempty_leader) : "memory");
                }
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
                sts64(redt + (uint32_t)((lg * OZQ_NT + c0 + lane) * 8), q2[0]);
            }
            ++n;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int et = tid - 64;                                 // 0 .. 127 = column of the tile
            double s2 = 0.0;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) s2 += lds64(redt + (uint32_t)((w4 * OZQ_NT + et) * 8));
            g.part_ssq[(long)ib * g.ldpart + cb * OZQ_NT + et] = s2;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    oz_cluster_sync();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
gpk_oz2_vargemm_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapK, const Oz2Args g)
{
    extern __shared__ unsigned char oz_raw[];
    const uint32_t base = (smem_u32(oz_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + OZ2_NSTG * OZ2_STAGE, bar_empty = bar_full + 8 * OZ2_NSTG;
    const uint32_t bar_tfull = bar_empty + 8 * OZ2_NSTG, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
    const uint32_t red = base + OZ2_NSTG * OZ2_STAGE + 256;          // [4 lane groups][128 columns]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ib, cb;
    oz_tile_of((int)blockIdx.x, g.nb, g.ncb, g.group, ib, cb);
    const int nkb = (ib + 1) * OZ2_T / OZ2_KB;

    if (tid == 0) {
        for (int s = 0; s < OZ2_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 1);
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
            int it = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const int ns = pass == 0 ? 4 : OZ_S;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % OZ2_NSTG;
                    if (it >= OZ2_NSTG) oz_mbar_wait(bar_empty + 8 * s, (uint32_t)((it / OZ2_NSTG - 1) & 1));
                    const uint32_t st = base + s * OZ2_STAGE;
                    mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(2 * ns * OZ2_SL));
                    for (int q = 0; q < ns; ++q) {
                        tma_load_2d(st + q * OZ2_SL, &mapP, kb * OZ2_KB, q * g.NP + ib * OZ2_T, bar_full + 8 * s);
                        tma_load_2d(st + (8 + q) * OZ2_SL, &mapK, kb * OZ2_KB, q * g.rows + cb * OZ2_T, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = oz_idesc(OZ2_T, OZ2_T);
            int it = 0;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {                                      // the epilogue has drained the first four accumulators
                    oz_mbar_wait(bar_tempty, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % OZ2_NSTG;
                    oz_mbar_wait(bar_full + 8 * s, (uint32_t)((it / OZ2_NSTG) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * OZ2_STAGE;
                    if (pass == 0) {
#pragma unroll
                        for (int lvl = 0; lvl < 4; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
                                oz_mma(tmem + (uint32_t)(lvl * OZ2_T), oz_desc32(st + a * OZ2_SL), oz_desc32(st + (8 + lvl - a) * OZ2_SL),
                                       idesc, (uint32_t)((kb | a) != 0));
                    } else {
#pragma unroll
                        for (int lvl = 4; lvl < OZ_S; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
                                oz_mma(tmem + (uint32_t)((lvl - 4) * OZ2_T), oz_desc32(st + a * OZ2_SL),
                                       oz_desc32(st + (8 + lvl - a) * OZ2_SL), idesc, (uint32_t)((kb | a) != 0));
                    }
                    oz_commit(bar_empty + 8 * s);
                }
                oz_commit(bar_tfull);
            }
        }
    } else {
        const int lg = warp & 3;
        const int rl = lg * 32 + lane, row = ib * OZ2_T + rl;
        const double rs = ldexp(1.0, g.eP[row] + g.eK);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        double* sc = g.scratch + ((size_t)(smid % OZ2_SCRATCH_SLOTS) * OZ2_T + rl) * OZ2_T;
        const uint32_t lane_base = tmem + ((uint32_t)(lg * 32) << 16);
        // pass 1 result (levels 0..3) -> scratch
        oz_mbar_wait(bar_tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < OZ2_T; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = 3; lvl >= 0; --lvl) {
                uint32_t d[32];
                oz_tmem_ld32(lane_base + (uint32_t)(lvl * OZ2_T + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(sc + c0 + j) = make_double2(v[j], v[j + 1]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 64) mbar_arrive(bar_tempty);
        // pass 2 result (levels 4..6) + scratch -> squares -> column sums
        oz_mbar_wait(bar_tfull, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < OZ2_T; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = OZ_S - 1; lvl >= 4; --lvl) {
                uint32_t d[32];
                oz_tmem_ld32(lane_base + (uint32_t)((lvl - 4) * OZ2_T + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
            double q2[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 hi = *reinterpret_cast<const double2*>(sc + c0 + j);
                const double x0 = (v[j] + hi.x) * rs, x1 = (v[j + 1] + hi.y) * rs;
                q2[j] = x0 * x0;
                q2[j + 1] = x1 * x1;
            }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; ++j) {
                    const double keep2 = up ? q2[j + w] : q2[j], send2 = up ? q2[j] : q2[j + w];
                    q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, w);
                }
            }
            sts64(red + (uint32_t)((lg * OZ2_T + c0 + lane) * 8), q2[0]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;
        double s2 = 0.0;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) s2 += lds64(red + (uint32_t)((w4 * OZ2_T + et) * 8));
        g.part_ssq[(long)ib * g.ldpart + cb * OZ2_T + et] = s2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------
// Covariance builder for the int8 path: gpk_cov_tma_kernel's tile loop (TMA-staged pre-scaled train operand, thread =
// 2 train points x CC candidates), but the fp64 K* never reaches HBM: every value k = amp * prod f(q) leaves as its
// OZ_S balanced base-256 digits (Kq[s][cand][j], two adjacent int8 per thread and slice), and the posterior mean's share
// sum_j k(c, j) alpha_j of this 128-column tile is reduced over the 64 threads of a candidate group and written to
// part_mu[tile][cand] (summed in fixed order by gpk_finish_kernel: deterministic).  Replaces K* store (8 B / element)
// + split kernel (8 B read, 8 B written) + mean dot (8 B read) by 8 B written per element.
// ---------------------------------------------------------------------------------------
template <int CC>
__global__ void __launch_bounds__(256, CC == 8 ? 2 : 4)
gpk_cov_oz_kernel(const __grid_constant__ CUtensorMap mapX, const KSpec ks, int n,
                  const double* __restrict__ cand, int dc, long m,
                  const double* __restrict__ lower, const double* __restrict__ upper,
                  const double* __restrict__ alpha, int eK, int8_t* __restrict__ Kq, long ldq, long slice_stride,
                  double* __restrict__ part_mu, long ldpart, int gx, int gy, int trigger)
{
    // trigger: this grid is the PRIMARY of a programmatic dependent launch (score_dev): all its CTAs are resident, so the
    // int8 contraction of the previous chunk, launched right behind it on the same stream, may start now and run beside it
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Work items (train tile bx of 128 columns, candidate group by of 4 CC rows) are walked with stride gridDim.x: a grid of
    // gx * gy CTAs does one item each; the look-ahead launch uses one or two CTAs per SM (see score_dev: they are
    // resident before the contraction's CTAs arrive, which is what lets the two kernels overlap).
    constexpr int TC = 4 * CC;
    extern __shared__ unsigned char cov_raw[];
    const int tid = threadIdx.x;
    const int nt = ks.n_terms;
    const uint32_t base = (cov_smem_u32(cov_raw) + 127u) & ~127u;
    const uint32_t xs = base;                                   // nt x 128 doubles
    const uint32_t scb = xs + (uint32_t)nt * 1024u;             // TC x nt doubles
    const uint32_t bar = scb + (uint32_t)(TC * nt) * 8u;
    const uint32_t mred = bar + 8u;                             // 8 warps x CC doubles: cross-warp mean reduction
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int jp = tid & 63, cgp = tid >> 6, lane = tid & 31;
    const double sc = ldexp(1.0, -eK);
    uint32_t phase = 0;
    for (long w = blockIdx.x; w < (long)gx * gy; w += gridDim.x, phase ^= 1u) {
        const int bx = (int)(w % gx);
        const long c0 = (w / gx) * TC;
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)nt * 1024u) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(xs), "l"((uint64_t)&mapX), "r"(bar), "r"(bx * 128), "r"(0) : "memory");
        }
        for (int e = tid; e < TC * nt; e += 256) {
            const int c = e / nt, t = e - c * nt;
            const long ci = c0 + c;
            double v = 0.0;
            if (ci < m) {
                const int a = ks.axis[t];
                v = cand[ci * dc + a];
                if (lower != nullptr) v = (v - lower[a]) / (upper[a] - lower[a]);
                v *= ks.scale[t];
            }
            asm volatile("st.shared.f64 [%0], %1;" :: "r"(scb + (uint32_t)e * 8u), "d"(v) : "memory");
        }
        __syncthreads();
        {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        }
        double q[CC][2], pr[CC][2];
#pragma unroll
        for (int c = 0; c < CC; ++c) { q[c][0] = q[c][1] = 0.0; pr[c][0] = pr[c][1] = 1.0; }
        const uint32_t xrow = xs + (uint32_t)jp * 16u;
        const uint32_t srow = scb + (uint32_t)(cgp * CC * nt) * 8u;
        for (int t = 0; t < nt; ++t) {
            double x0, x1;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(xrow + (uint32_t)t * 1024u));
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                double sv;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(sv) : "r"(srow + (uint32_t)(c * nt + t) * 8u));
                const double d0 = sv - x0, d1 = sv - x1;
                q[c][0] = fma(d0, d0, q[c][0]);
                q[c][1] = fma(d1, d1, q[c][1]);
            }
            if (ks.last[t]) {
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    pr[c][0] *= gpk_radial_q(ks.family, q[c][0]);
                    pr[c][1] *= gpk_radial_q(ks.family, q[c][1]);
                    q[c][0] = q[c][1] = 0.0;
                }
            }
        }
        const int j0 = bx * 128 + 2 * jp;
        const bool v0 = j0 < n, v1 = j0 + 1 < n;
        const double a0 = v0 ? alpha[j0] : 0.0, a1 = v1 ? alpha[j0 + 1] : 0.0;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            const long ci = c0 + cgp * CC + c;
            const bool cv = ci < m;
            const double k0 = (cv && v0) ? ks.amp * pr[c][0] : 0.0;
            const double k1 = (cv && v1) ? ks.amp * pr[c][1] : 0.0;
            // digits: two adjacent int8 per slice
            const unsigned long long y0 = oz_digits(k0 * sc), y1 = oz_digits(k1 * sc);
            int8_t* dst = Kq + ci * ldq + j0;
#pragma unroll
            for (int s2 = 0; s2 < OZ_S; ++s2)
                *reinterpret_cast<uint16_t*>(dst + (long)s2 * slice_stride) = (uint16_t)(oz_digit_of(y0, s2) | (oz_digit_of(y1, s2) << 8));
            // mean share of this tile: reduce over the 64 threads (2 warps) of the candidate group, fixed order
            double pm = fma(k0, a0, k1 * a1);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) pm += __shfl_xor_sync(0xffffffffu, pm, off);
            if (lane == 0) asm volatile("st.shared.f64 [%0], %1;" :: "r"(mred + (uint32_t)(((tid >> 5) * CC + c) * 8)), "d"(pm) : "memory");
        }
        __syncthreads();
        if (tid < TC) {
            const int g = tid / CC, c = tid - g * CC;           // candidate group g = warps 2g, 2g + 1
            double lo2, hi2;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lo2) : "r"(mred + (uint32_t)(((2 * g) * CC + c) * 8)));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(hi2) : "r"(mred + (uint32_t)(((2 * g + 1) * CC + c) * 8)));
            part_mu[(long)bx * ldpart + c0 + tid] = lo2 + hi2;
        }
        __syncthreads();                                        // the next item overwrites the operand tiles and mred
    }
}
inline size_t cov_oz_smem_bytes(int n_terms, int cc) { return (size_t)n_terms * 1024 + (size_t)4 * cc * n_terms * 8 + 8 + 8 * cc * 8 + 128; }

// ---------------------------------------------------------------------------------------
// int8 tensor-pipe issue-rate peak of this GPU (roofline denominator of gpk_oz_vargemm_kernel in bench.py): every SM
// issues `iters` kind::i8 MMAs of 128 x 128 x 32 on one shared-memory operand pair (contents irrelevant) into one
// TMEM accumulator; no loads in the loop.  tools/microbench/i8_umma_probe.cu checks the same instruction against a
// CPU integer GEMM.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) gpk_peak_i8_kernel(int iters, int random_operands)
{
    extern __shared__ unsigned char pk_raw[];
    const uint32_t base = (smem_u32(pk_raw) + 1023u) & ~1023u;            // A: 128 x 64 B, B: 128 x 64 B (64B swizzle atoms)
    const uint32_t bar = base + 2 * 8192, slot = bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    // operands: a constant pattern (no switching activity in the datapath) or pseudo-random bytes (the statistics of real
    // digit slices: this is what decides how far the power limit lets the pipe run in a long measurement)
    for (int e = tid; e < 4096; e += 128) {
        uint32_t w = 0x01010101u * (e & 3);
        if (random_operands) { w = (uint32_t)e * 2654435761u + blockIdx.x * 40503u; w ^= w >> 15; w *= 2246822519u; w ^= w >> 13; }
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + 4u * e), "r"(w) : "memory");
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
    if (tid == 0) {
        const uint32_t idesc = oz_idesc(128, 128);
        for (int i = 0; i < iters; ++i) {
            oz_mma(tmem, oz_desc(base), oz_desc(base + 8192), idesc, (uint32_t)(i != 0));
            oz_mma(tmem, oz_desc(base + 32), oz_desc(base + 8192 + 32), idesc, 1u);
        }
        oz_commit(bar);
        oz_mbar_wait(bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}
