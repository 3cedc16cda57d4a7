// ozaki_probe.cu — stand-alone prototype for VERDICT item 8: the variance contraction
//     V = L^-1 K*^T ,  ssq_c = sum_i V_ic^2 ,  mu_c = sum_i V_ic z_i
// on the int8 tensor pipe (tcgen05.mma kind::i8, TMEM accumulators) through an error-free (Ozaki) split of both
// fp64 operands into S = 7 balanced base-256 digits (-128 .. 127), instead of fp64 DMMA.
//
//   P  (N x N, lower triangular, fp64)  ->  Pq[s][i][k] int8,  P[i][k]  ~ 2^eP[i] sum_s Pq[s][i][k] 2^(-8 (s+1))
//   K* (M x N, fp64, 0 < k <= amp)      ->  Kq[t][c][k] int8,  K*[c][k] ~ 2^eK    sum_t Kq[t][c][k] 2^(-8 (t+1))
//   V[i][c] = 2^(eP[i] + eK) sum_lvl 2^(-8 (lvl + 2)) sum_{s + t = lvl} <Pq[s][i][:], Kq[t][c][:]>      (lvl < S)
// Every slice-pair product is an exact int32 GEMM; the pairs of one level share one TMEM accumulator
// ((lvl + 1) * K * 128^2 < 2^31 for K <= 16384, S = 7), so a 128 x 64 tile keeps S = 7 accumulators of 64 columns = 448
// of the 512 TMEM columns.  Per 64-byte k-block the CTA stages all 7 + 7 slices (84 KB) once and issues 28 pairs x 2 MMAs.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue (TMEM -> fp64, scales,
// column reductions).  Checked against an 80-bit CPU reference on real GP data (Matern-5/2, N = 1024) and timed on a
// C2-sized synthetic problem (N = 4096, 16384 candidates).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ozaki_probe.bin ozaki_probe.cu -lcuda && ./ozaki_probe.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { printf("{\"error\": \"%s -> %s (line %d)\"}\n", #call, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int S = 7;                    // slices per operand
constexpr int TM = 128, TN = 64;        // tile: 128 rows of P x 64 candidates
constexpr int KBY = 64;                 // k-block in bytes (= int8 elements): one 64B-swizzle atom row
constexpr int UMMA_K = 32;
constexpr int NSTG = 2;
constexpr int A_SLICE = TM * KBY, B_SLICE = TN * KBY;                   // 8192, 4096 bytes
constexpr int STAGE = S * (A_SLICE + B_SLICE);                          // 98304 bytes
constexpr int OZ_THREADS = 192;                                          // 6 warps
constexpr int OZ_SMEM = NSTG * STAGE + 1024 + 256 + 4 * TN * 2 * 8;

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
__device__ __forceinline__ void tma_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// K-major SWIZZLE_64B matrix descriptor: atoms of 8 rows x 64 bytes (512 B), stride byte offset 512, version 1, layout 4
__device__ __forceinline__ uint64_t umma_desc64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
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
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
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

// ---------------------------------------------------------------------------------------
// operand split: one thread per element.  mode 0: per-row exponent from e[row]; mode 1: one exponent e[0] for all rows.
// q[s][row][col] (ld = cols), balanced base-256 digits.
// ---------------------------------------------------------------------------------------
// exponent e with |x| 2^-e in the balanced digit interval [-128/255, 127/255) for all |x| <= amax
__host__ __device__ inline int oz_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int ex;
    const double m = frexp(amax, &ex);
    return ex + 1 + (m * 128.0 >= 127.49 ? 1 : 0);
}
__global__ void oz_rowmax_kernel(const double* __restrict__ A, long rows, long cols, int* __restrict__ e) {
    const long r = blockIdx.x;
    double m = 0.0;
    for (long c = threadIdx.x; c < cols; c += blockDim.x) m = fmax(m, fabs(A[r * cols + c]));
    __shared__ double sh[256];
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) e[r] = oz_exponent(sh[0]);
}
__global__ void oz_split_kernel(const double* __restrict__ A, long rows, long cols, const int* __restrict__ e, int mode,
                                int8_t* __restrict__ q) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const long r = idx / cols;
    double v = ldexp(A[idx], -(mode == 0 ? e[r] : e[0]));
#pragma unroll
    // balanced base-256 digits by integer arithmetic: X = rint(v 2^56); the bytes of X + 0x80..80 are the digits + 128
    const long long X = __double2ll_rn(v * 72057594037927936.0);
    const unsigned long long y = (unsigned long long)(X + 0x0080808080808080LL) ^ 0x0080808080808080ULL;
#pragma unroll
    for (int s = 0; s < S; ++s) q[(long)s * rows * cols + idx] = (int8_t)((y >> (8 * (S - 1 - s))) & 0xFFull);
}

// ---------------------------------------------------------------------------------------
// the contraction: one CTA per tile (row block ib, 64-candidate block cb), longest contractions first
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void oz_tile_of(int id, int nb, int ncb, int group, int& ib, int& cb) {
    const int full = ncb / group;
    int grp = id / (nb * group), gsz = group;
    if (grp >= full) { grp = full; gsz = ncb - full * group; }
    id -= grp * nb * group;
    ib = nb - 1 - id / gsz;
    cb = grp * group + id % gsz;
}

struct OzArgs {
    int group;                          // candidate blocks per L2-resident group (0: plain order)
    int nb, ncb;                        // row blocks of P (128 rows), candidate blocks (64)
    int N, Mc;                          // P is N x N, K* is Mc x N
    const int* eP; const int* eK;       // exponents
    const double* z;
    double* part_ssq; double* part_mu;  // [nb][Mc]
};

// CL > 1: CL CTAs of a cluster work on the same row block and CL neighbouring candidate blocks; every CTA loads
// S / CL of the L^-1 slice tiles of a k-block and multicasts them to the whole cluster (its K* slices stay private),
// so the L^-1 traffic per CTA drops by CL.  A stage may be refilled once ALL CTAs of the cluster have consumed it:
// the MMA issuers commit to the empty barrier of every CTA (count CL).
template <int CL, int ORDER>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_vargemm_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapK, const OzArgs g)
{
    extern __shared__ unsigned char raw[];
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + NSTG * STAGE, bar_empty = bar_full + 8 * NSTG, bar_tmem = bar_empty + 8 * NSTG;
    const uint32_t tmem_slot = bar_tmem + 8;
    const uint32_t red = base + NSTG * STAGE + 256;                  // [4 warps][64 cols][2] doubles
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = CL > 1 ? (int)cluster_rank() : 0;
    const int cid = (int)blockIdx.x / CL, cpr = g.ncb / CL;          // cluster id, clusters per row block
    int ib, cb;
    if (g.group > 0) { int cc; oz_tile_of(cid, g.nb, cpr, g.group / CL, ib, cc); cb = cc * CL + crank; }   // groups of clusters
    else { ib = g.nb - 1 - cid / cpr; cb = (cid % cpr) * CL + crank; }
    const int nkb = (ib + 1) * TM / KBY;                             // lower triangle: columns < (ib + 1) * 128
    const uint16_t cmask = (uint16_t)((1u << CL) - 1);

    if (tid == 0) {
        for (int s = 0; s < NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CL); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                  // every CTA's barriers exist before anyone signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTG;
                if (kb >= NSTG) mbar_wait(bar_empty + 8 * s, (uint32_t)((kb / NSTG - 1) & 1));
                const uint32_t st = base + s * STAGE;
                mbar_expect_tx(bar_full + 8 * s, STAGE);
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    if (CL == 1) tma_2d(st + q * A_SLICE, &mapP, kb * KBY, q * g.N + ib * TM, bar_full + 8 * s);
                    else if (q % CL == crank) tma_2d_mc(st + q * A_SLICE, &mapP, kb * KBY, q * g.N + ib * TM, bar_full + 8 * s, cmask);
                    tma_2d(st + S * A_SLICE + q * B_SLICE, &mapK, kb * KBY, q * g.Mc + cb * TN, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(TM, TN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTG;
                mbar_wait(bar_full + 8 * s, (uint32_t)((kb / NSTG) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE;
                if (ORDER == 0) {
#pragma unroll
                for (int lvl = 0; lvl < S; ++lvl)
#pragma unroll
                    for (int a = 0; a <= lvl; ++a) {
                        const int b = lvl - a;
#pragma unroll
                        for (int k = 0; k < KBY / UMMA_K; ++k)
                            umma_i8(tmem + (uint32_t)(lvl * TN), umma_desc64(st + a * A_SLICE + k * UMMA_K),
                                    umma_desc64(st + S * A_SLICE + b * B_SLICE + k * UMMA_K), idesc,
                                    (uint32_t)((kb | a | k) != 0));
                    }
                } else {
                    // same products, consecutive instructions share the (128 x 32) slice tile of L^-1: k outer, a, then b
#pragma unroll
                for (int k = 0; k < KBY / UMMA_K; ++k)
#pragma unroll
                    for (int a = 0; a < S; ++a)
#pragma unroll
                        for (int b = 0; b < S - a; ++b)
                            umma_i8(tmem + (uint32_t)((a + b) * TN), umma_desc64(st + a * A_SLICE + k * UMMA_K),
                                    umma_desc64(st + S * A_SLICE + b * B_SLICE + k * UMMA_K), idesc,
                                    (uint32_t)((kb | a | k) != 0));
                }
                if (CL == 1) umma_commit(bar_empty + 8 * s);         // frees the stage when these MMAs have read it
                else umma_commit_mc(bar_empty + 8 * s, cmask);       // ... in every CTA of the cluster
            }
            umma_commit(bar_tmem);                                   // all accumulators final
        }
    } else {
        // ---------------- epilogue: warps 2..5 own TMEM lanes 32 (warp % 4) .. + 31 = tile rows ----------------
        const int lg = warp & 3;
        const int row = ib * TM + lg * 32 + lane;
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const double zr = g.z[row];
        const double rs = ldexp(1.0, g.eP[row] + g.eK[0]);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {                      // 32 candidates at a time (registers)
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = S - 1; lvl >= 0; --lvl) {                 // least significant level first
                uint32_t d[32];
                tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(lvl * TN + half * 32), d);
                const double sc = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sc, v[j]);
            }
            // column sums over the warp's 32 rows: transposed butterfly, lane l ends with column l
            double q2[32], qm[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double x = v[j] * rs; q2[j] = x * x; qm[j] = x * zr; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; ++j) {
                    // keep the half of the columns selected by this lane bit, send the other half to the partner
                    const double keep2 = up ? q2[j + w] : q2[j], send2 = up ? q2[j] : q2[j + w];
                    const double keepm = up ? qm[j + w] : qm[j], sendm = up ? qm[j] : qm[j + w];
                    q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, w);
                    qm[j] = keepm + __shfl_xor_sync(0xffffffffu, sendm, w);
                }
            }
            // lane l now holds column (bit-reversal-free mapping): after the steps w = 16 .. 1 the surviving column is
            // c = sum over bits of (lane & w) -> exactly `lane`
            const uint32_t slot = red + (uint32_t)(((lg * TN) + half * 32 + lane) * 16);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(slot), "d"(q2[0]), "d"(qm[0]) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;                                     // 0 .. 127
        if (et < TN) {
            double s2 = 0.0, sm = 0.0;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
                double a, b;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(red + (uint32_t)((w4 * TN + et) * 16)));
                s2 += a; sm += b;
            }
            g.part_ssq[(long)ib * g.Mc + cb * TN + et] = s2;
            g.part_mu[(long)ib * g.Mc + cb * TN + et] = sm;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                  // no CTA leaves while peers may still signal / write it
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}


// =======================================================================================
// v7: CTA PAIR (tcgen05 cta_group::2).  kind::i8 reads both operands from shared memory (128 B / clock / SM), so the
// 128 x 64 MMAs above are operand-fetch bound (6 KB -> 48 clocks for 33 clocks of arithmetic).  A pair of CTAs on one
// TPC issues ONE MMA of M = 256 (two row blocks of L^-1, 128 rows in each CTA's shared memory and TMEM) x N candidates;
// each CTA stages only HALF of the K* slice tiles (N / 2 candidate rows) and the pair's tensor cores share them:
// 4 KB + N/2 x 32 B per CTA and MMA.  The leader CTA (rank 0) issues the MMAs and owns the "full" barriers (both CTAs'
// TMA loads signal them: cp.async.bulk.tensor ... cta_group::2 with the barrier address mapped to rank 0); "empty" and
// "accumulators final" arrive in both CTAs through the multicast commit.  `swap` selects which CTA stages which half
// of the candidate rows (checked at run time against the single-CTA kernel).
// =======================================================================================
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

template <int NT>                                   // candidates per pair tile
struct PairCfg {
    static constexpr int BH = NT / 2;               // K* rows staged per CTA
    static constexpr int BH_SLICE = BH * KBY;
    static constexpr int STAGE = S * (A_SLICE + BH_SLICE);
    static constexpr int NSTG = (3 * STAGE + 8192 <= 227 * 1024) ? 3 : 2;
    static constexpr int SMEM = NSTG * STAGE + 1024 + 256 + 4 * NT * 2 * 8;
};

template <int NT>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_pair_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapKh, const OzArgs g, const int swap)
{
    using C = PairCfg<NT>;
    extern __shared__ unsigned char raw[];
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + C::NSTG * C::STAGE, bar_empty = bar_full + 8 * C::NSTG, bar_tmem = bar_empty + 8 * C::NSTG;
    const uint32_t tmem_slot = bar_tmem + 8;
    const uint32_t red = base + C::NSTG * C::STAGE + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)cluster_rank();
    int ibp, cb;
    oz_tile_of((int)blockIdx.x / 2, g.nb / 2, g.ncb, g.group, ibp, cb);
    const int ib = 2 * ibp + crank;
    const int nkb = (2 * ibp + 2) * TM / KBY;                        // both CTAs run the longer of the two contractions
                                                                     // (the upper triangle of L^-1 is stored as zeros)
    if (tid == 0) {
        for (int s = 0; s < C::NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                              // both CTAs' barriers and TMEM exist
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            const int crow = cb * NT + (crank ^ swap) * C::BH;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % C::NSTG;
                if (kb >= C::NSTG) mbar_wait(bar_empty + 8 * s, (uint32_t)((kb / C::NSTG - 1) & 1));
                const uint32_t st = base + s * C::STAGE;
                const uint32_t lbar = map_to_rank(bar_full + 8 * s, 0);
                if (crank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * C::STAGE);      // both CTAs' bytes land on this barrier
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    tma_2d_pair(st + q * A_SLICE, &mapP, kb * KBY, q * g.N + ib * TM, lbar);
                    tma_2d_pair(st + S * A_SLICE + q * C::BH_SLICE, &mapKh, kb * KBY, q * g.Mc + crow, lbar);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = umma_idesc(2 * TM, NT);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % C::NSTG;
                mbar_wait(bar_full + 8 * s, (uint32_t)((kb / C::NSTG) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * C::STAGE;
#pragma unroll
                for (int lvl = 0; lvl < S; ++lvl)
#pragma unroll
                    for (int a = 0; a <= lvl; ++a) {
                        const int b = lvl - a;
#pragma unroll
                        for (int k = 0; k < KBY / UMMA_K; ++k)
                            umma_i8_pair(tmem + (uint32_t)(lvl * NT), umma_desc64(st + a * A_SLICE + k * UMMA_K),
                                         umma_desc64(st + S * A_SLICE + b * C::BH_SLICE + k * UMMA_K), idesc,
                                         (uint32_t)((kb | a | k) != 0));
                    }
                umma_commit_pair(bar_empty + 8 * s);                 // frees stage s in BOTH CTAs
            }
            umma_commit_pair(bar_tmem);                              // accumulators final, both CTAs' epilogues go
        }
    } else {
        const int lg = warp & 3;
        const int row = ib * TM + lg * 32 + lane;
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const double zr = g.z[row];
        const double rs = ldexp(1.0, g.eP[row] + g.eK[0]);
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = S - 1; lvl >= 0; --lvl) {
                uint32_t d[32];
                tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(lvl * NT + c0), d);
                const double sc = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sc, v[j]);
            }
            double q2[32], qm[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double x = v[j] * rs; q2[j] = x * x; qm[j] = x * zr; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; ++j) {
                    const double keep2 = up ? q2[j + w] : q2[j], send2 = up ? q2[j] : q2[j + w];
                    const double keepm = up ? qm[j + w] : qm[j], sendm = up ? qm[j] : qm[j + w];
                    q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, w);
                    qm[j] = keepm + __shfl_xor_sync(0xffffffffu, sendm, w);
                }
            }
            const uint32_t slot = red + (uint32_t)(((lg * NT) + c0 + lane) * 16);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(slot), "d"(q2[0]), "d"(qm[0]) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;
        if (et < NT) {
            double s2 = 0.0, sm = 0.0;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
                double a, b;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(red + (uint32_t)((w4 * NT + et) * 16)));
                s2 += a; sm += b;
            }
            g.part_ssq[(long)ib * g.Mc + cb * NT + et] = s2;
            g.part_mu[(long)ib * g.Mc + cb * NT + et] = sm;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                              // the peer's shared memory / TMEM are no longer in use
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}


// =======================================================================================
// v8: CTA pair AND two passes: M = 256 x N = 128 per pair.  Per CTA and MMA 4 KB (own 128 rows of an L^-1 slice) + 2 KB
// (half of the 128 K* rows) of operand fetch for 66 clocks of arithmetic: tensor-bound, where the single-CTA 128 x 64
// tile is fetch-bound (48 clocks for 33).  TMEM holds 128 columns per level, so the levels go in two passes over the
// contraction: first the three least significant (4..6, all 7 slices), drained to fp64 in an L2-resident scratch tile,
// then levels 0..3 (slices 0..3) added on top in the same least-significant-first order as the one-pass kernel.
// =======================================================================================
constexpr int P2_NT = 128;
constexpr int P2_BH = P2_NT / 2;                    // 64 K* rows per CTA
constexpr int P2_BH_SLICE = P2_BH * KBY;            // 4096
constexpr int P2_STAGE = S * (A_SLICE + P2_BH_SLICE);   // 86016
constexpr int P2_NSTG = 2;
constexpr int P2_SMEM = P2_NSTG * P2_STAGE + 1024 + 256 + 4 * P2_NT * 2 * 8;
constexpr int P2_LOW = 4;                           // levels P2_LOW .. S-1 in pass 0, 0 .. P2_LOW-1 in pass 1

__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_pair2_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapKh, const OzArgs g,
                double* __restrict__ scratch, const int swap)
{
    extern __shared__ unsigned char raw[];
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + P2_NSTG * P2_STAGE, bar_empty = bar_full + 8 * P2_NSTG;
    const uint32_t bar_tfull = bar_empty + 8 * P2_NSTG, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
    const uint32_t red = base + P2_NSTG * P2_STAGE + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)cluster_rank();
    int ibp, cb;
    oz_tile_of((int)blockIdx.x / 2, g.nb / 2, g.ncb, g.group, ibp, cb);     // g.ncb = candidate blocks of 128 here
    const int ib = 2 * ibp + crank;
    const int nkb = (2 * ibp + 2) * TM / KBY;

    if (tid == 0) {
        for (int s = 0; s < P2_NSTG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 2);                                   // one arrival per CTA of the pair (used on rank 0)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            const int crow = cb * P2_NT + (crank ^ swap) * P2_BH;
            int it = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const int ns = pass == 0 ? S : P2_LOW;              // slices needed: levels >= 4 touch all, levels < 4 only 0..3
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % P2_NSTG;
                    if (it >= P2_NSTG) mbar_wait(bar_empty + 8 * s, (uint32_t)((it / P2_NSTG - 1) & 1));
                    const uint32_t st = base + s * P2_STAGE;
                    const uint32_t lbar = map_to_rank(bar_full + 8 * s, 0);
                    if (crank == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)(2 * ns * (A_SLICE + P2_BH_SLICE)));
                    for (int q = 0; q < ns; ++q) {
                        tma_2d_pair(st + q * A_SLICE, &mapP, kb * KBY, q * g.N + ib * TM, lbar);
                        tma_2d_pair(st + S * A_SLICE + q * P2_BH_SLICE, &mapKh, kb * KBY, q * g.Mc + crow, lbar);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = umma_idesc(2 * TM, P2_NT);
            int it = 0;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {                                     // both CTAs' epilogues have drained the low levels
                    mbar_wait(bar_tempty, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % P2_NSTG;
                    mbar_wait(bar_full + 8 * s, (uint32_t)((it / P2_NSTG) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * P2_STAGE;
                    if (pass == 0) {
#pragma unroll
                        for (int lvl = P2_LOW; lvl < S; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
#pragma unroll
                                for (int k = 0; k < KBY / UMMA_K; ++k)
                                    umma_i8_pair(tmem + (uint32_t)((lvl - P2_LOW) * P2_NT), umma_desc64(st + a * A_SLICE + k * UMMA_K),
                                                 umma_desc64(st + S * A_SLICE + (lvl - a) * P2_BH_SLICE + k * UMMA_K), idesc,
                                                 (uint32_t)((kb | a | k) != 0));
                    } else {
#pragma unroll
                        for (int lvl = 0; lvl < P2_LOW; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
#pragma unroll
                                for (int k = 0; k < KBY / UMMA_K; ++k)
                                    umma_i8_pair(tmem + (uint32_t)(lvl * P2_NT), umma_desc64(st + a * A_SLICE + k * UMMA_K),
                                                 umma_desc64(st + S * A_SLICE + (lvl - a) * P2_BH_SLICE + k * UMMA_K), idesc,
                                                 (uint32_t)((kb | a | k) != 0));
                    }
                    umma_commit_pair(bar_empty + 8 * s);
                }
                umma_commit_pair(bar_tfull);
            }
        }
    } else {
        const int lg = warp & 3;
        const int rl = lg * 32 + lane, row = ib * TM + rl;
        const double zr = g.z[row];
        const double rs = ldexp(1.0, g.eP[row] + g.eK[0]);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        double* sc = scratch + (size_t)(smid % 256) * TM * P2_NT;   // [column][row]: lanes = rows -> coalesced
        const uint32_t lane_base = tmem + ((uint32_t)(lg * 32) << 16);
        // pass 0: levels S-1 .. P2_LOW -> fp64 partial in scratch
        mbar_wait(bar_tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < P2_NT; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = S - 1; lvl >= P2_LOW; --lvl) {
                uint32_t d[32];
                tmem_ld32(lane_base + (uint32_t)((lvl - P2_LOW) * P2_NT + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) sc[(size_t)(c0 + j) * TM + rl] = v[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 64) mbar_arrive_remote(map_to_rank(bar_tempty, 0));
        // pass 1: levels P2_LOW-1 .. 0 on top
        mbar_wait(bar_tfull, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < P2_NT; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = sc[(size_t)(c0 + j) * TM + rl];
#pragma unroll 1
            for (int lvl = P2_LOW - 1; lvl >= 0; --lvl) {
                uint32_t d[32];
                tmem_ld32(lane_base + (uint32_t)(lvl * P2_NT + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
            double q2[32], qm[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double x = v[j] * rs; q2[j] = x * x; qm[j] = x * zr; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; ++j) {
                    const double keep2 = up ? q2[j + w] : q2[j], send2 = up ? q2[j] : q2[j + w];
                    const double keepm = up ? qm[j + w] : qm[j], sendm = up ? qm[j] : qm[j + w];
                    q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, w);
                    qm[j] = keepm + __shfl_xor_sync(0xffffffffu, sendm, w);
                }
            }
            const uint32_t slot = red + (uint32_t)(((lg * P2_NT) + c0 + lane) * 16);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(slot), "d"(q2[0]), "d"(qm[0]) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;                                     // 0 .. 127 = column
        double s2 = 0.0, sm = 0.0;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
            double a, b;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(red + (uint32_t)((w4 * P2_NT + et) * 16)));
            s2 += a; sm += b;
        }
        g.part_ssq[(long)ib * g.Mc + cb * P2_NT + et] = s2;
        g.part_mu[(long)ib * g.Mc + cb * P2_NT + et] = sm;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// =======================================================================================
// v4: 128 x 128 tiles in TWO passes over the contraction (levels 0..3, then 4..7): the int8 MMA reads both operands
// from shared memory at 128 B / clock / SM, so a 128 x 64 MMA (6 KB) is operand-fetch bound at 48 cycles instead of
// its 33 compute cycles; 128 x 128 (8 KB, 64 cycles) is balanced.  TMEM holds 4 accumulators of 128 columns per pass;
// the fp64 partial result of pass 1 waits in an L2-resident scratch tile (one per SM).  32-byte k-blocks (SWIZZLE_32B),
// 3 stages of 64 KB (pass 1 fills half of a stage: 4 + 4 slice tiles, pass 2 all 8 + 8).
// =======================================================================================
constexpr int KB2 = 32;
constexpr int T2 = 128;
constexpr int SL2 = T2 * KB2;                       // 4096 bytes per slice tile
constexpr int STAGE2 = 16 * SL2;                    // 65536
constexpr int NSTG2 = 3;
constexpr int OZ2_SMEM = NSTG2 * STAGE2 + 1024 + 256 + 4 * T2 * 8;

__device__ __forceinline__ uint64_t umma_desc32(uint32_t smem_addr) {      // K-major SWIZZLE_32B: 8 rows x 32 B atoms
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
__device__ __forceinline__ void mbar_arrive1(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

struct Oz2Args {
    int group;
    int nb, ncb;                        // row blocks of P, candidate blocks of 128
    int N, Mc;
    const int* eP; const int* eK;
    double* part_ssq;                   // [nb][Mc]
    double* scratch;                    // [nsmid][128][128]
};

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz2_vargemm_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapK, const Oz2Args g)
{
    extern __shared__ unsigned char raw[];
    const uint32_t base = (s_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + NSTG2 * STAGE2, bar_empty = bar_full + 8 * NSTG2;
    const uint32_t bar_tfull = bar_empty + 8 * NSTG2, bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
    const uint32_t red = base + NSTG2 * STAGE2 + 256;                // [4 lane groups][128 columns]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ib, cb;
    if (g.group > 0) oz_tile_of((int)blockIdx.x, g.nb, g.ncb, g.group, ib, cb);
    else { ib = g.nb - 1 - (int)blockIdx.x / g.ncb; cb = (int)blockIdx.x % g.ncb; }
    const int nkb = (ib + 1) * T2 / KB2;

    if (tid == 0) {
        for (int s = 0; s < NSTG2; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
                const int ns = pass == 0 ? 4 : S;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NSTG2;
                    if (it >= NSTG2) mbar_wait(bar_empty + 8 * s, (uint32_t)((it / NSTG2 - 1) & 1));
                    const uint32_t st = base + s * STAGE2;
                    mbar_expect_tx(bar_full + 8 * s, (uint32_t)(2 * ns * SL2));
                    for (int q = 0; q < ns; ++q) {
                        tma_2d(st + q * SL2, &mapP, kb * KB2, q * g.N + ib * T2, bar_full + 8 * s);
                        tma_2d(st + (8 + q) * SL2, &mapK, kb * KB2, q * g.Mc + cb * T2, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(T2, T2);
            int it = 0;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {                                      // the epilogue has read the first four accumulators
                    mbar_wait(bar_tempty, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NSTG2;
                    mbar_wait(bar_full + 8 * s, (uint32_t)((it / NSTG2) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * STAGE2;
                    if (pass == 0) {
#pragma unroll
                        for (int lvl = 0; lvl < 4; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
                                umma_i8(tmem + (uint32_t)(lvl * T2), umma_desc32(st + a * SL2), umma_desc32(st + (8 + lvl - a) * SL2),
                                        idesc, (uint32_t)((kb | a) != 0));
                    } else {
#pragma unroll
                        for (int lvl = 4; lvl < S; ++lvl)
#pragma unroll
                            for (int a = 0; a <= lvl; ++a)
                                umma_i8(tmem + (uint32_t)((lvl - 4) * T2), umma_desc32(st + a * SL2), umma_desc32(st + (8 + lvl - a) * SL2),
                                        idesc, (uint32_t)((kb | a) != 0));
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_tfull);
            }
        }
    } else {
        const int lg = warp & 3;
        const int rl = lg * 32 + lane, row = ib * T2 + rl;
        const double rs = ldexp(1.0, g.eP[row] + g.eK[0]);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        double* sc = g.scratch + ((size_t)smid * T2 + rl) * T2;
        const uint32_t lane_base = tmem + ((uint32_t)(lg * 32) << 16);
        // ---- pass 1 result -> scratch
        mbar_wait(bar_tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < T2; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = 3; lvl >= 0; --lvl) {
                uint32_t d[32];
                tmem_ld32(lane_base + (uint32_t)(lvl * T2 + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(sc + c0 + j) = make_double2(v[j], v[j + 1]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 64) mbar_arrive1(bar_tempty);
        // ---- pass 2 result + scratch -> squares -> column sums
        mbar_wait(bar_tfull, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < T2; c0 += 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0;
#pragma unroll 1
            for (int lvl = S - 1; lvl >= 4; --lvl) {
                uint32_t d[32];
                tmem_ld32(lane_base + (uint32_t)((lvl - 4) * T2 + c0), d);
                const double sf = ldexp(1.0, -8 * (lvl + 2));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fma((double)(int)d[j], sf, v[j]);
            }
            double q2[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 h = *reinterpret_cast<const double2*>(sc + c0 + j);
                const double x0 = (v[j] + h.x) * rs, x1 = (v[j + 1] + h.y) * rs;
                q2[j] = x0 * x0; q2[j + 1] = x1 * x1;
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
            asm volatile("st.shared.f64 [%0], %1;" :: "r"(red + (uint32_t)((lg * T2 + c0 + lane) * 8)), "d"(q2[0]) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - 64;
        double s2 = 0.0;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
            double a;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(red + (uint32_t)((w4 * T2 + et) * 8)));
            s2 += a;
        }
        g.part_ssq[(long)ib * g.Mc + cb * T2 + et] = s2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

__global__ void oz_finish_kernel(const double* part_ssq, const double* part_mu, int nb, long Mc, double* ssq, double* mu) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Mc) return;
    double a = 0.0, b = 0.0;
    for (int p = 0; p < nb; ++p) { a += part_ssq[(long)p * Mc + c]; b += part_mu[(long)p * Mc + c]; }
    ssq[c] = a; mu[c] = b;
}

// ---------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static void make_map(CUtensorMap* map, void* base, long rows, long cols, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols};
    cuuint32_t box[2] = {(cuuint32_t)KBY, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)r); exit(1); }
}

static void make_map32(CUtensorMap* map, void* base, long rows, long cols) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols};
    cuuint32_t box[2] = {(cuuint32_t)KB2, (cuuint32_t)T2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled(32B) %d\"}\n", (int)r); exit(1); }
}

struct Problem { int N, M; std::vector<double> P, Ks, z; double amp; };

static double matern52(double r2) { double r = std::sqrt(5.0 * r2); return (1.0 + r + 5.0 * r2 / 3.0) * std::exp(-r); }

// real GP data: K = Matern-5/2 (metric D/4) + 1e-3 I on uniform X, P = chol(K)^-1 (plain C++, N ~ 1024), K* for M candidates
static Problem gp_problem(int N, int M, int D) {
    Problem pr; pr.N = N; pr.M = M; pr.amp = 1.0;
    std::vector<double> X((size_t)N * D), Xs((size_t)M * D), K((size_t)N * N);
    srand(1234);
    for (auto& v : X) v = rand() / (double)RAND_MAX;
    for (auto& v : Xs) v = rand() / (double)RAND_MAX;
    const double metric = D / 4.0;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) {
            double r2 = 0; for (int d = 0; d < D; ++d) { double t = X[(size_t)i * D + d] - X[(size_t)j * D + d]; r2 += t * t / metric; }
            K[(size_t)i * N + j] = matern52(r2) + (i == j ? 1e-3 : 0.0);
        }
    for (int j = 0; j < N; ++j) {                               // in-place lower Cholesky
        double d = K[(size_t)j * N + j];
        for (int k = 0; k < j; ++k) d -= K[(size_t)j * N + k] * K[(size_t)j * N + k];
        d = std::sqrt(d); K[(size_t)j * N + j] = d;
        for (int i = j + 1; i < N; ++i) {
            double s = K[(size_t)i * N + j];
            for (int k = 0; k < j; ++k) s -= K[(size_t)i * N + k] * K[(size_t)j * N + k];
            K[(size_t)i * N + j] = s / d;
        }
    }
    pr.P.assign((size_t)N * N, 0.0);                            // P = L^-1, column by column
    for (int c = 0; c < N; ++c) {
        pr.P[(size_t)c * N + c] = 1.0 / K[(size_t)c * N + c];
        for (int i = c + 1; i < N; ++i) {
            double s = 0; for (int k = c; k < i; ++k) s -= K[(size_t)i * N + k] * pr.P[(size_t)k * N + c];
            pr.P[(size_t)i * N + c] = s / K[(size_t)i * N + i];
        }
    }
    pr.Ks.resize((size_t)M * N);
    for (int c = 0; c < M; ++c)
        for (int j = 0; j < N; ++j) {
            double r2 = 0; for (int d = 0; d < D; ++d) { double t = Xs[(size_t)c * D + d] - X[(size_t)j * D + d]; r2 += t * t / metric; }
            pr.Ks[(size_t)c * N + j] = matern52(r2);
        }
    pr.z.resize(N);
    for (auto& v : pr.z) v = rand() / (double)RAND_MAX - 0.5;
    return pr;
}

static Problem synthetic_problem(int N, int M) {               // timing only: lower-triangular noise with a decaying profile
    Problem pr; pr.N = N; pr.M = M; pr.amp = 1.0;
    pr.P.assign((size_t)N * N, 0.0); pr.Ks.resize((size_t)M * N); pr.z.resize(N);
    srand(99);
    for (int i = 0; i < N; ++i) for (int k = 0; k <= i; ++k) pr.P[(size_t)i * N + k] = (rand() / (double)RAND_MAX - 0.5) * std::exp(-0.002 * (i - k));
    for (auto& v : pr.Ks) v = rand() / (double)RAND_MAX;
    for (auto& v : pr.z) v = rand() / (double)RAND_MAX - 0.5;
    return pr;
}

struct Result { std::vector<double> ssq, mu; float ms_split_k, ms_gemm; };

template <int CL, int ORDER>
static void launch_oz(int grid, const CUtensorMap& mapP, const CUtensorMap& mapK, const OzArgs& a) {
    CKC(cudaFuncSetAttribute(oz_vargemm_kernel<CL, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = OZ_SMEM; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CKC(cudaLaunchKernelEx(&cfg, oz_vargemm_kernel<CL, ORDER>, mapP, mapK, a));
}

template <int NT>
static void launch_pair(int nb, int ncb, const CUtensorMap& mapP, const CUtensorMap& mapKh, const OzArgs& a, int swap) {
    using C = PairCfg<NT>;
    CKC(cudaFuncSetAttribute(oz_pair_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nb * ncb); cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CKC(cudaLaunchKernelEx(&cfg, oz_pair_kernel<NT>, mapP, mapKh, a, swap));
}

static void launch_pair2(int nb, int ncb128, const CUtensorMap& mapP, const CUtensorMap& mapKh, OzArgs a, double* scratch, int swap, int group) {
    CKC(cudaFuncSetAttribute(oz_pair2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
    a.ncb = ncb128; a.group = group;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nb * ncb128); cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = P2_SMEM; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CKC(cudaLaunchKernelEx(&cfg, oz_pair2_kernel, mapP, mapKh, a, scratch, swap));
}

// cl: 20 / 21 = CTA pair + two passes (256 x 128), K* halves in rank order / swapped;
// cl: 1 / 2 = cluster size of the multicast variant, 4 = A-sharing issue order, 10 / 11 = CTA pair (cta_group::2), K* halves
// in rank order / swapped
static Result run_gpu(const Problem& pr, int reps, int cl = 1, bool two_pass = false, bool grouped = false) {
    const int N = pr.N, M = pr.M, nb = N / TM, ncb = M / TN;
    double *dP, *dK, *dz, *dpss, *dpmu, *dssq, *dmu;
    int8_t *dPq, *dKq;
    int *deP, *deK;
    CKC(cudaMalloc(&dP, (size_t)N * N * 8)); CKC(cudaMalloc(&dK, (size_t)M * N * 8)); CKC(cudaMalloc(&dz, (size_t)N * 8));
    CKC(cudaMalloc(&dPq, (size_t)S * N * N)); CKC(cudaMalloc(&dKq, (size_t)S * M * N));
    CKC(cudaMalloc(&deP, (size_t)N * 4)); CKC(cudaMalloc(&deK, 4));
    CKC(cudaMalloc(&dpss, (size_t)nb * M * 8)); CKC(cudaMalloc(&dpmu, (size_t)nb * M * 8));
    CKC(cudaMalloc(&dssq, (size_t)M * 8)); CKC(cudaMalloc(&dmu, (size_t)M * 8));
    CKC(cudaMemcpy(dP, pr.P.data(), (size_t)N * N * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dK, pr.Ks.data(), (size_t)M * N * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dz, pr.z.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
    int ek = oz_exponent(pr.amp);                    // 0 < K* <= amp: one exponent for the whole matrix
    CKC(cudaMemcpy(deK, &ek, 4, cudaMemcpyHostToDevice));
    oz_rowmax_kernel<<<N, 256>>>(dP, N, N, deP);
    oz_split_kernel<<<(unsigned)(((size_t)N * N + 255) / 256), 256>>>(dP, N, N, deP, 0, dPq);
    CKC(cudaGetLastError());
    CUtensorMap mapP, mapK, mapP32, mapK32;
    make_map(&mapP, dPq, (long)S * N, N, TM);
    make_map(&mapK, dKq, (long)S * M, N, TN);
    CUtensorMap mapKh;
    make_map(&mapKh, dKq, (long)S * M, N, TN / 2);
    make_map32(&mapP32, dPq, (long)S * N, N);
    make_map32(&mapK32, dKq, (long)S * M, N);
    double* dscr = nullptr;
    CKC(cudaMalloc(&dscr, (size_t)256 * T2 * T2 * 8));
    CKC(cudaFuncSetAttribute(oz2_vargemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM));
    Oz2Args a2; a2.group = grouped ? 16 : 0; a2.nb = nb; a2.ncb = M / T2; a2.N = N; a2.Mc = M; a2.eP = deP; a2.eK = deK; a2.part_ssq = dpss; a2.scratch = dscr;
    OzArgs a; a.group = grouped ? 32 : 0; a.nb = nb; a.ncb = ncb; a.N = N; a.Mc = M; a.eP = deP; a.eK = deK; a.z = dz; a.part_ssq = dpss; a.part_mu = dpmu;
    cudaEvent_t e0, e1, e2;
    CKC(cudaEventCreate(&e0)); CKC(cudaEventCreate(&e1)); CKC(cudaEventCreate(&e2));
    Result res; res.ms_split_k = res.ms_gemm = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CKC(cudaEventRecord(e0));
        oz_split_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(dK, M, N, deK, 1, dKq);
        CKC(cudaEventRecord(e1));
        if (two_pass) oz2_vargemm_kernel<<<nb * (M / T2), OZ_THREADS, OZ2_SMEM>>>(mapP32, mapK32, a2);
        else if (cl >= 20) launch_pair2(nb, M / P2_NT, mapP, mapK, a, dscr, cl - 20, grouped ? 16 : 8);
        else if (cl >= 10) launch_pair<TN>(nb, ncb, mapP, mapKh, a, cl - 10);
        else if (cl == 1) launch_oz<1, 0>(nb * ncb, mapP, mapK, a);
        else if (cl == 2) launch_oz<2, 0>(nb * ncb, mapP, mapK, a);
        else launch_oz<1, 1>(nb * ncb, mapP, mapK, a);            // cl == 4 slot reused: A-sharing issue order
        oz_finish_kernel<<<(M + 255) / 256, 256>>>(dpss, dpmu, nb, M, dssq, dmu);
        CKC(cudaEventRecord(e2));
        CKC(cudaGetLastError());
        CKC(cudaDeviceSynchronize());
        float m1, m2;
        CKC(cudaEventElapsedTime(&m1, e0, e1)); CKC(cudaEventElapsedTime(&m2, e1, e2));
        res.ms_split_k = std::min(res.ms_split_k, m1); res.ms_gemm = std::min(res.ms_gemm, m2);
    }
    res.ssq.resize(M); res.mu.resize(M);
    CKC(cudaMemcpy(res.ssq.data(), dssq, (size_t)M * 8, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(res.mu.data(), dmu, (size_t)M * 8, cudaMemcpyDeviceToHost));
    cudaFree(dP); cudaFree(dK); cudaFree(dz); cudaFree(dPq); cudaFree(dKq); cudaFree(deP); cudaFree(deK);
    cudaFree(dpss); cudaFree(dpmu); cudaFree(dssq); cudaFree(dmu); cudaFree(dscr);
    return res;
}

int main(int argc, char** argv) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    g_encode = (EncodeFn)p;
    if (argc > 1 && std::string(argv[1]) == "pair2") {
        Problem pr = gp_problem(1024, 256, 16);
        Result r = run_gpu(pr, 1, 1, false, true), p0 = run_gpu(pr, 1, 20, false, true), p1 = run_gpu(pr, 1, 21, false, true);
        double d0 = 0, d1 = 0;
        for (int c = 0; c < pr.M; ++c) {
            d0 = std::max(d0, std::fabs(r.ssq[c] - p0.ssq[c]) + std::fabs(r.mu[c] - p0.mu[c]));
            d1 = std::max(d1, std::fabs(r.ssq[c] - p1.ssq[c]) + std::fabs(r.mu[c] - p1.mu[c]));
        }
        const int swap = d1 < d0 ? 1 : 0;
        Problem big = synthetic_problem(4096, 16384);
        Result t = run_gpu(big, 4, 1, false, true), tp = run_gpu(big, 4, 20 + swap, false, true);
        double dbig = 0;
        for (int c = 0; c < big.M; ++c) dbig = std::max(dbig, std::fabs(t.ssq[c] - tp.ssq[c]));
        printf("{\"probe\": \"CTA pair (cta_group::2) 256x128 in two passes, S=%d\", \"max_abs_diff_halves_in_rank_order\": %.3e, "
               "\"max_abs_diff_halves_swapped\": %.3e, \"swap\": %d, \"single_cta_ms_gemm\": %.4f, \"pair_two_pass_ms_gemm\": %.4f, "
               "\"c2_max_abs_diff\": %.3e}\n", S, d0, d1, swap, t.ms_gemm, tp.ms_gemm, dbig);
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "pair") {
        // CTA-pair kernel alone (its own process: a protocol mistake would hang, not just miscompute)
        Problem pr = gp_problem(1024, 256, 16);
        Result r = run_gpu(pr, 1, 1, false, true), p0 = run_gpu(pr, 1, 10, false, true), p1 = run_gpu(pr, 1, 11, false, true);
        double d0 = 0, d1 = 0;
        for (int c = 0; c < pr.M; ++c) {
            d0 = std::max(d0, std::fabs(r.ssq[c] - p0.ssq[c]) + std::fabs(r.mu[c] - p0.mu[c]));
            d1 = std::max(d1, std::fabs(r.ssq[c] - p1.ssq[c]) + std::fabs(r.mu[c] - p1.mu[c]));
        }
        const int swap = d1 < d0 ? 1 : 0;
        Problem big = synthetic_problem(4096, 16384);
        Result t = run_gpu(big, 4, 1, false, true), tp = run_gpu(big, 4, 10 + swap, false, true);
        double dbig = 0;
        for (int c = 0; c < big.M; ++c) dbig = std::max(dbig, std::fabs(t.ssq[c] - tp.ssq[c]));
        printf("{\"probe\": \"CTA pair (cta_group::2) 256x%d, S=%d\", \"max_abs_diff_halves_in_rank_order\": %.3e, "
               "\"max_abs_diff_halves_swapped\": %.3e, \"swap\": %d, \"single_cta_ms_gemm\": %.4f, \"pair_ms_gemm\": %.4f, "
               "\"c2_max_abs_diff\": %.3e}\n", TN, S, d0, d1, swap, t.ms_gemm, tp.ms_gemm, dbig);
        return 0;
    }
    // ---- accuracy on real GP data (N = 1024, D = 16, 256 candidates) against an 80-bit CPU contraction
    Problem pr = gp_problem(1024, 256, 16);
    Result r = run_gpu(pr, 1, 1), r2 = run_gpu(pr, 1, 2), r4 = run_gpu(pr, 1, 4), rt = run_gpu(pr, 1, 1, true);
    double worst_var2 = 0;
    double cl_diff = 0;
    for (int c = 0; c < pr.M; ++c)
        cl_diff = std::max(cl_diff, std::max(std::fabs(r.ssq[c] - r2.ssq[c]), std::fabs(r.ssq[c] - r4.ssq[c])) + std::max(std::fabs(r.mu[c] - r2.mu[c]), std::fabs(r.mu[c] - r4.mu[c])));
    double worst_var = 0, worst_mu = 0, var_min = 1e300;
    for (int c = 0; c < pr.M; ++c) {
        long double ssq = 0, mu = 0;
        for (int i = 0; i < pr.N; ++i) {
            long double v = 0;
            for (int k = 0; k <= i; ++k) v += (long double)pr.P[(size_t)i * pr.N + k] * (long double)pr.Ks[(size_t)c * pr.N + k];
            ssq += v * v; mu += v * (long double)pr.z[i];
        }
        const double var_ref = (double)((long double)pr.amp - ssq), var = pr.amp - r.ssq[c];
        var_min = std::min(var_min, var_ref);
        worst_var = std::max(worst_var, std::fabs(var - var_ref) / std::max(std::fabs(var_ref), 1e-6 * pr.amp));
        worst_mu = std::max(worst_mu, std::fabs(r.mu[c] - (double)mu) / std::max(std::fabs((double)mu), 1.0));
        worst_var2 = std::max(worst_var2, std::fabs((pr.amp - rt.ssq[c]) - var_ref) / std::max(std::fabs(var_ref), 1e-6 * pr.amp));
    }
    // ---- timing on the C2 shape
    Problem big = synthetic_problem(4096, 16384);
    Result t = run_gpu(big, 4, 1), t2 = run_gpu(big, 4, 2), t4 = run_gpu(big, 4, 4), tt = run_gpu(big, 4, 1, true);
    Result tg = run_gpu(big, 4, 1, false, true), ttg = run_gpu(big, 4, 1, true, true);
    Result tg2 = run_gpu(big, 4, 2, false, true), tg4 = run_gpu(big, 4, 4, false, true);
    Result rg2 = run_gpu(pr, 1, 2, false, true), rg4 = run_gpu(pr, 1, 4, false, true);
    double grp_cl_diff = 0;
    for (int c = 0; c < pr.M; ++c) grp_cl_diff = std::max(grp_cl_diff, std::max(std::fabs(r.ssq[c] - rg2.ssq[c]), std::fabs(r.ssq[c] - rg4.ssq[c])));
    Result rg = run_gpu(pr, 1, 1, false, true), rtg = run_gpu(pr, 1, 1, true, true);
    double grp_diff = 0;
    for (int c = 0; c < pr.M; ++c) grp_diff = std::max(grp_diff, std::max(std::fabs(r.ssq[c] - rg.ssq[c]), std::fabs(rt.ssq[c] - rtg.ssq[c])));
    const double flops = 16384.0 * (4096.0 * 4096.0 + 2 * 4096.0);
    printf("{\"probe\": \"Ozaki int8 variance contraction, S=%d slices, tile %dx%d\", \"scaled_var_err_vs_80bit\": %.3e, "
           "\"mu_err\": %.3e, \"var_min\": %.3e, \"c2_chunk_ms_gemm\": %.4f, \"c2_chunk_ms_split_kstar\": %.4f, "
           "\"fp64_equiv_tflops_gemm\": %.2f, \"fp64_equiv_tflops_incl_split\": %.2f, \"dmma_reference_tflops\": 35.2, \"cluster2_ms_gemm\": %.4f, \"a_sharing_order_ms_gemm\": %.4f, \"cluster_vs_plain_max_abs_diff\": %.3e, \"two_pass_128x128_ms_gemm\": %.4f, \"two_pass_scaled_var_err\": %.3e, \"two_pass_fp64_equiv_tflops\": %.2f, \"grouped_ms_gemm\": %.4f, \"two_pass_grouped_ms_gemm\": %.4f, \"grouped_max_abs_diff\": %.3e, \"grouped_cluster2_ms_gemm\": %.4f, \"grouped_cluster4_ms_gemm\": %.4f, \"grouped_cluster_max_abs_diff\": %.3e}\n",
           S, TM, TN, worst_var, worst_mu, var_min, t.ms_gemm, t.ms_split_k, flops / (t.ms_gemm * 1e-3) / 1e12,
           flops / ((t.ms_gemm + t.ms_split_k) * 1e-3) / 1e12, t2.ms_gemm, t4.ms_gemm, cl_diff, tt.ms_gemm, worst_var2, flops / (tt.ms_gemm * 1e-3) / 1e12, tg.ms_gemm, ttg.ms_gemm, grp_diff, tg2.ms_gemm, tg4.ms_gemm, grp_cl_diff);
    return 0;
}
