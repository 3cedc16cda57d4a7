// Blocked diagonal-block kernel (option diag = 3; same contract as gpk_potrf_diag_fused_kernel):
//   L_kk = chol(A_kk) and inv(L_kk) of one 128 x 128 diagonal block, one CTA of 256 threads.
//
// The column-by-column kernel pays one block-wide barrier per pivot (128 intervals of ~950 cycles; 3/4 of its
// issued instructions are not arithmetic).  Here the block is processed in 8 panels of 16 columns with two
// barriers per panel:
//   S  "factor + solve": EVERY warp factorises the published 16 x 16 diagonal sub-block redundantly in registers
//      (lane & 15 = row, columns broadcast with shuffles, the rsqrt of the next pivot computed one column ahead so the
//      serial chain per pivot is mul -> fma -> rsqrt).  The broadcast column is exactly what a forward substitution
//      against the sub-block needs, so each thread carries one 16-vector through the same loop for free:
//      threads 0..111 a row of the panel below the sub-block (L_ik = A_ik L_kk^-T), threads 128..255 a column of
//      the inverse's row block (X_k. = L_kk^-1 Xtilde_k.).  Finished values go to K / P / Q straight from registers.
//   U  "update + publish": rank-16 update of the 8 x 8 cyclic register tiles of the trailing matrix and of the
//      inverse's residual (kept in shared memory, staged through registers for the panel's k loop), operands read
//      with conflict-free / broadcast 64-bit shared loads; then the next panel's sub-block and rows are published.
// The arithmetic order per element (pivots subtracted in increasing order) is the one of the column-by-column
// kernel, so the factors agree to rounding.  The panel loop is rolled and only U is specialised per panel (static
// register indices): the kernel runs once per launch on one SM, i.e. out of a cold instruction cache, and a first
// fully unrolled version (38k instructions) was fetch-bound and slower than the kernel it replaces.
#pragma once

constexpr int D3PS = 17;      // row stride (doubles) of the panel staging array: odd -> thread-per-row and
                              // 16-rows-per-half-warp 64-bit loads are bank-conflict-free
constexpr int D3XS = 130;     // row stride of the inverse's residual
struct __align__(16) D3Smem {
    double xs[128 * D3XS];        // Xtilde / finished rows of inv(L_kk), row-major
    double pan[2][128 * D3PS];    // pan[i][k]: column 16 kb + k of the current panel, row i (raw, then solved)
    double din[2][16 * 17];       // the 16 x 16 diagonal sub-block as published
};
constexpr int DIAG3_SMEM = (int)sizeof(D3Smem);

__device__ __forceinline__ double d3_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1/sqrt(p): hardware seed (MUFU.RSQ64H, ~22 bits) + two Newton steps -> ~1 ulp; 9 instructions, no branches
// (the library rsqrt() carries special-case paths this chain does not need: p is a checked positive finite pivot).
__device__ __forceinline__ double d3_rsqrt(double p)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    double t = p * y, e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = p * y;
    e = fma(-t, y, 1.0);
    return fma(0.5 * y, e, y);
}

__device__ __forceinline__ void d3_store16(double* __restrict__ dst, const double (&v)[16])
{
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<double2*>(dst + 2 * q) = make_double2(v[2 * q], v[2 * q + 1]);
}

// ---- publish the raw sub-block and the panel below it (register indices are static)
template <int KB>
__device__ __forceinline__ void d3_publish(const double (&A)[8][8], D3Smem& sm, int ty, int tx)
{
    constexpr int PB = KB & 1;
    sm.din[PB][ty * 17 + tx] = A[KB][KB];
#pragma unroll
    for (int a = KB + 1; a < 8; ++a) sm.pan[PB][(ty + 16 * a) * D3PS + tx] = A[a][KB];
}

// ---- U: rank-16 updates after panel KB, then publish panel KB + 1
template <int KB>
__device__ __forceinline__ void d3_update(double (&A)[8][8], D3Smem& sm, int ty, int tx)
{
    constexpr int PB = KB & 1;
    const double* pr = sm.pan[PB] + ty * D3PS;
    const double* pc = sm.pan[PB] + tx * D3PS;
    const double* xrow = sm.xs + (16 * KB) * D3XS + tx;
    double* xown = sm.xs + ty * D3XS + tx;
    double xa[8][8];
#pragma unroll
    for (int a = KB + 1; a < 8; ++a)
#pragma unroll
        for (int b = 0; b <= KB; ++b) xa[a][b] = xown[16 * a * D3XS + 16 * b];
#pragma unroll 2
    for (int k = 0; k < 16; ++k) {
        double lr[8], lc[8], xc[8];
#pragma unroll
        for (int a = KB + 1; a < 8; ++a) lr[a] = pr[16 * a * D3PS + k];
#pragma unroll
        for (int b = KB + 1; b < 8; ++b) lc[b] = pc[16 * b * D3PS + k];
#pragma unroll
        for (int b = 0; b <= KB; ++b) xc[b] = xrow[k * D3XS + 16 * b];
#pragma unroll
        for (int b = KB + 1; b < 8; ++b)
#pragma unroll
            for (int a = b; a < 8; ++a) A[a][b] = fma(-lr[a], lc[b], A[a][b]);
#pragma unroll
        for (int b = 0; b <= KB; ++b)
#pragma unroll
            for (int a = KB + 1; a < 8; ++a) xa[a][b] = fma(-lr[a], xc[b], xa[a][b]);
    }
#pragma unroll
    for (int a = KB + 1; a < 8; ++a)
#pragma unroll
        for (int b = 0; b <= KB; ++b) xown[16 * a * D3XS + 16 * b] = xa[a][b];
    d3_publish<KB + 1>(A, sm, ty, tx);
}

#define D3_CASES7(F, ...) switch (kbp) { case 0: F<0>(__VA_ARGS__); break; case 1: F<1>(__VA_ARGS__); break; \
    case 2: F<2>(__VA_ARGS__); break; case 3: F<3>(__VA_ARGS__); break; case 4: F<4>(__VA_ARGS__); break; \
    case 5: F<5>(__VA_ARGS__); break; default: F<6>(__VA_ARGS__); break; }

__global__ void __launch_bounds__(256, 1)
gpk_potrf_diag_blocked_kernel(double* __restrict__ K, long ld, int kb,
                              double* __restrict__ P, double* __restrict__ Q, long ldp,
                              int* __restrict__ status, double* __restrict__ logdet_part,
                              long long* __restrict__ prof)
{
    extern __shared__ __align__(16) unsigned char d3_raw[];
    D3Smem& sm = *reinterpret_cast<D3Smem*>(d3_raw);
    __shared__ int s_bad;

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, r = lane & 15;
    cudaGridDependencySynchronize();      // programmatic dependent launch (see gpk_gemm_nt_kernel)
    if (*status != 0) return;
    if (tid == 0) s_bad = 0;
    const bool stamp = prof != nullptr && tid == 255;      // a thread outside warp 0 (which owns the outputs of S)
    if (stamp) prof[0] = clock64();

    double* Kt = K + (long)kb * 128 * ld + (long)kb * 128;
    double* Pt = P + (long)kb * 128 * ldp + (long)kb * 128;
    double* Qt = Q + (long)kb * 128 * ldp + (long)kb * 128;
    double A[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ty + 16 * a, c = tx + 16 * b;
            A[a][b] = (b <= a && c <= i) ? Kt[(long)i * ld + c] : 0.0;
        }
    for (int e = tid; e < 128 * D3XS; e += 256) {          // Xtilde = I
        const int i = e / D3XS, c = e - i * D3XS;
        sm.xs[e] = (i == c) ? 1.0 : 0.0;
    }
    d3_publish<0>(A, sm, ty, tx);
    double lsum = 0.0;
    __syncthreads();
    if (stamp) prof[1] = clock64();

#pragma unroll 1
    for (int kbp = 0; kbp < 8; ++kbp) {   // panel loop rolled: phase S exists once in the binary
        const int pb = kbp & 1;
        double* pan = sm.pan[pb];
        const double* din = sm.din[pb];

        // ---- S: factor the 16 x 16 sub-block (every warp, redundantly) + one forward substitution per thread
        const int irow = 16 * (kbp + 1) + tid;                   // panel row of threads 0..111
        const int ccol = tid - 128;                              // inverse column of threads 128..255
        const bool is_pan = tid < 128 && irow < 128;
        const bool is_x = tid >= 128 && ccol < 16 * (kbp + 1);
        double d[16], v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) d[c] = (c <= r) ? din[r * 17 + c] : 0.0;
        {
            const double* src = is_pan ? pan + irow * D3PS : sm.xs + (16 * kbp) * D3XS + (is_x ? ccol : 0);
            const int step = is_pan ? 1 : D3XS;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                v[k] = 0.0;
                if (is_pan || is_x) v[k] = src[k * step];        // idle threads must not touch what others write
            }
        }
        const double p0 = d3_shfl(d[0], 0);
        bool ok = (p0 > 0.0) && !isinf(p0);
        int bad = ok ? 0 : kb * 128 + 16 * kbp + 1;      // first failing pivot (1-based), in a register: no branches
        const double r0 = d3_rsqrt(p0);
        double rs = ok ? r0 : 1.0;
        double x = d3_shfl(d[0], 1), y = d3_shfl(d[1], 1);       // row j+1: its entry in column j and its diagonal
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double l = d[j] * rs;              // column j of L_D for rows >= j (lane j: sqrt of the pivot)
            d[j] = l;
            const double yj = v[j] * rs;             // forward substitution: unknown j of this thread's vector
            v[j] = yj;
            double rs_next = 1.0;
            if (j < 15) {                            // next pivot, by the same fma its owner applies below
                const double ln = x * rs;
                const double pn = fma(-ln, ln, y);
                ok = (pn > 0.0) && !isinf(pn);
                bad = (bad == 0 && !ok) ? kb * 128 + 16 * kbp + j + 2 : bad;
                const double rn = d3_rsqrt(pn);      // unconditional: stays in the straight-line block
                rs_next = ok ? rn : 1.0;
            }
#pragma unroll
            for (int c = j + 1; c < 16; ++c) {
                const double lc = d3_shfl(l, c);     // L_D[c][j], from the lane that owns row c
                d[c] = fma(-l, lc, d[c]);
                v[c] = fma(-lc, yj, v[c]);
            }
            if (j < 14) {
                x = d3_shfl(d[j + 1], j + 2);
                y = d3_shfl(d[j + 2], j + 2);
            }
            rs = rs_next;
        }
        if (is_pan) {                                 // finished panel row: operand of U and final L values
#pragma unroll
            for (int k = 0; k < 16; ++k) pan[irow * D3PS + k] = v[k];
            d3_store16(Kt + (long)irow * ld + 16 * kbp, v);
        }
        if (tid >= 128) {                             // finished column of the inverse's row block (zeros right of it)
            if (is_x) {
#pragma unroll
                for (int k = 0; k < 16; ++k) sm.xs[(16 * kbp + k) * D3XS + ccol] = v[k];
            }
            d3_store16(Qt + (long)ccol * ldp + 16 * kbp, v);                             // Q = P^T (upper)
#pragma unroll
            for (int k = 0; k < 16; ++k) Pt[(long)(16 * kbp + k) * ldp + ccol] = v[k];   // P (lower), coalesced
        }
        if (tid < 16) {                               // warp 0 publishes the factor of the sub-block
            double row[16];
            double mine = 1.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                row[c] = (c <= r) ? d[c] : 0.0;
                if (c == r) mine = d[c];
            }
            d3_store16(Kt + (long)(16 * kbp + r) * ld + 16 * kbp, row);
            lsum += log(mine);
            if (tid == 0 && bad != 0 && s_bad == 0) s_bad = bad;
        }
        __syncthreads();
        if (stamp) prof[2 + 2 * kbp] = clock64();
        if (kbp == 7) break;

        // ---- U: zero fill right of the sub-block, rank-16 updates, publish the next panel
        for (int b = kbp + 1; b < 8; ++b) Kt[(long)(16 * kbp + ty) * ld + 16 * b + tx] = 0.0;
        D3_CASES7(d3_update, A, sm, ty, tx)
        __syncthreads();
        if (stamp) prof[3 + 2 * kbp] = clock64();
    }

    if (tid < 32) {                       // log-det partial: sum over the 16 rows of the 8 panels
        double s = (tid < 16) ? lsum : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (tid == 0) {
            logdet_part[kb] = s;
            if (s_bad != 0) atomicCAS(status, 0, s_bad);
        }
    }
    if (stamp) prof[33] = clock64();
}
