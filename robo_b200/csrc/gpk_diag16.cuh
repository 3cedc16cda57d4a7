// Blocked diagonal-block kernel (option diag = 3; same contract as gpk_potrf_diag_fused_kernel):
//   L_kk = chol(A_kk) and inv(L_kk) of one 128 x 128 diagonal block, one CTA of 256 threads.
//
// The column-by-column kernel pays one block-wide barrier per pivot (128 intervals of ~950 cycles; 3/4 of its
// issued instructions are not arithmetic).  Here the block is processed in 8 panels of 16 columns with two
// barriers per panel:
//   S  "factor + solve": warp 0 eliminates the published 16 x 16 diagonal sub-block in registers (lane & 15 = row)
//      in square-root-free form (L' D L'^T): the serial chain per pivot is mul -> fma -> reciprocal, the raw column is
//      broadcast with shuffles before its pivot's reciprocal is known, and the 16 rsqrt that give the Cholesky factor
//      L = L' D^1/2 run in parallel after the last pivot.  Every column of L' is published through shared memory and
//      signalled with a named barrier (bar.arrive, one id per pivot); warps 1..4 wait on that id (bar.sync) and apply
//      it to one unit-lower forward substitution per thread -- the 128 - 16(p+1) rows of the panel below the
//      sub-block (L_ik = A_ik L_kk^-T) and the 16(p+1) columns of the inverse's row block (X_k. = L_kk^-1 Xtilde_k.),
//      always 128 vectors -- then scale by D^-1/2.  (Earlier versions, measured: all 8 warps factorising redundantly
//      with the substitutions in the same loop = 4.7k cycles per panel, DP-issue bound; one warp carrying the sqrt
//      chain AND 32 substitutions = 5.2k.)
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
    double colp[16][16];          // colp[j][c] = L_D[c][j]: column j as published by the factorising warp
    double rsp[16];               // 1 / L_D[j][j]
    double ub[2][16];             // raw column of the sub-block about to be eliminated (factorising warp only)
    double ldd[16 * 17];          // the factorised sub-block (lower), for the coalesced store in U
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

// 1/p: hardware seed (MUFU.RCP64H) + two Newton steps; 5 instructions on the pivot chain
__device__ __forceinline__ double d3_rcp(double p)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    y = fma(y, fma(-p, y, 1.0), y);
    return fma(y, fma(-p, y, 1.0), y);
}

// named barriers between the factorising warp (arrive) and the four substitution warps (sync): 32 + 128 threads
__device__ __forceinline__ void d3_bar_arrive(int id) { asm volatile("bar.arrive %0, 160;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void d3_bar_sync(int id) { asm volatile("bar.sync %0, 160;" :: "r"(id) : "memory"); }

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
    const bool stamp = prof != nullptr && tid == 159;      // last substitution thread
    if (stamp) prof[0] = clock64();

    double* Kt = K + (long)kb * 128 * ld + (long)kb * 128;
    double* Pt = P + (long)kb * 128 * ldp + (long)kb * 128;
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

        // ---- S: warp 0 factorises the 16 x 16 sub-block, warps 1..4 run the 128 forward substitutions
        const int u = tid - 32;                                  // vector of threads 32..159
        const int nrows = 128 - 16 * (kbp + 1);                  // panel rows below the sub-block
        const bool is_vec = u >= 0 && u < 128;
        const bool is_pan = is_vec && u < nrows;
        const int irow = 16 * (kbp + 1) + u;                     // panel row (is_pan)
        const int ccol = u - nrows;                              // column of the inverse's row block (is_x)
        const bool fine = stamp && kbp == 3;
        if (fine) prof[34] = clock64();
        if (prof != nullptr && kbp == 3 && tid == 0) prof[46] = clock64();
        if (tid < 32) {
            // -- the factorising warp.  Square-root-free elimination (L' D L'^T with unit-lower L'): the serial chain
            // per pivot is  mul -> fma -> reciprocal  (the raw column is broadcast before its pivot's reciprocal is
            // known), and the 16 rsqrt that turn L' D^1/2 into the Cholesky factor run in parallel at the end.
            double d[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) d[c] = (c <= r) ? din[r * 17 + c] : 0.0;
            double p = d3_shfl(d[0], 0);
            bool ok = (p > 0.0) && !isinf(p);
            int bad = ok ? 0 : kb * 128 + 16 * kbp + 1;  // first failing pivot (1-based), in a register: no branches
            double myp = p;                              // lane r keeps pivot r
            const double i0 = d3_rcp(p);
            double inv = ok ? i0 : 1.0;
            double x = d3_shfl(d[0], 1), y = d3_shfl(d[1], 1);   // row j+1: its entry in column j and its diagonal
            if (lane < 16) sm.ub[0][r] = d[0];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                // column j of the unit-lower factor, published for the substitutions.  The signal comes FIRST: the
                // barrier instruction fences the scheduler, and the two independent halves of this pivot -- the
                // reciprocal of the next pivot (the serial chain) and the update with column j -- must be free to
                // interleave on this one in-order warp (with the chain ahead of the signal they ran back to back:
                // 200 cycles per pivot instead of ~100).
                const double t = d[j] * inv;
                if (lane < 16) sm.colp[j][r] = t;
                if (j < 14) d3_bar_arrive(1 + j);
                const double t1 = x * inv;
                const double pn = fma(-t1, x, y);
                ok = (pn > 0.0) && !isinf(pn);
                bad = (bad == 0 && !ok) ? kb * 128 + 16 * kbp + j + 2 : bad;
                const double rn = d3_rcp(pn);            // unconditional: stays in the straight-line block
                const double inv_next = ok ? rn : 1.0;
                myp = (r == j + 1) ? pn : myp;
                if (j == 14) {                           // all pivots known: reciprocal square roots, then the last signal
                    const double rsq = d3_rsqrt(myp);
                    if (lane < 16) sm.rsp[r] = rsq;
                    d3_bar_arrive(15);
                }
                // raw entries (c, j) come back as broadcast loads of the vector stored one pivot earlier (they do not
                // wait for the reciprocal; 2 x 15 shuffles per pivot cost a single warp ~4 issue cycles each)
                const double* ubj = sm.ub[j & 1];
#pragma unroll
                for (int c = j + 1; c < 16; ++c) d[c] = fma(-t, ubj[c], d[c]);
                if (lane < 16) sm.ub[(j + 1) & 1][r] = d[j + 1];
                __syncwarp();
                if (j < 14) {
                    x = d3_shfl(d[j + 1], j + 2);
                    y = d3_shfl(d[j + 2], j + 2);
                }
                inv = inv_next;
                if (prof != nullptr && kbp == 3 && tid == 0 && (j & 3) == 3) prof[42 + (j >> 2)] = clock64();
            }
            if (prof != nullptr && kbp == 3 && tid == 0) prof[45] = clock64();
            __syncwarp();
            if (tid < 16) {                               // the Cholesky factor of the sub-block, for the store in U
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const double sc = (c == r) ? myp : d[c];          // raw column entry (or the pivot) * 1/sqrt(pivot c)
                    sm.ldd[r * 17 + c] = (c <= r) ? sc * sm.rsp[c] : 0.0;
                }
                lsum += 0.5 * log(myp);
                if (tid == 0 && bad != 0 && s_bad == 0) s_bad = bad;
            }
        } else if (is_vec) {
            // -- one forward substitution per thread against the unit-lower factor, column by column as published
            double v[16];
            {
                const double* src = is_pan ? pan + irow * D3PS : sm.xs + (16 * kbp) * D3XS + ccol;
                const int step = is_pan ? 1 : D3XS;
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = src[k * step];
            }
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                d3_bar_sync(1 + j);
                const double wj = v[j];
                const double* col = sm.colp[j];
                if ((j + 1) & 1) v[j + 1] = fma(-col[j + 1], wj, v[j + 1]);
#pragma unroll
                for (int c = (j + 2) & ~1; c < 16; c += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(col + c);
                    v[c] = fma(-l2.x, wj, v[c]);
                    v[c + 1] = fma(-l2.y, wj, v[c + 1]);
                }
                if (fine && (j & 3) == 3) prof[35 + (j >> 2)] = clock64();
            }
            if (fine) prof[38] = clock64();
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] *= sm.rsp[k];          // D^-1/2
            if (is_pan) {                                 // finished panel row: operand of U
#pragma unroll
                for (int k = 0; k < 16; ++k) pan[irow * D3PS + k] = v[k];
            } else {                                      // finished column of the inverse's row block
#pragma unroll
                for (int k = 0; k < 16; ++k) sm.xs[(16 * kbp + k) * D3XS + ccol] = v[k];
            }
        }
        if (fine) prof[39] = clock64();
        __syncthreads();
        if (stamp) prof[2 + 2 * kbp] = clock64();
        // ---- U: coalesced stores of what panel kbp finished (they drain behind the arithmetic), rank-16 updates,
        // publish the next panel.  The zeros right of the sub-block and Q = P^T are written off the critical chain
        // (gpk_diag_prezero_kernel before the factorisation, gpk_diag_qfill_kernel after it).
        {
            const long grow = 16 * kbp + ty;                                     // row of the panel's row block
            Kt[grow * ld + 16 * kbp + tx] = sm.ldd[ty * 17 + tx];
            for (int a = kbp + 1; a < 8; ++a)                                    // solved panel rows below
                Kt[(long)(ty + 16 * a) * ld + 16 * kbp + tx] = pan[(ty + 16 * a) * D3PS + tx];
#pragma unroll
            for (int b = 0; b < 8; ++b) Pt[grow * ldp + tx + 16 * b] = sm.xs[grow * D3XS + tx + 16 * b];   // P (lower)
        }
        if (kbp == 7) break;
        if (fine) prof[40] = clock64();
        D3_CASES7(d3_update, A, sm, ty, tx)
        if (fine) prof[41] = clock64();
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

// Off-chain helpers of the blocked kernel: zero what lies right of the 16 x 16 sub-blocks in every diagonal tile of
// K (the covariance builder leaves the symmetric values there; later GEMMs read the tile as a lower-triangular
// operand), and Q's diagonal tiles = transposed diagonal tiles of P.  One CTA per diagonal tile.
__global__ void __launch_bounds__(256) gpk_diag_prezero_kernel(double* __restrict__ K, long ld)
{
    double* Kt = K + (long)blockIdx.x * 128 * ld + (long)blockIdx.x * 128;
    for (int e = threadIdx.x; e < 128 * 128; e += 256) {
        const int i = e >> 7, c = e & 127;
        if ((c >> 4) > (i >> 4)) Kt[(long)i * ld + c] = 0.0;
    }
}

__global__ void __launch_bounds__(256) gpk_diag_qfill_kernel(const double* __restrict__ P, double* __restrict__ Q,
                                                             long ldp, const int* __restrict__ status)
{
    __shared__ double t[32][33];
    if (*status != 0) return;
    const double* Pt = P + (long)blockIdx.x * 128 * ldp + (long)blockIdx.x * 128;
    double* Qt = Q + (long)blockIdx.x * 128 * ldp + (long)blockIdx.x * 128;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;          // 32 x 8
    for (int bi = 0; bi < 4; ++bi)
        for (int bj = 0; bj < 4; ++bj) {
            for (int q = 0; q < 4; ++q) t[ly + 8 * q][lx] = Pt[(long)(32 * bi + ly + 8 * q) * ldp + 32 * bj + lx];
            __syncthreads();
            for (int q = 0; q < 4; ++q) Qt[(long)(32 * bj + ly + 8 * q) * ldp + 32 * bi + lx] = t[lx][ly + 8 * q];
            __syncthreads();
        }
}
