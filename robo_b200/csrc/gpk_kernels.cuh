// gpk_kernels.cuh — covariance builders, diagonal-block Cholesky/inverse, scoring epilogue.
#pragma once
#include "gpk_internal.cuh"

// ---------------------------------------------------------------------------------------
// Covariance tile builder.  out[c][j] = k(cand_c, train_j)   (row-major, ld = ldo)
//   train points: transposed, column-contiguous  Xt[axis][j]  (coalesced across threads)
//   candidates  : row-major raw inputs, optionally scaled (x - lower) / (upper - lower)
// CTA = 128 train points x 32 candidates, 256 threads, 16 candidates per thread.
// Rows c >= m and columns j >= n are written as exact zeros (padding must not contribute to
// the contractions that follow).  tri != 0: skip tiles entirely above the diagonal (K build).
// Used for K (cand = train), K* (scoring), K** (full_cov) and kernel.get_value.
// ---------------------------------------------------------------------------------------
template <int CPT>
__global__ void __launch_bounds__(256, CPT == 16 ? 1 : 2)
gpk_cov_kernel(const KSpec ks, const double* __restrict__ Xt, long ldx, int n,
               const double* __restrict__ cand, int dc, long m,
               const double* __restrict__ lower, const double* __restrict__ upper,
               double* __restrict__ out, long ldo, int tri)
{
    // CPT candidates per thread: 16 (tile 128 x 32, standalone launches) or 8 (tile 128 x 16, ~60
    // registers so that a CTA fits next to a resident variance-GEMM CTA when the two overlap)
    constexpr int TC = 2 * CPT;
    __shared__ double sc[TC][GPK_MAX_TERMS + 1];
    const int tid = threadIdx.x;
    const int j = blockIdx.x * 128 + (tid & 127);
    const long c0 = (long)blockIdx.y * TC;
    if (tri && (long)blockIdx.x * 128 > c0 + TC - 1) return;

    const int nt = ks.n_terms;
    for (int e = tid; e < TC * nt; e += 256) {
        int c = e / nt, t = e - c * nt;
        long ci = c0 + c;
        double v = 0.0;
        if (ci < m) {
            int a = ks.axis[t];
            v = cand[ci * dc + a];
            if (lower != nullptr) v = (v - lower[a]) / (upper[a] - lower[a]);
        }
        sc[c][t] = v;
    }
    __syncthreads();

    const int cg = (tid >> 7) * CPT;
    const bool jv = j < n;
    double r2[CPT], pr[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) { r2[c] = 0.0; pr[c] = 1.0; }
    for (int t = 0; t < nt; ++t) {
        const double xj = jv ? Xt[(long)ks.axis[t] * ldx + j] : 0.0;
        const double im = ks.inv_metric[t];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            double d = sc[cg + c][t] - xj;
            r2[c] = fma(d * d, im, r2[c]);
        }
        if (ks.last[t]) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) { pr[c] *= gpk_radial(ks.family, r2[c]); r2[c] = 0.0; }
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        long ci = c0 + cg + c;
        out[ci * ldo + j] = (jv && ci < m) ? ks.amp * pr[c] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// Covariance tile builder, TMA-staged (default).  Same contract as gpk_cov_kernel, other operand layout:
//   train side : TERM-major, pre-scaled   Xs[t][j] = x_j[axis_t] * sqrt(c_f / metric_t)   (gpk_termmajor_kernel),
//                one cp.async.bulk.tensor.2d (box n_terms x 128 columns, no swizzle) per CTA into shared memory,
//                completion on an mbarrier; threads read their two train points with one 16-byte LDS per term
//   candidates : row-major raw inputs, scaled (bounds, then the same per-term factor) while filling shared memory
// so the inner loop is  d = s - x ; q = fma(d, d, q)  : 2 FP64 instructions per (pair, term) instead of 3, and one
// broadcast LDS per CC pairs instead of one per pair (the old kernel was co-limited by the LSU: 33 % of the DFMA peak).
// CTA = 128 train points x 4 CC candidates, 256 threads, thread = 2 train points x CC candidates (CC = 8: 128 x 32 tile;
// CC = 4: 128 x 16 tile, <= 64 registers so that a CTA fits next to a resident variance-GEMM CTA).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cov_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CC>
__global__ void __launch_bounds__(256, CC == 8 ? 2 : 4)
gpk_cov_tma_kernel(const __grid_constant__ CUtensorMap mapX, const KSpec ks, int n,
                   const double* __restrict__ cand, int dc, long m,
                   const double* __restrict__ lower, const double* __restrict__ upper,
                   double* __restrict__ out, long ldo, int tri)
{
    constexpr int TC = 4 * CC;
    extern __shared__ unsigned char cov_raw[];
    const int tid = threadIdx.x;
    const long c0 = (long)blockIdx.y * TC;
    if (tri && (long)blockIdx.x * 128 > c0 + TC - 1) return;
    const int nt = ks.n_terms;
    const uint32_t base = (cov_smem_u32(cov_raw) + 127u) & ~127u;
    const uint32_t xs = base;                                   // nt x 128 doubles
    const uint32_t scb = xs + (uint32_t)nt * 1024u;             // TC x nt doubles
    const uint32_t bar = scb + (uint32_t)(TC * nt) * 8u;        // 8-byte aligned mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)nt * 1024u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(xs), "l"((uint64_t)&mapX), "r"(bar), "r"((int)(blockIdx.x * 128)), "r"(0) : "memory");
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
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar) : "memory");
    }
    const int jp = tid & 63, cgp = tid >> 6;
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
    const int j0 = blockIdx.x * 128 + 2 * jp;
    const bool v0 = j0 < n, v1 = j0 + 1 < n;
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        const long ci = c0 + cgp * CC + c;
        const bool cv = ci < m;
        double2 o;
        o.x = (cv && v0) ? ks.amp * pr[c][0] : 0.0;
        o.y = (cv && v1) ? ks.amp * pr[c][1] : 0.0;
        *reinterpret_cast<double2*>(out + ci * ldo + j0) = o;
    }
}

inline size_t cov_tma_smem_bytes(int n_terms, int cc) { return (size_t)n_terms * 1024 + (size_t)4 * cc * n_terms * 8 + 8 + 128; }

// Xs[t][j] = scaled x_j[axis_t] * scale_t  (term-major operand of gpk_cov_tma_kernel), zero padded to ldx columns
__global__ void gpk_termmajor_kernel(const KSpec ks, const double* __restrict__ X, long n, int d,
                                     const double* __restrict__ lower, const double* __restrict__ upper,
                                     double* __restrict__ Xs, long ldx)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)ks.n_terms * ldx;
    if (idx >= total) return;
    int t = (int)(idx / ldx);
    long j = idx - (long)t * ldx;
    double v = 0.0;
    if (j < n) {
        const int a = ks.axis[t];
        v = X[j * d + a];
        if (lower != nullptr) v = (v - lower[a]) / (upper[a] - lower[a]);
        v *= ks.scale[t];
    }
    Xs[idx] = v;
}

// Xt[a][j] = X[j][a] (optionally scaled), zero padded to ldx columns.
__global__ void gpk_transpose_kernel(const double* __restrict__ X, long n, int d,
                                     const double* __restrict__ lower, const double* __restrict__ upper,
                                     double* __restrict__ Xt, long ldx)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)d * ldx;
    if (idx >= total) return;
    int a = (int)(idx / ldx);
    long j = idx - (long)a * ldx;
    double v = 0.0;
    if (j < n) {
        v = X[j * d + a];
        if (lower != nullptr) v = (v - lower[a]) / (upper[a] - lower[a]);
    }
    Xt[idx] = v;
}

// After the K build: diagonal += diag_add (george: yerr^2 + TINY), unit diagonal on padding rows,
// and the augmented right-hand-side row  K[NP][j] = y_j - mean  (so the blocked Cholesky also
// produces z = L^-1 (y - mean) as row NP of the factor; rows NP+1.. stay zero).
__global__ void gpk_kfix_kernel(double* __restrict__ K, long ld, int n, int NP, double diag_add,
                                const double* __restrict__ y, double mean)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP) return;
    K[(long)i * ld + i] = (i < n) ? K[(long)i * ld + i] + diag_add : 1.0;
    K[(long)NP * ld + i] = (i < n) ? y[i] - mean : 0.0;
}

// ---------------------------------------------------------------------------------------
// Diagonal block: L_kk = chol(A_kk) in shared memory, then inv(L_kk) in place.
//   writes L_kk (lower, upper zeroed) back to K,
//   inv(L_kk) to the diagonal block of P (lower) and its transpose to Q (upper),
//   sum log diag(L_kk) to logdet_part[kb], and flags a non-positive pivot in *status
//   (1 + global pivot index), which is what scipy.linalg.cholesky reports as LinAlgError.
// One CTA, 256 threads, dynamic smem 128 x 129 doubles + 3 x 128.
// ---------------------------------------------------------------------------------------
constexpr int DIAG_SMEM = (128 * 129 + 2 * 128) * 8;

__global__ void __launch_bounds__(256)
gpk_potrf_diag_kernel(double* __restrict__ K, long ld, int kb,
                      double* __restrict__ P, double* __restrict__ Q, long ldp,
                      int* __restrict__ status, double* __restrict__ logdet_part)
{
    extern __shared__ double dsm[];
    double (*A)[129] = (double (*)[129])dsm;
    double* dsq = dsm + 128 * 129;       // sqrt of pivots
    double* vbuf = dsq + 128;
    __shared__ double pb2[256];
    __shared__ int s_bad;

    const int tid = threadIdx.x;
    if (*status != 0) return;
    if (tid == 0) s_bad = 0;

    double* Kt = K + (long)kb * 128 * ld + (long)kb * 128;
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        A[r][c] = Kt[(long)r * ld + c];
    }
    __syncthreads();

    const int i = tid & 127, h = tid >> 7;
    // ---- elimination: after step j, column j holds the un-scaled column, A[j][j] = pivot d_j
    for (int j = 0; j < 128; ++j) {
        double d = A[j][j];
        if (!(d > 0.0) || isinf(d)) {
            if (tid == 0 && s_bad == 0) s_bad = kb * 128 + j + 1;
            d = 1.0;
        }
        const double invd = 1.0 / d;
        if (i > j) {
            const double lij = A[i][j] * invd;
            for (int c = j + 1 + h; c <= i; c += 2) A[i][c] -= lij * A[c][j];
        }
        __syncthreads();
    }
    if (tid < 128) {
        double d = A[tid][tid];
        if (!(d > 0.0) || isinf(d)) d = 1.0;
        dsq[tid] = sqrt(d);
    }
    __syncthreads();
    // ---- finalise L and write it back
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        double v = 0.0;
        if (c < r) { v = A[r][c] / dsq[c]; }
        else if (c == r) v = dsq[c];
        Kt[(long)r * ld + c] = v;
    }
    __syncthreads();
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        if (c < r) A[r][c] = A[r][c] / dsq[c];
    }
    if (tid < 32) {
        double s = 0.0;
        for (int q = tid; q < 128; q += 32) s += log(dsq[q]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (tid == 0) logdet_part[kb] = s;
    }
    __syncthreads();
    if (tid == 0 && s_bad != 0) atomicCAS(status, 0, s_bad);

    // ---- in-place inverse of the lower-triangular block, right to left (LAPACK dtrti2 order)
    for (int j = 127; j >= 0; --j) {
        const double ajj = 1.0 / dsq[j];
        if (tid < 128 && tid > j) vbuf[tid] = A[tid][j];
        __syncthreads();
        double part = 0.0;
        if (i > j) {
            for (int c = j + 1 + h; c <= i; c += 2) part = fma(A[i][c], vbuf[c], part);
        }
        pb2[tid] = part;
        __syncthreads();
        if (h == 0 && i > j) A[i][j] = -(pb2[i] + pb2[128 + i]) * ajj;
        if (tid == 0) A[j][j] = ajj;
    }
    __syncthreads();
    double* Pt = P + (long)kb * 128 * ldp + (long)kb * 128;
    double* Qt = Q + (long)kb * 128 * ldp + (long)kb * 128;
    for (int e = tid; e < 128 * 128; e += 256) {
        int r = e >> 7, c = e & 127;
        Pt[(long)r * ldp + c] = (c <= r) ? A[r][c] : 0.0;
        Qt[(long)r * ldp + c] = (c >= r) ? A[c][r] : 0.0;
    }
}

constexpr int DIAG2_SMEM = (128 * 129 + 5 * 128) * 8;

// ---------------------------------------------------------------------------------------
// Diagonal block, register-tiled fused version (default; same contract as gpk_potrf_diag_kernel).
// Thread (ty, tx) = (tid / 16, tid % 16) keeps the 8 x 8 cyclic sub-tile A[ty + 16a][tx + 16b] in registers; per
// column j the 16 owner threads publish the column through a double-buffered shared vector, everyone
// scales it by rsqrt(pivot) and applies the rank-1 update to its registers.  The forward substitution L X = I runs one column behind the
// factorisation inside the SAME barrier interval (row j-1 of X and column j of A are published
// before the one __syncthreads of step j), so the block costs 128 barrier intervals instead of 256.
// L columns reach the substitution through registers (the scaled column every thread already
// computed for the rank-1 update), not through shared memory.
// ---------------------------------------------------------------------------------------
template <int JBP>
__device__ __forceinline__ void diag_x_publish(double (&X)[8][8], double* rb, double rs_prev, int ty, int tx, int jjp)
{
    if (ty == jjp) {              // owners of row jm = 16*JBP + jjp finish it: X[jm][c] *= 1/L[jm][jm]
#pragma unroll
        for (int b = 0; b <= JBP; ++b) {
            const double x = X[JBP][b] * rs_prev;
            X[JBP][b] = x;
            rb[tx + 16 * b] = x;
        }
    }
}

template <int JBP>
__device__ __forceinline__ void diag_x_update(double (&X)[8][8], const double (&lr)[8], const double* rb,
                                              int ty, int tx, int jjp)
{
    // X[i][c] -= L[i][jm] X[jm][c] for the rows below jm = 16 JBP + jjp.  No per-element predicates: the rows of
    // block JBP that are not below jm get a zero multiplier (one select), columns right of jm hold zeros in rb.
    double xr[8], lx[8];
    const double* rbc = rb + tx;
#pragma unroll
    for (int b = 0; b <= JBP; ++b) xr[b] = rbc[16 * b];
#pragma unroll
    for (int a = JBP; a < 8; ++a) lx[a] = lr[a];
    lx[JBP] = (ty > jjp) ? lr[JBP] : 0.0;
#pragma unroll
    for (int a = JBP; a < 8; ++a)
#pragma unroll
        for (int b = 0; b <= JBP; ++b) X[a][b] = fma(-lx[a], xr[b], X[a][b]);
}

template <int JB>
__device__ __forceinline__ void diag_fused_block(double (&A)[8][8], double (&X)[8][8], double (&lr)[8], double& rs_prev,
                                                 double& rs_cur, double* colbuf, double* rowbuf, double* dnext,
                                                 int ty, int tx, int tid, int kb, int* s_bad)
{
    // Invariants.  lr[] enters holding the scaled column j-1 of L (this thread's rows) and leaves holding column j.
    // rs_cur = rsqrt(pivot j) was computed during the previous step from the published diagonal element.
    // Entries above the diagonal inside the diagonal blocks (row < column) are never read by valid entries and
    // are left to hold garbage (this removes the per-element predicates; the final store masks them).
    for (int jj = 0; jj < 16; ++jj) {
        const int j = JB * 16 + jj;
        double* cb = colbuf + (j & 1) * 128;
        double* rb = rowbuf + ((j + 1) & 1) * 128;       // parity of jm = j - 1
        if (tx == jj) {
            double* cbw = cb + ty;
#pragma unroll
            for (int a = JB; a < 8; ++a) cbw[16 * a] = A[a][JB];
        }
        if (jj < 15) {
            if (ty == jj + 1 && tx == jj + 1) dnext[j & 1] = A[JB][JB];
        } else if (JB < 7) {
            if (ty == 0 && tx == 0) dnext[j & 1] = A[JB < 7 ? JB + 1 : 7][JB < 7 ? JB + 1 : 7];
        }
        if (jj > 0) diag_x_publish<JB>(X, rb, rs_prev, ty, tx, jj - 1);
        else if (JB > 0) diag_x_publish<(JB > 0 ? JB - 1 : 0)>(X, rb, rs_prev, ty, tx, 15);
        __syncthreads();
        double d = cb[j];
        if (!(d > 0.0) || isinf(d)) {
            if (tid == 0 && *s_bad == 0) *s_bad = kb * 128 + j + 1;
            d = 1.0;
        }
        const double rs = rs_cur;            // == rsqrt(d)
        const double sq = d * rs;
        double rs_next = 1.0;
        if (j < 127) {                       // next pivot: the same fma its owner applies below
            const double ln = cb[j + 1] * rs;
            const double dn = fma(-ln, ln, dnext[j & 1]);
            rs_next = (dn > 0.0 && !isinf(dn)) ? rsqrt(dn) : 1.0;
        }
        // forward-substitution step for row j-1, with the previous column of L still in lr[]
        if (jj > 0) diag_x_update<JB>(X, lr, rb, ty, tx, jj - 1);
        else if (JB > 0) diag_x_update<(JB > 0 ? JB - 1 : 0)>(X, lr, rb, ty, tx, 15);
        // column j of L, scaled
        double lc[8];
        const double* cbr = cb + ty;
        const double* cbc = cb + tx;
#pragma unroll
        for (int a = JB; a < 8; ++a) lr[a] = cbr[16 * a] * rs;
#pragma unroll
        for (int b = JB; b < 8; ++b) lc[b] = cbc[16 * b] * rs;
        if (tx == jj) {                      // owners keep the finished column (rows above the diagonal: don't care)
#pragma unroll
            for (int a = JB + 1; a < 8; ++a) A[a][JB] = lr[a];
            A[JB][JB] = (ty == jj) ? sq : lr[JB];
        }
        // rank-1 update; only the column block that contains finished columns needs a predicate
        if (tx > jj) {
#pragma unroll
            for (int a = JB; a < 8; ++a) A[a][JB] = fma(-lr[a], lc[JB], A[a][JB]);
        }
#pragma unroll
        for (int b = JB + 1; b < 8; ++b)
#pragma unroll
            for (int a = b; a < 8; ++a) A[a][b] = fma(-lr[a], lc[b], A[a][b]);
        rs_prev = rs;
        rs_cur = rs_next;
    }
}

__global__ void __launch_bounds__(256, 1)
gpk_potrf_diag_fused_kernel(double* __restrict__ K, long ld, int kb,
                            double* __restrict__ P, double* __restrict__ Q, long ldp,
                            int* __restrict__ status, double* __restrict__ logdet_part)
{
    extern __shared__ double dsm[];
    double (*Ls)[129] = (double (*)[129])dsm;
    double* colbuf = dsm + 128 * 129;     // 2 x 128
    double* rowbuf = colbuf + 256;        // 2 x 128
    double* dnext = rowbuf + 256;         // 2: diagonal element of the next pivot
    __shared__ int s_bad;

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    cudaGridDependencySynchronize();      // programmatic dependent launch (see gpk_gemm_nt_kernel)
    if (*status != 0) return;
    if (tid == 0) s_bad = 0;

    double* Kt = K + (long)kb * 128 * ld + (long)kb * 128;
    double A[8][8], X[8][8], lrp[8];          // lrp: the scaled column of L of the previous step
    double rs_prev = 1.0, rs_cur = 1.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        lrp[a] = 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ty + 16 * a, c = tx + 16 * b;
            A[a][b] = (c <= i) ? Kt[(long)i * ld + c] : 0.0;
            X[a][b] = (i == c) ? 1.0 : 0.0;
        }
    }
    if (tid == 0) dnext[1] = A[0][0];     // first pivot (slot 1: step 0 publishes the next one into slot 0)
    __syncthreads();
    {
        const double d0 = dnext[1];
        rs_cur = (d0 > 0.0 && !isinf(d0)) ? rsqrt(d0) : 1.0;
    }
    diag_fused_block<0>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<1>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<2>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<3>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<4>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<5>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<6>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    diag_fused_block<7>(A, X, lrp, rs_prev, rs_cur, colbuf, rowbuf, dnext, ty, tx, tid, kb, &s_bad);
    // last row of X (row 127) only needs its scaling; nobody reads the broadcast copy, but slower threads may still be
    // reading this buffer for the update of step 127 (racecheck): write the copy to the other parity buffer
    diag_x_publish<7>(X, rowbuf + 128, rs_prev, ty, tx, 15);

    // ---------------- publish L, log-det, status ----------------
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ty + 16 * a, c = tx + 16 * b;
            const double v = (c <= i) ? A[a][b] : 0.0;
            if (i == c) colbuf[i] = v;                     // diagonal of L for the log-det
            Kt[(long)i * ld + c] = v;
        }
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int q = tid; q < 128; q += 32) s += log(colbuf[q]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (tid == 0) {
            logdet_part[kb] = s;
            if (s_bad != 0) atomicCAS(status, 0, s_bad);
        }
    }
    // ---------------- publish L^-1 (P lower) and its transpose (Q upper) ----------------
    double* Pt = P + (long)kb * 128 * ldp + (long)kb * 128;
    double* Qt = Q + (long)kb * 128 * ldp + (long)kb * 128;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ty + 16 * a, c = tx + 16 * b;
            const double v = (c <= i) ? X[a][b] : 0.0;
            Ls[i][c] = v;
            Pt[(long)i * ldp + c] = v;
        }
    __syncthreads();
    for (int e = tid; e < 128 * 128; e += 256) {
        const int r = e >> 7, c = e & 127;
        Qt[(long)r * ldp + c] = (c >= r) ? Ls[c][r] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// Scoring epilogue: sum the per-row-block partials in fixed order, finish mean / variance,
// apply the output transform + clip (gaussian_process.py:282-294), the acquisition closed form,
// and a per-block arg-max with numpy.argmax tie-breaking.
// ---------------------------------------------------------------------------------------
struct FinishArgs {
    const double* part_mu; const double* part_ssq; long ldpart; int nparts;
    long m;                 // valid candidates in this chunk
    long base;              // global index of the chunk's first candidate
    double kss;             // k(x*, x*) = amplitude (stationary kernels)
    double mean;            // GP constant mean
    int norm_out; double y_mean, y_std;
    int acq_kind; double eta, par;
    double* out_mu; double* out_var; double* out_acq;    // chunk-local device arrays, may be NULL
    BestPair* block_best;   // one per block
    unsigned long long* n_negative;
    const double* mu_direct; // when set: mean - f.mean = K* alpha computed in fp64 beside the int8 variance contraction
};

__global__ void __launch_bounds__(256) gpk_finish_kernel(const FinishArgs f)
{
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    double val = 0.0;
    long long idx = -1;
    if (c < f.m) {
        double ssq = 0.0, mu = 0.0;
        if (f.mu_direct != nullptr) {
            for (int p = 0; p < f.nparts; ++p) ssq += f.part_ssq[(long)p * f.ldpart + c];
            mu = f.mu_direct[c];
        } else {
            for (int p = 0; p < f.nparts; ++p) {
                ssq += f.part_ssq[(long)p * f.ldpart + c];
                mu += f.part_mu[(long)p * f.ldpart + c];
            }
        }
        double var = f.kss - ssq;
        mu += f.mean;
        if (f.norm_out) { mu = mu * f.y_std + f.y_mean; var = var * (f.y_std * f.y_std); }
        if (var < GPK_EPS) var = GPK_EPS;                  // np.clip(var, eps, inf); NaN stays NaN
        if (f.out_mu) f.out_mu[c] = mu;
        if (f.out_var) f.out_var[c] = var;
        if (f.acq_kind != GPK_ACQ_NONE) {
            val = gpk_acq_value(f.acq_kind, mu, var, f.eta, f.par);
            if (f.out_acq) f.out_acq[c] = val;
            if (f.acq_kind == GPK_ACQ_EI && val < 0.0 && f.n_negative) atomicAdd(f.n_negative, 1ULL);
            idx = f.base + c;
        }
    }
    if (f.acq_kind == GPK_ACQ_NONE || f.block_best == nullptr) return;
    // block arg-max
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, val, off);
        long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (gpk_better(ov, oi, val, idx)) { val = ov; idx = oi; }
    }
    __shared__ double sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = val; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (gpk_better(sv[w], si[w], val, idx)) { val = sv[w]; idx = si[w]; }
        f.block_best[blockIdx.x].val = val;
        f.block_best[blockIdx.x].idx = idx;
    }
}

// Final arg-max over block results, merged into *best (which may hold the running best of
// earlier chunks; idx < 0 means empty).
__global__ void __launch_bounds__(256) gpk_argmax_final_kernel(const BestPair* __restrict__ bb, int nblocks,
                                                               BestPair* __restrict__ best)
{
    double val = 0.0;
    long long idx = -1;
    for (int b = threadIdx.x; b < nblocks; b += 256)
        if (gpk_better(bb[b].val, bb[b].idx, val, idx)) { val = bb[b].val; idx = bb[b].idx; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, val, off);
        long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (gpk_better(ov, oi, val, idx)) { val = ov; idx = oi; }
    }
    __shared__ double sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = val; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (gpk_better(sv[w], si[w], val, idx)) { val = sv[w]; idx = si[w]; }
        if (gpk_better(val, idx, best->val, best->idx)) { best->val = val; best->idx = idx; }
    }
}

// Acquisition closed form on supplied moments (gpk_acq_moments).
__global__ void gpk_acq_moments_kernel(const double* __restrict__ mu, const double* __restrict__ var, long m,
                                       int kind, double eta, double par, double* __restrict__ out,
                                       unsigned long long* n_negative)
{
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double v = gpk_acq_value(kind, mu[c], var[c], eta, par);
    out[c] = v;
    if (kind == GPK_ACQ_EI && v < 0.0) atomicAdd(n_negative, 1ULL);
}

// full_cov epilogue: output transform, then (clip != 0) every entry clipped to >= eps
// (gaussian_process.py:282-294 applies the clip to the whole matrix).  clip == 0 keeps the raw posterior
// covariance, negative off-diagonal entries included: what george's sample_conditional draws from
// (gaussian_process.py:324; only predict() clips).
__global__ void gpk_cov_finish_kernel(double* __restrict__ cov, long ld, long m, int norm_out, double y_std, int clip)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    long r = idx / m, c = idx - r * m;
    double v = cov[r * ld + c];
    if (norm_out) v *= y_std * y_std;
    if (clip && v < GPK_EPS) v = GPK_EPS;
    cov[r * ld + c] = v;
}


// ---------------------------------------------------------------------------------------
// Marginal-likelihood gradient, trace pass (gaussian_process.py:168-191 with the noise term
// corrected):  d(-ll)/d theta_p = -1/2 sum_ij A_ij dK_ij/d theta_p,  A = alpha alpha^T - K^-1.
// Same tiling as the covariance builder (128 columns j x 32 rows i per CTA, lower tiles only,
// off-diagonal pairs counted twice); dK/d theta is recomputed on the fly and never stored
// (the reference materialises an (N, N, H) array, :181-182).  Per CTA it writes nv = n_terms + 2
// partial sums: [sum w k, sum_t ..., trace A]; a second kernel adds the CTAs in fixed order.
//   dk/d log_amp      = k
//   dk/d log_metric_t = -k * (dlog f / d r2)(r2_g) * (x_t - x'_t)^2 / metric_t      (t in group g)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gpk_grad_trace_kernel(const KSpec ks, const double* __restrict__ Xt, long ldx, int n,
                      const double* __restrict__ Xrow, int dc,
                      const double* __restrict__ Kinv, long ldk, const double* __restrict__ alpha,
                      double* __restrict__ part)
{
    __shared__ double sc[32][GPK_MAX_TERMS + 1];
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int nt = ks.n_terms, nv = nt + 2;
    const long bid = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const int j = blockIdx.x * 128 + (tid & 127);
    const long c0 = (long)blockIdx.y * 32;
    if ((long)blockIdx.x * 128 > c0 + 31) {                  // tile entirely above the diagonal
        if (tid < nv) part[bid * nv + tid] = 0.0;
        return;
    }
    for (int e = tid; e < 32 * nt; e += 256) {
        int c = e / nt, t = e - c * nt;
        long ci = c0 + c;
        sc[c][t] = (ci < n) ? Xrow[ci * dc + ks.axis[t]] : 0.0;
    }
    __syncthreads();

    const int cg = (tid >> 7) * 16;
    const bool jv = j < n;
    double gl[GPK_MAX_TERMS];                                // per-term partial sums (local memory)
    for (int t = 0; t < nt; ++t) gl[t] = 0.0;
    double gamp = 0.0, gtr = 0.0;

    double r2[16], wk[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { r2[c] = 0.0; wk[c] = 1.0; }
    for (int t = 0; t < nt; ++t) {                           // pass 1: k(x_i, x_j) / amp
        const double xj = jv ? Xt[(long)ks.axis[t] * ldx + j] : 0.0;
        const double im = ks.inv_metric[t];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double d = sc[cg + c][t] - xj;
            r2[c] = fma(d * d, im, r2[c]);
        }
        if (ks.last[t]) {
#pragma unroll
            for (int c = 0; c < 16; ++c) { wk[c] *= gpk_radial(ks.family, r2[c]); r2[c] = 0.0; }
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {                           // weights w_ij * k_ij
        const long i = c0 + cg + c;
        double w = 0.0;
        if (jv && i < n && j <= i) {
            const double a = alpha[i] * alpha[j] - Kinv[i * ldk + j];
            if (i == j) { gtr += a; w = a; } else w = 2.0 * a;
        }
        wk[c] = w * ks.amp * wk[c];
        gamp += wk[c];
    }
    int t0 = 0;
    while (t0 < nt) {                                        // pass 2: group by group
        int t1 = t0;
        while (!ks.last[t1]) ++t1;
#pragma unroll
        for (int c = 0; c < 16; ++c) r2[c] = 0.0;
        for (int t = t0; t <= t1; ++t) {
            const double xj = jv ? Xt[(long)ks.axis[t] * ldx + j] : 0.0;
            const double im = ks.inv_metric[t];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double d = sc[cg + c][t] - xj;
                r2[c] = fma(d * d, im, r2[c]);
            }
        }
        double coef[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) coef[c] = -wk[c] * gpk_radial_dlog(ks.family, r2[c]);
        for (int t = t0; t <= t1; ++t) {
            const double xj = jv ? Xt[(long)ks.axis[t] * ldx + j] : 0.0;
            double sacc = 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double d = sc[cg + c][t] - xj;
                sacc = fma(coef[c], d * d, sacc);
            }
            gl[t] += sacc * ks.inv_metric[t];
        }
        t0 = t1 + 1;
    }
    // block reduction of the nv values (fixed order: lanes, then warps)
    for (int v = 0; v < nv; ++v) {
        double x = (v == 0) ? gamp : (v == nv - 1 ? gtr : gl[v - 1]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((tid & 31) == 0) red[tid >> 5] = x;
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < 8; ++w) sum += red[w];
            part[bid * nv + v] = sum;
        }
        __syncthreads();
    }
}

// out[v] = -1/2 * sum over CTAs (fixed order); the noise entry is scaled by sigma^2.
__global__ void gpk_grad_final_kernel(const double* __restrict__ part, long nblocks, int nv, double noise_var,
                                      double* __restrict__ out)
{
    const int v = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (long b = threadIdx.x; b < nblocks; b += 256) s += part[b * nv + v];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = -0.5 * sh[0] * (v == nv - 1 ? noise_var : 1.0);
}

// ---------------------------------------------------------------------------------------
// Reductions over the hyper-parameter samples of a GP-MCMC model (A, B are [n_models][m]):
//   mode 0: out1 = mean_i A_i                       (MarginalizationGPMCMC.compute, marginalization.py:121)
//   mode 1: out1 = mean_i A_i ; out2 = var_i(A_i) + mean_i B_i, clipped at eps
//           (GaussianProcessMCMC.predict, gaussian_process_mcmc.py:235-247; np.var is the
//            two-pass population variance)
// ---------------------------------------------------------------------------------------
__global__ void gpk_reduce_models_kernel(const double* __restrict__ A, const double* __restrict__ B, int n_models,
                                         long m, int mode, double* __restrict__ out1, double* __restrict__ out2)
{
    long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int i = 0; i < n_models; ++i) s += A[(long)i * m + c];
    const double mean = s / (double)n_models;
    out1[c] = mean;
    if (mode == 1) {
        double v = 0.0, sb = 0.0;
        for (int i = 0; i < n_models; ++i) {
            double d = A[(long)i * m + c] - mean;
            v += d * d;
            sb += B[(long)i * m + c];
        }
        double r = v / (double)n_models + sb / (double)n_models;
        if (r < GPK_EPS) r = GPK_EPS;
        out2[c] = r;
    }
}

// ---------------------------------------------------------------------------------------
// On-device candidate generation for RandomSampling.maximize (random_sampling.py:38-47):
//   candidate i < n_uniform : lower + (upper - lower) * U[0,1)^d          (init_random_uniform)
//   otherwise               : clip(incumbent + scale * N(0,1)^d, lower, upper)
// Philox4x32-10 counter-based generator keyed by (seed, global candidate index, coordinate pair):
// the stream depends on nothing else, so results are identical for any chunking or GPU count.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void gpk_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double gpk_u01(uint32_t lo, uint32_t hi) {      // 53-bit uniform in [0, 1)
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (double)(v >> 11) * 1.1102230246251565e-16;
}

__global__ void gpk_candidates_kernel(unsigned long long seed, long first, long count, long n_uniform, int d,
                                      const double* __restrict__ lower, const double* __restrict__ upper,
                                      const double* __restrict__ incumbent, double scale,
                                      double* __restrict__ out)
{
    const int npair = (d + 1) / 2;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * npair) return;
    const long li = t / npair;
    const int pb = (int)(t - li * npair);
    const unsigned long long gi = (unsigned long long)(first + li);
    uint32_t r[4];
    gpk_philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)pb, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u0 = gpk_u01(r[0], r[1]), u1 = gpk_u01(r[2], r[3]);
    double v0, v1;
    const int a0 = 2 * pb, a1 = 2 * pb + 1;
    if ((long)gi < n_uniform) {
        // explicit rounding steps (no fma contraction): bit-identical to numpy's lower + (upper-lower)*u
        v0 = __dadd_rn(lower[a0], __dmul_rn(upper[a0] - lower[a0], u0));
        if (a1 < d) v1 = __dadd_rn(lower[a1], __dmul_rn(upper[a1] - lower[a1], u1));
    } else {                                              // Box-Muller, u in (0, 1]
        const double rad = sqrt(-2.0 * log(1.0 - u0));
        double sn, cs;
        sincospi(2.0 * u1, &sn, &cs);
        v0 = fmin(fmax(incumbent[a0] + scale * rad * cs, lower[a0]), upper[a0]);
        if (a1 < d) v1 = fmin(fmax(incumbent[a1] + scale * rad * sn, lower[a1]), upper[a1]);
    }
    out[li * d + a0] = v0;
    if (a1 < d) out[li * d + a1] = v1;
}

// ---------------------------------------------------------------------------------------
// Predictive gradients (the API robo/acquisition_functions/ei.py:80-85,
// pi.py:65-71, lcb.py:66-69 expect from a model but no reference model implements):
//   d mu / d x*_a   =  sum_j alpha_j        d k(x*, x_j) / d x*_a
//   d var / d x*_a  = -2 sum_j (K^-1 k*)_j  d k(x*, x_j) / d x*_a          (k** is constant: stationary kernels)
//   d k / d x*_t    =  k * (dlog f / d r2)(r2_g) * 2 (x*_t - x_jt) / metric_t    per kernel term t (axis a = axis[t])
// One CTA per candidate, threads stride over the training points, block reduction per term, chain
// rule of the input scaling (1 / (upper - lower)) and of the output transform applied at the end.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gpk_predict_grad_kernel(const KSpec ks, const double* __restrict__ Xt, long ldx, int n,
                        const double* __restrict__ cand, int dc,
                        const double* __restrict__ lower, const double* __restrict__ upper,
                        const double* __restrict__ alpha, const double* __restrict__ Wt, long ldw,
                        int norm_out, double y_std, double* __restrict__ dmu, double* __restrict__ dvar)
{
    __shared__ double xs[GPK_MAX_TERMS];
    __shared__ double red[16];
    const int tid = threadIdx.x, nt = ks.n_terms;
    const long c = blockIdx.x;
    for (int t = tid; t < nt; t += 256) {
        const int a = ks.axis[t];
        double v = cand[c * dc + a];
        if (lower != nullptr) v = (v - lower[a]) / (upper[a] - lower[a]);
        xs[t] = v;
    }
    for (int a = tid; a < dc; a += 256) { dmu[c * dc + a] = 0.0; dvar[c * dc + a] = 0.0; }
    __syncthreads();
    double gm[GPK_MAX_TERMS], gv[GPK_MAX_TERMS];
    for (int t = 0; t < nt; ++t) { gm[t] = 0.0; gv[t] = 0.0; }
    for (int j = tid; j < n; j += 256) {
        double k = ks.amp, r2 = 0.0;
        for (int t = 0; t < nt; ++t) {                                  // kernel value
            const double d = xs[t] - Xt[(long)ks.axis[t] * ldx + j];
            r2 = fma(d * d, ks.inv_metric[t], r2);
            if (ks.last[t]) { k *= gpk_radial(ks.family, r2); r2 = 0.0; }
        }
        const double ka = k * alpha[j], kw = -2.0 * k * Wt[c * ldw + j];
        int t0 = 0;
        while (t0 < nt) {                                               // group by group
            int t1 = t0;
            while (!ks.last[t1]) ++t1;
            r2 = 0.0;
            for (int t = t0; t <= t1; ++t) {
                const double d = xs[t] - Xt[(long)ks.axis[t] * ldx + j];
                r2 = fma(d * d, ks.inv_metric[t], r2);
            }
            const double ratio = 2.0 * gpk_radial_dlog(ks.family, r2);
            for (int t = t0; t <= t1; ++t) {
                const double d = xs[t] - Xt[(long)ks.axis[t] * ldx + j];
                const double f = ratio * d * ks.inv_metric[t];
                gm[t] = fma(ka, f, gm[t]);
                gv[t] = fma(kw, f, gv[t]);
            }
            t0 = t1 + 1;
        }
    }
    for (int t = 0; t < nt; ++t) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            double x = pass == 0 ? gm[t] : gv[t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if ((tid & 31) == 0) red[pass * 8 + (tid >> 5)] = x;
        }
        __syncthreads();
        if (tid == 0) {
            double sm = 0.0, sv = 0.0;
            for (int w = 0; w < 8; ++w) { sm += red[w]; sv += red[8 + w]; }
            const int a = ks.axis[t];
            double scale = 1.0;
            if (lower != nullptr) scale = 1.0 / (upper[a] - lower[a]);
            if (norm_out) { sm *= y_std; sv *= y_std * y_std; }
            dmu[c * dc + a] += sm * scale;
            dvar[c * dc + a] += sv * scale;
        }
        __syncthreads();
    }
}

// Acquisition value and gradient from moments and their gradients (ei.py:76-85, pi.py:61-71, lcb.py:65-69).
__global__ void gpk_acq_grad_kernel(const double* __restrict__ mu, const double* __restrict__ var,
                                    const double* __restrict__ dmu, const double* __restrict__ dvar, long m, int d,
                                    int kind, double eta, double par, double* __restrict__ f, double* __restrict__ df)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * d) return;
    const long c = idx / d;
    const double s = sqrt(var[c]);
    const double dm = dmu[idx], ds = dvar[idx] / (2.0 * s);
    double g;
    if (kind == GPK_ACQ_EI) {
        const double z = (eta - mu[c] - par) / s;
        g = -dm * gpk_ndtr(z) + ds * gpk_norm_pdf(z);
    } else if (kind == GPK_ACQ_PI) {
        const double z = (eta - mu[c] - par) / s;
        g = -(gpk_norm_pdf(z) / s) * (dm + ds * z);
    } else {
        g = -(dm - par * ds);
    }
    df[idx] = g;
    if (idx == c * d) f[c] = gpk_acq_value(kind, mu[c], var[c], eta, par);
}

// ---------------------------------------------------------------------------------------
// fp64 issue-rate peaks of the GPU we run on (register-resident operands, no memory traffic): the
// denominators of the roofline reported by bench.py.  DMMA m8n8k4 = the tensor pipe every GEMM of this
// library uses; DFMA = the vector pipe of the covariance builder.
// ---------------------------------------------------------------------------------------
__global__ void gpk_peak_dmma_kernel(double* out, int iters)
{
    double c[16][2], a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void gpk_peak_dfma_kernel(double* out, int iters)
{
    double a[8];
    const double b = 1.0000001, c = 0.9999999;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}
