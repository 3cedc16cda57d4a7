/* kmat.c — TEST INFRASTRUCTURE (oracle): multi-threaded restatement of george's kernel evaluation
 * for the stationary radial kernels RoBO uses, so that the CPU oracle can cover the full-size
 * BASELINE configurations (2^20 candidates, N = 8192) in seconds instead of minutes.
 *
 * Same arithmetic, in the same order, as oracle/george_oracle.py (which restates george 0.3's
 * Matern52Kernel / Matern32Kernel / ExpSquaredKernel as called from
 * robo/models/gaussian_process.py:106-119 through george.GP.compute / predict):
 *     r2  = sum over the group's axes, in order, of  (x_a - y_a) * (x_a - y_a) / metric_a
 *     f   = (1 + r + 5 r2 / 3) exp(-r), r = sqrt(5 r2)          Matern-5/2
 *           (1 + r) exp(-r),            r = sqrt(3 r2)          Matern-3/2
 *           exp(-r2 / 2)                                         ExpSquared
 *     k   = amp * prod over groups f(r2_g)      (Product of ConstantKernel and radial kernels;
 *           the product is accumulated left to right starting from amp, like Product._value)
 * Checked against george_oracle.get_value in tests/test_oracle_golden.py (<= 2 ulp: libm exp vs numpy exp).
 * Never linked or loaded by the product (robo_b200/); only tests/, bench.py's CPU arms and smoke() use oracle/.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared -o liboracle_kmat.so kmat.c -lm   (oracle/build_c.py)
 */
#include <math.h>
#include <stddef.h>

static double radial(int family, double r2) {
    if (family == 0) { double r = sqrt(5.0 * r2); return (1.0 + r + 5.0 * r2 / 3.0) * exp(-r); }
    if (family == 1) return exp(-0.5 * r2);
    { double r = sqrt(3.0 * r2); return (1.0 + r) * exp(-r); }
}

/* out[i * n2 + j] = k(X1[i], X2[j]); X1 (n1 x d), X2 (n2 x d) row-major.
 * terms t = 0..n_terms-1: axis[t], metric[t] (NOT inverted: the oracle divides), last[t] = 1 closes a group. */
void oracle_kmat(int family, double amp, int n_terms, const int* axis, const int* last, const double* metric,
                 const double* X1, long n1, const double* X2, long n2, int d, double* out)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n1; ++i) {
        const double* x = X1 + i * d;
        double* o = out + i * n2;
        for (long j = 0; j < n2; ++j) {
            const double* y = X2 + j * d;
            double k = amp, r2 = 0.0;
            for (int t = 0; t < n_terms; ++t) {
                const double df = x[axis[t]] - y[axis[t]];
                r2 += df * df / metric[t];
                if (last[t]) { k = k * radial(family, r2); r2 = 0.0; }
            }
            o[j] = k;
        }
    }
}

/* row-wise sum of squares of a column-major-by-candidate block: ssq[j] = sum_i V[i * m + j]^2 (V is n x m) */
void oracle_colsumsq(const double* V, long n, long m, double* ssq)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < m; ++j) {
        double s = 0.0;
        for (long i = 0; i < n; ++i) { const double v = V[i * m + j]; s += v * v; }
        ssq[j] = s;
    }
}

static double radial_dr2(int family, double r2) {
    if (family == 0) { double r = sqrt(5.0 * r2); return -(5.0 / 6.0) * (1.0 + r) * exp(-r); }
    if (family == 1) return -0.5 * exp(-0.5 * r2);
    { double r = sqrt(3.0 * r2); return -1.5 * exp(-r); }
}

/* g[0] = sum_ij A_ij K_ij (d K / d log amp = K), g[1 + t] = sum_ij A_ij dK_ij / d log metric_t with
 *   d f(r2_g) / d log metric_t = -f'(r2_g) (x_t - y_t)^2 / metric_t        (george_oracle._RadialKernel._gradient)
 *   d (k1 k2) = d k1 * k2                                                   (george_oracle.Product._gradient)
 * i.e. the einsum('ijk,ij') of robo/models/gaussian_process.py:186 without materialising the (N, N, H) array
 * (18 GB at N = 8192, D = 32).  X: (n x d) row-major, A: (n x n) row-major.  n_terms <= 64. */
void oracle_grad_trace(int family, double amp, int n_terms, const int* axis, const int* last, const double* metric,
                       const double* X, long n, int d, const double* A, double* g)
{
    for (int p = 0; p <= n_terms; ++p) g[p] = 0.0;
#pragma omp parallel
    {
        double acc[65];
        for (int p = 0; p <= n_terms; ++p) acc[p] = 0.0;
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            const double* x = X + i * d;
            for (long j = 0; j < n; ++j) {
                const double* y = X + j * d;
                double sq[64], fg[64], dfg[64];
                int gid[64], ng = 0;
                double r2 = 0.0;
                for (int t = 0; t < n_terms; ++t) {
                    const double df = x[axis[t]] - y[axis[t]];
                    sq[t] = df * df / metric[t];
                    r2 += sq[t];
                    gid[t] = ng;
                    if (last[t]) { fg[ng] = radial(family, r2); dfg[ng] = radial_dr2(family, r2); ++ng; r2 = 0.0; }
                }
                double k = amp;
                for (int q = 0; q < ng; ++q) k = k * fg[q];
                const double a = A[i * n + j];
                acc[0] += a * k;
                for (int t = 0; t < n_terms; ++t) {
                    double others = amp;
                    for (int q = 0; q < ng; ++q) if (q != gid[t]) others = others * fg[q];
                    acc[1 + t] += a * (-dfg[gid[t]] * sq[t] * others);
                }
            }
        }
#pragma omp critical
        for (int p = 0; p <= n_terms; ++p) g[p] += acc[p];
    }
}
