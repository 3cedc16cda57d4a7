// gpk_internal.cuh — shared device helpers and structs (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/gpk.h"

#define GPK_TILE 128          // block size of every blocked algorithm (rows per tile)
#define GPK_EPS 2.220446049250313e-16

// Kernel specification passed by value to the covariance-building kernels.
// k(x,x') = amp * prod_g f( sum_{t in g} (x[axis_t]-x'[axis_t])^2 * inv_metric_t )
struct KSpec {
    int family;
    int n_terms;
    double amp;
    int axis[GPK_MAX_TERMS];
    int last[GPK_MAX_TERMS];        // 1 if term t closes its product group
    double inv_metric[GPK_MAX_TERMS];
    double scale[GPK_MAX_TERMS];    // sqrt(c_f / metric_t): coordinates pre-scaled so that q = sum (s - s')^2 is the
                                    // radial argument directly (c_f = 5 Matern-5/2, 3 Matern-3/2, 1/2 ExpSquared)
};

// f as a function of q = c_f * r2 (pre-scaled coordinates, gpk_cov_tma_kernel)
__device__ __forceinline__ double gpk_radial_q(int family, double q) {
    if (family == GPK_MATERN52) {
        double r = sqrt(q);
        return fma(q, 1.0 / 3.0, 1.0 + r) * exp(-r);           // 1 + r + 5 r2 / 3
    } else if (family == GPK_EXPSQUARED) {
        return exp(-q);
    } else {
        double r = sqrt(q);
        return (1.0 + r) * exp(-r);
    }
}

// f(r2) for the radial families (oracle/george_oracle.py: Matern52Kernel._f etc.).
__device__ __forceinline__ double gpk_radial(int family, double r2) {
    if (family == GPK_MATERN52) {
        double r = sqrt(5.0 * r2);
        return (1.0 + r + 5.0 * r2 / 3.0) * exp(-r);
    } else if (family == GPK_EXPSQUARED) {
        return exp(-0.5 * r2);
    } else {
        double r = sqrt(3.0 * r2);
        return (1.0 + r) * exp(-r);
    }
}

// d f / d r2 (george_oracle.py: _dfdr2)
__device__ __forceinline__ double gpk_radial_dr2(int family, double r2) {
    if (family == GPK_MATERN52) {
        double r = sqrt(5.0 * r2);
        return -(5.0 / 6.0) * (1.0 + r) * exp(-r);
    } else if (family == GPK_EXPSQUARED) {
        return -0.5 * exp(-0.5 * r2);
    } else {
        double r = sqrt(3.0 * r2);
        return -1.5 * exp(-r);
    }
}

// d log f / d r2 (ratio f'/f in closed form: no 0/0 when f underflows)
__device__ __forceinline__ double gpk_radial_dlog(int family, double r2) {
    if (family == GPK_MATERN52) {
        double r = sqrt(5.0 * r2);
        return -(5.0 / 6.0) * (1.0 + r) / (1.0 + r + 5.0 * r2 / 3.0);
    } else if (family == GPK_EXPSQUARED) {
        return -0.5;
    } else {
        double r = sqrt(3.0 * r2);
        return -1.5 / (1.0 + r);
    }
}

// ---- standard normal helpers (scipy.special.ndtr / log_ndtr / norm.pdf restated) -------
__device__ __forceinline__ double gpk_ndtr(double z) {
    return 0.5 * erfc(-z * 0.70710678118654752440);
}
__device__ __forceinline__ double gpk_norm_pdf(double z) {
    return exp(-0.5 * z * z) * 0.39894228040143267794;       // 1/sqrt(2 pi)
}
__device__ __forceinline__ double gpk_norm_logpdf(double z) {
    return -0.5 * z * z - 0.91893853320467274178;             // log sqrt(2 pi)
}
__device__ __forceinline__ double gpk_log_ndtr(double z) {
    double t = z * 0.70710678118654752440;
    if (z < -1.0) return log(erfcx(-t) * 0.5) - t * t;
    return log1p(-0.5 * erfc(t));
}

// Acquisition closed forms on (mu, var); var already clipped/un-normalised.
// kind: gpk_acq_kind.  Mirrors ei.py:70-78, log_ei.py:72-120, pi.py:61-63, lcb.py:65.
__device__ __forceinline__ double gpk_acq_value(int kind, double m, double v, double eta, double par) {
    double s = sqrt(v);
    if (kind == GPK_ACQ_EI) {
        double z = (eta - m - par) / s;
        if (z < -30.0) {
            // deep lower tail: z Phi(z) + phi(z) = phi(z) (1 - |z| R(|z|)), R = Mills ratio
            // = sqrt(pi/2) erfcx(|z|/sqrt 2).  Same value as the direct form (both lose ~z^2 ulp to
            // cancellation) but the sign is decided in the normal range, so EI never turns negative
            // when phi and Phi go subnormal (|z| > 37.6), where the reference's scipy result is >= 0.
            double az = -z;
            double bracket = 1.0 - az * 1.25331413731550025121 * erfcx(az * 0.70710678118654752440);
            return s * (gpk_norm_pdf(z) * bracket);
        }
        return s * (z * gpk_ndtr(z) + gpk_norm_pdf(z));
    } else if (kind == GPK_ACQ_PI) {
        return gpk_ndtr((eta - m - par) / s);
    } else if (kind == GPK_ACQ_LCB) {
        return -(m - par * s);
    } else if (kind == GPK_ACQ_LOG_EI) {
        double f_min = eta - par;
        double z = (f_min - m) / s;
        const double ninf = -INFINITY;
        if (fabs(f_min - m) == 0.0) {                       // log_ei.py:85-89
            return (s > 0.0) ? log(s) + gpk_norm_logpdf(z) : ninf;
        } else if (s == 0.0) {                              // log_ei.py:92-96
            return (m < f_min) ? log(f_min - m) : ninf;
        } else {
            double b = log(s) + gpk_norm_logpdf(z);         // log_ei.py:99
            if (f_min > m) {                                // log_ei.py:101-107
                double a = log(f_min - m) + gpk_log_ndtr(z);
                return fmax(a, b) + log(1.0 + exp(-fabs(b - a)));
            } else {                                        // log_ei.py:114-120
                double a = log(m - f_min) + gpk_log_ndtr(z);
                if (a >= b) return ninf;
                return b + log(1.0 - exp(a - b));
            }
        }
    }
    return 0.0;
}

// numpy.argmax ordering: NaN beats everything, then larger value, then lower index.
__device__ __forceinline__ bool gpk_better(double va, long long ia, double vb, long long ib) {
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    bool na = isnan(va), nb = isnan(vb);
    if (na || nb) {
        if (na && nb) return ia < ib;
        return na;
    }
    if (va > vb) return true;
    if (va < vb) return false;
    return ia < ib;
}

struct BestPair { double val; long long idx; };
