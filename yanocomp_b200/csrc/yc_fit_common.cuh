// Device helpers shared by the CTA-per-position fit kernels (yc_fit.cu) and the
// fused warp-per-position fit kernel (yc_fit_warp.cu).
#pragma once
#include "yc_common.cuh"

namespace yc {

#define YC_LOG_2PI 1.8378770664093453

// NumPy's add.reduce over a contiguous axis of length W (pairwise_sum with
// its 8-accumulator unrolling once W >= 8), on separately rounded terms.
template <int W>
__device__ __forceinline__ double np_sum(const double (&a)[W]) {
    if (W < 8) {
        double s = a[0];
#pragma unroll
        for (int j = 1; j < W; ++j) s = add_rn(s, a[j]);
        return s;
    } else {
        double s = add_rn(add_rn(add_rn(a[0], a[1]), add_rn(a[2], a[3])),
                          add_rn(add_rn(a[4], a[5]), add_rn(a[6], a[7])));
#pragma unroll
        for (int j = 8; j < W; ++j) s = add_rn(s, a[j]);
        return s;
    }
}
template <int W>
__device__ __forceinline__ double sqdist(const double (&x)[W], const double *c) {
    double t[W];
#pragma unroll
    for (int j = 0; j < W; ++j) { const double d = sub_rn(x[j], c[j]); t[j] = mul_rn(d, d); }
    return np_sum<W>(t);
}
// Gaussian of one label from its exact sufficient statistics
// (oracle MVNormal.from_samples: blank N(0,I), one fit at inertia 0).
template <int W>
__device__ __noinline__ bool ml_gaussian(const double *S, double *mu, double *cov) {
    // S = [S0, S1[W], S2 upper triangle]
    const double s0 = S[0];
    if (s0 < 1e-7) {
        #pragma unroll 1
        for (int a = 0; a < W; ++a) { mu[a] = 0.0; for (int c2 = 0; c2 < W; ++c2) cov[a * W + c2] = (a == c2) ? 1.0 : 0.0; }
        return true;
    }
    double u[W];
    #pragma unroll 1
    for (int a = 0; a < W; ++a) u[a] = div_rn(S[1 + a], s0);
    #pragma unroll 1
    for (int a = 0; a < W; ++a)
        #pragma unroll 1
        for (int c2 = 0; c2 < W; ++c2) {
            const int lo = a < c2 ? a : c2, hi = a < c2 ? c2 : a;
            const double s2 = S[1 + W + lo * W - lo * (lo - 1) / 2 + (hi - lo)];
            double t = sub_rn(s2, mul_rn(S[1 + a], u[c2]));
            t = sub_rn(t, mul_rn(u[a], S[1 + c2]));
            t = add_rn(t, mul_rn(s0, mul_rn(u[a], u[c2])));
            cov[a * W + c2] = div_rn(t, s0);
        }
    // inertia 0 blend: old*0 + new*1
    #pragma unroll 1
    for (int a = 0; a < W; ++a) mu[a] = u[a];
    // A cluster of 1..W rows has an EXACTLY singular ML covariance: whether its Cholesky pivot comes out as +1e-13
    // or -1e-13 is rounding noise (in LAPACK too), and EM started from such a matrix is chaotic.  The exact-arithmetic
    // answer is used instead of the noise: the fit fails (assumption A4, oracle switch A4_rank_deficient_init).
    if (s0 <= (double)W) return false;
    double linv[W * W], logdet;
    return cholesky_inverse<W>(cov, linv, logdet);
}

// parameter block in shared memory
template <int W>
struct EmPar {
    double mu[2][W];
    double cov[2][W * W];
    double linv[2][W * W];
    double cst[2];            // -0.5 (W log 2pi + log|cov|)
    double lo[W], hi[W];
    double ulogp;
    double logw[YC_MAX_SAMPLES][3];
    double ll;
    int ok;
};


}  // namespace yc
