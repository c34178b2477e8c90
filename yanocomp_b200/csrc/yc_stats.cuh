// Scalar statistics shared by host and device: Kolmogorov survival function,
// chi-square (1 dof) survival function, the G-tests of stats.py:302-338 and
// the fractional statistics of stats.py:341-355.  __host__ __device__ so the
// CPU test-suite can pin them against SciPy without a GPU (yc_*_host exports).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/yanocomp_b200.h"

#ifdef __CUDACC__
#define YC_HD __host__ __device__ inline
#define YC_HD_COLD __host__ __device__ __noinline__ inline      /* serial, cold: keep one copy */
#else
#define YC_HD inline
#define YC_HD_COLD inline
#endif

namespace yc {

// scipy.stats.kstwobign.sf == scipy.special.kolmogorov  (stats.py:48).
// Two-series evaluation (Kolmogorov 1933; cutover as in SciPy's cephes port):
//   x <= 0.82:  1 - sqrt(2 pi)/x * sum_{k>=1} exp(-(2k-1)^2 pi^2 / (8 x^2))
//   x >  0.82:  2 * sum_{k>=1} (-1)^(k-1) exp(-2 k^2 x^2)
YC_HD_COLD double kolmogorov_sf(double x) {
    if (isnan(x)) return x;
    if (x <= 0.040611972203751713) return 1.0;      // cdf underflows to 0
    const double pi = 3.14159265358979323846;
    if (x <= 0.82) {
        const double w = sqrt(2.0 * pi) / x;
        const double logu8 = -pi * pi / (x * x);    // log(u^8), u = exp(-pi^2/(8x^2))
        const double u = exp(logu8 / 8.0);
        if (u == 0.0) return 1.0;
        const double u8 = exp(logu8);
        const double u16 = u8 * u8;
        double P = 1.0 + u16 * u8;                  // 1 + u^24
        P = 1.0 + u16 * P;                          // 1 + u^16 (1 + u^24)
        P = 1.0 + u8 * P;                           // 1 + u^8 (1 + u^16 (1 + u^24))
        P = w * u * P;
        double sf = 1.0 - P;
        return sf < 0.0 ? 0.0 : (sf > 1.0 ? 1.0 : sf);
    }
    const double v = exp(-2.0 * x * x);
    const double v3 = v * v * v;
    double P = 1.0 - v3 * v3 * v * 1.0;             // 1 - v^7
    P = 1.0 - v3 * v * v * P;                       // 1 - v^5 (...)
    P = 1.0 - v3 * P;                               // 1 - v^3 (...)
    P = 2.0 * v * P;
    return P < 0.0 ? 0.0 : (P > 1.0 ? 1.0 : P);
}

// The gate of position_stats (stats.py:465): ks_stat >= min_ks and ks_p < fdr_threshold.  D is bit-exact, but the
// device p-value is only within a few ulp of scipy.special.kolmogorov (libm's exp / pow are not reproducible on
// the device), so a p-value within 2^-44 (relative) of the threshold passes the gate as well; the host re-decides
// exactly those positions with SciPy itself (engine._host_ks_pvalues) and drops the ones SciPy rejects.  Every
// filter decision is therefore the one the reference's own call makes.
#define YC_KS_P_GUARD 5.6843418860808015e-14
YC_HD bool ks_p_near_threshold(double pv, double thr) { return fabs(pv - thr) <= thr * YC_KS_P_GUARD; }
YC_HD bool ks_gate_open(double d, double pv, double min_ks, double thr) {
    return d >= min_ks && (pv < thr || ks_p_near_threshold(pv, thr));
}

// chi2.sf(g, 1) = erfc(sqrt(g/2))
YC_HD double chi2_sf_1dof(double g) {
    if (isnan(g)) return g;
    if (g <= 0.0) return 1.0;
    return erfc(sqrt(0.5 * g));
}

YC_HD double xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }

// -norm.ppf((1 - 95/100) / 2): the statsmodels Table2x2 CI multiplier at the
// alpha the reference passes (stats.py:354, alpha = 1 - ci/100 in floating point)
#define YC_Z_CI95 1.959963984540054

// G statistic of the first `ncol` columns of an nrow x 3 table (row stride 3), as
// scipy.stats.chi2_contingency(lambda_='log-likelihood') computes it: zero expected cell or
// dof 0 -> 0.0 (two_cond_g_test, stats.py:302-313); dof 1 (2 x 2, only with --no-uniform) gets
// Yates' continuity correction of magnitude min(0.5, |E - O|).
YC_HD_COLD double g_stat_table(const double *obs, int nrow, int ncol) {
    double rs[YC_MAX_SAMPLES], cs[YC_NCOMP] = {0.0, 0.0, 0.0}, tot = 0.0;
    _Pragma("unroll 1")
    for (int i = 0; i < nrow; ++i) {
        rs[i] = 0.0;
        _Pragma("unroll 1")
        for (int j = 0; j < ncol; ++j) { rs[i] += obs[i * 3 + j]; cs[j] += obs[i * 3 + j]; }
        tot += rs[i];
    }
    if (!(tot > 0.0)) return 0.0;
    _Pragma("unroll 1")
    for (int i = 0; i < nrow; ++i) if (rs[i] == 0.0) return 0.0;
    _Pragma("unroll 1")
    for (int j = 0; j < ncol; ++j) if (cs[j] == 0.0) return 0.0;
    const int dof = (nrow - 1) * (ncol - 1);
    if (dof == 0) return 0.0;
    double g = 0.0;
    _Pragma("unroll 1")
    for (int i = 0; i < nrow; ++i) {
        _Pragma("unroll 1")
        for (int j = 0; j < ncol; ++j) {
            const double e = rs[i] * cs[j] / tot;
            double o = obs[i * 3 + j];
            if (dof == 1) {
                const double diff = e - o;
                o += fmin(0.5, fabs(diff)) * (diff > 0.0 ? 1.0 : (diff < 0.0 ? -1.0 : 0.0));
            }
            g += 2.0 * xlogy(o, o / e);
        }
    }
    return g;
}

// out[8] = g, hom_g, p, cntrl_frac, treat_frac, log_odds, lower_ci, upper_ci.
// preds: [nrow_c + nrow_t][3] rounded counts of the PRESENT samples, control
// rows first.  gmm_g_test (stats.py:316-338) + calculate_fractional_stats
// (stats.py:341-355).
YC_HD_COLD void gtest_fractions(const double *preds, int nrow_c, int nrow_t, double thr, double *out,
                                int ncomp = 3) {
    const double nan_ = nan("");
    double c[3] = {0, 0, 0}, t[3] = {0, 0, 0};
    _Pragma("unroll 1")
    for (int i = 0; i < nrow_c; ++i) for (int j = 0; j < 3; ++j) c[j] += preds[i * 3 + j];
    _Pragma("unroll 1")
    for (int i = 0; i < nrow_t; ++i) for (int j = 0; j < 3; ++j) t[j] += preds[(nrow_c + i) * 3 + j];
    // heterogeneity 2x2 with Yates' correction, magnitude min(0.5, |E - O|)
    double o[4] = {c[0], c[1], t[0], t[1]};
    const double r0 = o[0] + o[1], r1 = o[2] + o[3], c0 = o[0] + o[2], c1 = o[1] + o[3];
    const double tot = r0 + r1;
    double het, p;
    if (tot == 0.0) {                       // all outliers: FloatingPointError path
        het = nan_; p = 1.0;
    } else if (r0 == 0.0 || r1 == 0.0 || c0 == 0.0 || c1 == 0.0) {
        het = 0.0; p = 1.0;                 // zero expected frequency: ValueError path
    } else {
        const double e[4] = {r0 * c0 / tot, r0 * c1 / tot, r1 * c0 / tot, r1 * c1 / tot};
        het = 0.0;
        _Pragma("unroll 1")
        for (int k = 0; k < 4; ++k) {
            const double diff = e[k] - o[k];
            const double mag = fmin(0.5, fabs(diff));
            const double sgn = diff > 0.0 ? 1.0 : (diff < 0.0 ? -1.0 : 0.0);
            const double oc = o[k] + mag * sgn;
            het += 2.0 * xlogy(oc, oc / e[k]);
        }
        p = chi2_sf_1dof(het);
    }
    double hom;
    if (p < thr) {
        hom = g_stat_table(preds, nrow_c, ncomp) + g_stat_table(preds + nrow_c * 3, nrow_t, ncomp);
        if (hom >= het) p = 1.0;
    } else {
        hom = nan_;
    }
    out[0] = het; out[1] = hom; out[2] = p;
    out[3] = c[1] / (c[0] + c[1]);
    out[4] = t[1] / (t[0] + t[1]);
    // Table2x2(shift_zeros=True)
    double f[4] = {c[0], c[1], t[0], t[1]};
    if (f[0] == 0.0 || f[1] == 0.0 || f[2] == 0.0 || f[3] == 0.0)
        _Pragma("unroll 1")
        for (int k = 0; k < 4; ++k) f[k] += 0.5;
    const double lor = log(f[0]) - log(f[1]) - log(f[2]) + log(f[3]);
    const double se = sqrt(1.0 / f[0] + 1.0 / f[1] + 1.0 / f[2] + 1.0 / f[3]);
    out[5] = lor; out[6] = lor - YC_Z_CI95 * se; out[7] = lor + YC_Z_CI95 * se;
}

}  // namespace yc
