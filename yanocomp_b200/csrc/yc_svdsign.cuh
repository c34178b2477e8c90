// First right singular vector of the standardised window, WITH the sign that
// numpy.linalg.svd (LAPACK dgesdd, JOBZ='S', path "M >> N") gives it.
//
// Why the sign matters: the KS statistic D of the reference is
// max |c1/n1 - c2/n2| evaluated in floating point at every pooled point of
// the projection (stats.py:34-43).  Mirroring the projection changes which
// (c1, c2) pairs are visited, and although the exact maximum is the same its
// ROUNDED value can differ by an ulp or two -- enough to flip `ks >= min_ks`
// for D sitting on the threshold (e.g. 20/100 vs 0.2).  So the axis
// orientation of LAPACK is part of the reference result and is reproduced
// here.
//
// dgesdd for an (M x d) matrix with M >= 11d/6 runs
//    dgeqr2 (Householder QR)  ->  dgebd2 on R  ->  dbdsdc -> dlasdq -> dbdsqr
//    (implicit-shift QR on the bidiagonal, starting from VT = I)  ->  dormbr.
// Only R = Q^T Z enters the right singular vectors, and R is fixed by
// G = Z^T Z up to the sign of each row; the signs come from the Householder
// pivots, which depend on the first d rows of Z only.  So one thread can
// replay the whole computation from G (d x d) and T = Z[0:d, :].
// The routines below follow LAPACK 3.11 (dlarfg, dgebd2, dbdsqr, dlartg,
// dlas2, dlasv2, dlasr) statement by statement; the Python prototype of the
// same code agrees with numpy.linalg.svd on the leading vector (sign
// included, |err| < 3e-13) on every one of 2e4 random windows with
// d in {3,5,7,9}.
#pragma once
#include <math.h>

namespace yc {
namespace lapack {

#define YC_LAPACK_EPS  1.1102230246251565e-16   /* dlamch('E') = 2^-53 */
#define YC_LAPACK_UNFL 2.2250738585072014e-308  /* dlamch('S') */

__host__ __device__ __forceinline__ double fsign(double a, double b) { return copysign(a, b); }

__host__ __device__ __noinline__ inline void dlartg(double f, double g, double &c, double &s, double &r) {
    if (g == 0.0) { c = 1.0; s = 0.0; r = f; }
    else if (f == 0.0) { c = 0.0; s = fsign(1.0, g); r = fabs(g); }
    else {
        const double d = sqrt(f * f + g * g);
        c = fabs(f) / d;
        r = fsign(d, f);
        s = g / r;
    }
}

__host__ __device__ __noinline__ inline void dlas2(double f, double g, double h, double &ssmin, double &ssmax) {
    const double fa = fabs(f), ga = fabs(g), ha = fabs(h);
    const double fhmn = fmin(fa, ha), fhmx = fmax(fa, ha);
    if (fhmn == 0.0) {
        ssmin = 0.0;
        if (fhmx == 0.0) ssmax = ga;
        else { const double q = fmin(fhmx, ga) / fmax(fhmx, ga); ssmax = fmax(fhmx, ga) * sqrt(1.0 + q * q); }
    } else if (ga < fhmx) {
        const double as = 1.0 + fhmn / fhmx, at = (fhmx - fhmn) / fhmx;
        const double au = (ga / fhmx) * (ga / fhmx);
        const double c = 2.0 / (sqrt(as * as + au) + sqrt(at * at + au));
        ssmin = fhmn * c; ssmax = fhmx / c;
    } else {
        const double au = fhmx / ga;
        if (au == 0.0) { ssmin = (fhmn * fhmx) / ga; ssmax = ga; }
        else {
            const double as = 1.0 + fhmn / fhmx, at = (fhmx - fhmn) / fhmx;
            const double c = 1.0 / (sqrt(1.0 + (as * au) * (as * au)) + sqrt(1.0 + (at * au) * (at * au)));
            ssmin = (fhmn * c) * au; ssmin = ssmin + ssmin;
            ssmax = ga / (c + c);
        }
    }
}

__host__ __device__ __noinline__ inline void dlasv2(double f, double g, double h, double &ssmin, double &ssmax,
                              double &snr, double &csr, double &snl, double &csl) {
    double ft = f, fa = fabs(f), ht = h, ha = fabs(h);
    int pmax = 1;
    const bool swap = ha > fa;
    if (swap) { pmax = 3; double t = ft; ft = ht; ht = t; t = fa; fa = ha; ha = t; }
    const double gt = g, ga = fabs(g);
    double clt, crt, slt, srt;
    if (ga == 0.0) { ssmin = ha; ssmax = fa; clt = 1.0; crt = 1.0; slt = 0.0; srt = 0.0; }
    else {
        bool gasmal = true;
        if (ga > fa) {
            pmax = 2;
            if ((fa / ga) < YC_LAPACK_EPS) {
                gasmal = false;
                ssmax = ga;
                ssmin = (ha > 1.0) ? fa / (ga / ha) : (fa / ga) * ha;
                clt = 1.0; slt = ht / gt; srt = 1.0; crt = ft / gt;
            }
        }
        if (gasmal) {
            const double d = fa - ha;
            double l = (d == fa) ? 1.0 : d / fa;
            const double m = gt / ft;
            double t = 2.0 - l;
            const double mm = m * m, tt = t * t;
            const double s = sqrt(tt + mm);
            const double r = (l == 0.0) ? fabs(m) : sqrt(l * l + mm);
            const double a = 0.5 * (s + r);
            ssmin = ha / a; ssmax = fa * a;
            if (mm == 0.0) {
                if (l == 0.0) t = fsign(2.0, ft) * fsign(1.0, gt);
                else t = gt / fsign(d, ft) + m / t;
            } else {
                t = (m / (s + t) + m / (r + l)) * (1.0 + a);
            }
            l = sqrt(t * t + 4.0);
            crt = 2.0 / l; srt = t / l;
            clt = (crt + srt * m) / a;
            slt = (ht / ft) * srt / a;
        }
    }
    if (swap) { csl = srt; snl = crt; csr = slt; snr = clt; }
    else { csl = clt; snl = slt; csr = crt; snr = srt; }
    double tsign;
    if (pmax == 1) tsign = fsign(1.0, csr) * fsign(1.0, csl) * fsign(1.0, f);
    else if (pmax == 2) tsign = fsign(1.0, snr) * fsign(1.0, csl) * fsign(1.0, g);
    else tsign = fsign(1.0, snr) * fsign(1.0, snl) * fsign(1.0, h);
    ssmax = fsign(ssmax, tsign);
    ssmin = fsign(ssmin, tsign * fsign(1.0, f) * fsign(1.0, h));
}

// dlasr('L','V',dir): rotation j acts on rows (ll+j, ll+j+1) of VT (D columns)
template <int D>
__host__ __device__ __noinline__ inline void dlasr_lv(double *VT, int ll, int nrot, const double *cs, const double *sn, bool forward) {
    _Pragma("unroll 1")
    for (int q = 0; q < nrot; ++q) {
        const int j = forward ? q : nrot - 1 - q;
        const double c = cs[j], s = sn[j];
        if (c != 1.0 || s != 0.0) {
            _Pragma("unroll 1")
            for (int k = 0; k < D; ++k) {
                const double a = VT[(ll + j + 1) * D + k], b = VT[(ll + j) * D + k];
                VT[(ll + j + 1) * D + k] = c * a - s * b;
                VT[(ll + j) * D + k] = s * a + c * b;
            }
        }
    }
}

// dbdsqr('U', n, ncvt=n, 0, 0): d, e (upper bidiagonal) -> singular values in
// d (decreasing), VT (row-major D x D, in/out) accumulates the right rotations.
// Indices in the comments are LAPACK's 1-based ones.  Returns false if the
// iteration limit is hit (never observed).
template <int D>
__host__ __device__ __noinline__ inline bool dbdsqr_vt(double *d, double *e, double *VT) {
    const int n = D;
    if (n > 1) {
        const double eps = YC_LAPACK_EPS;
        const double tolmul = fmax(10.0, fmin(100.0, pow(eps, -0.125)));
        const double tol = tolmul * eps;
        double smax = 0.0;
        _Pragma("unroll 1")
        for (int i = 0; i < n; ++i) smax = fmax(smax, fabs(d[i]));
        _Pragma("unroll 1")
        for (int i = 0; i < n - 1; ++i) smax = fmax(smax, fabs(e[i]));
        double sminoa = fabs(d[0]);
        if (sminoa != 0.0) {
            double mu = sminoa;
            _Pragma("unroll 1")
            for (int i = 1; i < n; ++i) {
                mu = fabs(d[i]) * (mu / (mu + fabs(e[i - 1])));
                sminoa = fmin(sminoa, mu);
                if (sminoa == 0.0) break;
            }
        }
        sminoa = sminoa / sqrt((double)n);
        const double thresh = fmax(tol * sminoa, 6.0 * ((double)n * ((double)n * YC_LAPACK_UNFL)));
        const int maxitdivn = 6 * n;
        int iterdivn = 0, it = -1, oldll = -1, oldm = -1, idir = 0;
        int m = n;
        double w1[D], w2[D];
        while (true) {
            if (m <= 1) break;
            if (it >= n) { it -= n; if (++iterdivn >= maxitdivn) return false; }
            smax = fabs(d[m - 1]);
            bool found = false;
            int ll = 0;
            _Pragma("unroll 1")
            for (int lll = 1; lll < m; ++lll) {
                ll = m - lll;
                const double abss = fabs(d[ll - 1]), abse = fabs(e[ll - 1]);
                if (abse <= thresh) { found = true; break; }
                smax = fmax(smax, fmax(abss, abse));
            }
            if (found) {
                e[ll - 1] = 0.0;
                if (ll == m - 1) { m -= 1; continue; }
            } else ll = 0;
            ll += 1;
            if (ll == m - 1) {          // 2 x 2 block
                double sigmn, sigmx, sinr, cosr, sinl, cosl;
                dlasv2(d[m - 2], e[m - 2], d[m - 1], sigmn, sigmx, sinr, cosr, sinl, cosl);
                d[m - 2] = sigmx; e[m - 2] = 0.0; d[m - 1] = sigmn;
                _Pragma("unroll 1")
                for (int k = 0; k < D; ++k) {        // drot(VT(m-1,:), VT(m,:), cosr, sinr)
                    const double x = VT[(m - 2) * D + k], y = VT[(m - 1) * D + k];
                    VT[(m - 2) * D + k] = cosr * x + sinr * y;
                    VT[(m - 1) * D + k] = cosr * y - sinr * x;
                }
                m -= 2;
                continue;
            }
            if (ll > oldm || m < oldll) idir = (fabs(d[ll - 1]) >= fabs(d[m - 1])) ? 1 : 2;
            double smin;
            bool conv = false;
            if (idir == 1) {
                if (fabs(e[m - 2]) <= fabs(tol) * fabs(d[m - 1])) { e[m - 2] = 0.0; continue; }
                double mu = fabs(d[ll - 1]);
                smin = mu;
                _Pragma("unroll 1")
                for (int lll = ll; lll < m; ++lll) {
                    if (fabs(e[lll - 1]) <= tol * mu) { e[lll - 1] = 0.0; conv = true; break; }
                    mu = fabs(d[lll]) * (mu / (mu + fabs(e[lll - 1])));
                    smin = fmin(smin, mu);
                }
            } else {
                if (fabs(e[ll - 1]) <= fabs(tol) * fabs(d[ll - 1])) { e[ll - 1] = 0.0; continue; }
                double mu = fabs(d[m - 1]);
                smin = mu;
                _Pragma("unroll 1")
                for (int lll = m - 1; lll >= ll; --lll) {
                    if (fabs(e[lll - 1]) <= tol * mu) { e[lll - 1] = 0.0; conv = true; break; }
                    mu = fabs(d[lll - 1]) * (mu / (mu + fabs(e[lll - 1])));
                    smin = fmin(smin, mu);
                }
            }
            if (conv) continue;
            oldll = ll; oldm = m;
            double shift, rr;
            if ((double)n * tol * (smin / smax) <= fmax(eps, 0.01 * tol)) shift = 0.0;
            else {
                double sll;
                if (idir == 1) { sll = fabs(d[ll - 1]); dlas2(d[m - 2], e[m - 2], d[m - 1], shift, rr); }
                else { sll = fabs(d[m - 1]); dlas2(d[ll - 1], e[ll - 1], d[ll], shift, rr); }
                if (sll > 0.0 && (shift / sll) * (shift / sll) < eps) shift = 0.0;
            }
            it += m - ll;
            const int nrot = m - ll;
            if (shift == 0.0) {
                if (idir == 1) {
                    double cs = 1.0, sn = 0.0, oldcs = 1.0, oldsn = 0.0, r;
                    _Pragma("unroll 1")
                    for (int i = ll; i < m; ++i) {
                        dlartg(d[i - 1] * cs, e[i - 1], cs, sn, r);
                        if (i > ll) e[i - 2] = oldsn * r;
                        dlartg(oldcs * r, d[i] * sn, oldcs, oldsn, d[i - 1]);
                        w1[i - ll] = cs; w2[i - ll] = sn;
                    }
                    const double h = d[m - 1] * cs;
                    d[m - 1] = h * oldcs; e[m - 2] = h * oldsn;
                    dlasr_lv<D>(VT, ll - 1, nrot, w1, w2, true);
                    if (fabs(e[m - 2]) <= thresh) e[m - 2] = 0.0;
                } else {
                    double cs = 1.0, sn = 0.0, oldcs = 1.0, oldsn = 0.0, r;
                    _Pragma("unroll 1")
                    for (int i = m; i > ll; --i) {
                        dlartg(d[i - 1] * cs, e[i - 2], cs, sn, r);
                        if (i < m) e[i - 1] = oldsn * r;
                        dlartg(oldcs * r, d[i - 2] * sn, oldcs, oldsn, d[i - 1]);
                        w1[i - ll - 1] = oldcs; w2[i - ll - 1] = -oldsn;
                    }
                    const double h = d[ll - 1] * cs;
                    d[ll - 1] = h * oldcs; e[ll - 1] = h * oldsn;
                    dlasr_lv<D>(VT, ll - 1, nrot, w1, w2, false);
                    if (fabs(e[ll - 1]) <= thresh) e[ll - 1] = 0.0;
                }
            } else {
                if (idir == 1) {
                    double f = (fabs(d[ll - 1]) - shift) * (fsign(1.0, d[ll - 1]) + shift / d[ll - 1]);
                    double g = e[ll - 1];
                    _Pragma("unroll 1")
                    for (int i = ll; i < m; ++i) {
                        double cosr, sinr, cosl, sinl, r;
                        dlartg(f, g, cosr, sinr, r);
                        if (i > ll) e[i - 2] = r;
                        f = cosr * d[i - 1] + sinr * e[i - 1];
                        e[i - 1] = cosr * e[i - 1] - sinr * d[i - 1];
                        g = sinr * d[i];
                        d[i] = cosr * d[i];
                        dlartg(f, g, cosl, sinl, r);
                        d[i - 1] = r;
                        f = cosl * e[i - 1] + sinl * d[i];
                        d[i] = cosl * d[i] - sinl * e[i - 1];
                        if (i < m - 1) { g = sinl * e[i]; e[i] = cosl * e[i]; }
                        w1[i - ll] = cosr; w2[i - ll] = sinr;
                    }
                    e[m - 2] = f;
                    dlasr_lv<D>(VT, ll - 1, nrot, w1, w2, true);
                    if (fabs(e[m - 2]) <= thresh) e[m - 2] = 0.0;
                } else {
                    double f = (fabs(d[m - 1]) - shift) * (fsign(1.0, d[m - 1]) + shift / d[m - 1]);
                    double g = e[m - 2];
                    _Pragma("unroll 1")
                    for (int i = m; i > ll; --i) {
                        double cosr, sinr, cosl, sinl, r;
                        dlartg(f, g, cosr, sinr, r);
                        if (i < m) e[i - 1] = r;
                        f = cosr * d[i - 1] + sinr * e[i - 2];
                        e[i - 2] = cosr * e[i - 2] - sinr * d[i - 1];
                        g = sinr * d[i - 2];
                        d[i - 2] = cosr * d[i - 2];
                        dlartg(f, g, cosl, sinl, r);
                        d[i - 1] = r;
                        f = cosl * e[i - 2] + sinl * d[i - 2];
                        d[i - 2] = cosl * d[i - 2] - sinl * e[i - 2];
                        if (i > ll + 1) { g = sinl * e[i - 3]; e[i - 3] = cosl * e[i - 3]; }
                        w1[i - ll - 1] = cosl; w2[i - ll - 1] = -sinl;
                    }
                    e[ll - 1] = f;
                    if (fabs(e[ll - 1]) <= thresh) e[ll - 1] = 0.0;
                    dlasr_lv<D>(VT, ll - 1, nrot, w1, w2, false);
                }
            }
        }
    }
    _Pragma("unroll 1")
    for (int i = 0; i < n; ++i)
        if (d[i] < 0.0) { d[i] = -d[i]; for (int k = 0; k < D; ++k) VT[i * D + k] = -VT[i * D + k]; }
    _Pragma("unroll 1")
    for (int i = 1; i < n; ++i) {       // sort into decreasing order
        int isub = 0;
        double smin = d[0];
        _Pragma("unroll 1")
        for (int j = 1; j < n + 1 - i; ++j) if (d[j] <= smin) { isub = j; smin = d[j]; }
        const int last = n - i;
        if (isub != last) {
            d[isub] = d[last]; d[last] = smin;
            _Pragma("unroll 1")
            for (int k = 0; k < D; ++k) { const double t = VT[isub * D + k]; VT[isub * D + k] = VT[last * D + k]; VT[last * D + k] = t; }
        }
    }
    return true;
}

// G = Z^T Z (row-major, symmetric), T = first D rows of Z (row-major).
// v receives the first right singular vector numpy.linalg.svd(Z) returns.
template <int D>
__host__ __device__ __noinline__ void first_right_singular_vector(const double *G, const double *Tin, double *v) {
    if (D == 1) { v[0] = 1.0; return; }          // dbdsdc: N = 1 -> VT = 1
    double S[D * D], T[D * D], A[D * D];
    _Pragma("unroll 1")
    for (int i = 0; i < D * D; ++i) { S[i] = G[i]; T[i] = Tin[i]; A[i] = 0.0; }
    // ---- R of dgeqr2 from the Gram matrix and the Householder pivots
    _Pragma("unroll 1")
    for (int i = 0; i < D; ++i) {
        const double alpha = T[i * D + i];
        const double xn2 = S[i * D + i] - alpha * alpha;
        if (!(xn2 > 0.0)) {
            _Pragma("unroll 1")
            for (int j = i; j < D; ++j) A[i * D + j] = T[i * D + j];
        } else {
            const double beta = -fsign(sqrt(S[i * D + i]), alpha);
            const double tau = (beta - alpha) / beta;
            const double inv = 1.0 / (alpha - beta);
            double w[D];
            A[i * D + i] = beta;
            _Pragma("unroll 1")
            for (int j = i + 1; j < D; ++j) {
                w[j] = T[i * D + j] + (S[i * D + j] - alpha * T[i * D + j]) * inv;
                A[i * D + j] = T[i * D + j] - tau * w[j];
            }
            _Pragma("unroll 1")
            for (int r = i + 1; r < D; ++r) {
                const double vr = T[r * D + i] * inv;
                _Pragma("unroll 1")
                for (int j = i + 1; j < D; ++j) T[r * D + j] -= tau * vr * w[j];
            }
        }
        _Pragma("unroll 1")
        for (int j = i + 1; j < D; ++j)
            _Pragma("unroll 1")
            for (int k = i + 1; k < D; ++k) S[j * D + k] -= A[i * D + j] * A[i * D + k];
    }
    // ---- dgebd2 on R (upper triangular, in A)
    double d[D], e[D], taup[D], vp[D * D];
    _Pragma("unroll 1")
    for (int i = 0; i < D; ++i) {
        // left reflector H(i) from A(i:, i)
        {
            const double alpha = A[i * D + i];
            double xn2 = 0.0;
            _Pragma("unroll 1")
            for (int r = i + 1; r < D; ++r) xn2 += A[r * D + i] * A[r * D + i];
            if (xn2 == 0.0) d[i] = alpha;
            else {
                const double xnorm = sqrt(xn2);
                const double beta = -fsign(hypot(alpha, xnorm), alpha);
                const double tau = (beta - alpha) / beta;
                const double inv = 1.0 / (alpha - beta);
                double vv[D];
                vv[i] = 1.0;
                _Pragma("unroll 1")
                for (int r = i + 1; r < D; ++r) vv[r] = A[r * D + i] * inv;
                _Pragma("unroll 1")
                for (int j = i + 1; j < D; ++j) {
                    double wj = 0.0;
                    _Pragma("unroll 1")
                    for (int r = i; r < D; ++r) wj += A[r * D + j] * vv[r];
                    _Pragma("unroll 1")
                    for (int r = i; r < D; ++r) A[r * D + j] -= tau * vv[r] * wj;
                }
                d[i] = beta;
            }
            A[i * D + i] = d[i];
        }
        if (i < D - 1) {
            // right reflector G(i) from A(i, i+1:)
            const double alpha = A[i * D + i + 1];
            double xn2 = 0.0;
            _Pragma("unroll 1")
            for (int c = i + 2; c < D; ++c) xn2 += A[i * D + c] * A[i * D + c];
            vp[i * D + i + 1] = 1.0;
            if (xn2 == 0.0) { e[i] = alpha; taup[i] = 0.0; }
            else {
                const double xnorm = sqrt(xn2);
                const double beta = -fsign(hypot(alpha, xnorm), alpha);
                const double tau = (beta - alpha) / beta;
                const double inv = 1.0 / (alpha - beta);
                _Pragma("unroll 1")
                for (int c = i + 2; c < D; ++c) vp[i * D + c] = A[i * D + c] * inv;
                _Pragma("unroll 1")
                for (int r = i + 1; r < D; ++r) {
                    double wr = 0.0;
                    _Pragma("unroll 1")
                    for (int c = i + 1; c < D; ++c) wr += A[r * D + c] * vp[i * D + c];
                    _Pragma("unroll 1")
                    for (int c = i + 1; c < D; ++c) A[r * D + c] -= tau * wr * vp[i * D + c];
                }
                e[i] = beta; taup[i] = tau;
            }
            A[i * D + i + 1] = e[i];
        }
    }
    double VT[D * D];
    _Pragma("unroll 1")
    for (int i = 0; i < D * D; ++i) VT[i] = 0.0;
    _Pragma("unroll 1")
    for (int i = 0; i < D; ++i) VT[i * D + i] = 1.0;
    dbdsqr_vt<D>(d, e, VT);
    // ---- dormbr('P','R','T'): VT <- VT G(n-1) ... G(1)  (only row 0 is needed)
    double row[D];
    _Pragma("unroll 1")
    for (int k = 0; k < D; ++k) row[k] = VT[k];
    _Pragma("unroll 1")
    for (int i = D - 2; i >= 0; --i) {
        if (taup[i] != 0.0) {
            double wv = 0.0;
            _Pragma("unroll 1")
            for (int c = i + 1; c < D; ++c) wv += row[c] * vp[i * D + c];
            _Pragma("unroll 1")
            for (int c = i + 1; c < D; ++c) row[c] -= taup[i] * wv * vp[i * D + c];
        }
    }
    _Pragma("unroll 1")
    for (int k = 0; k < D; ++k) v[k] = row[k];
}

}  // namespace lapack
}  // namespace yc
