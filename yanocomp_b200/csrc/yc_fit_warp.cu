// Kernels 2a + 2b + 3 fused, small-window fast path: ONE WARP (one 32-thread
// CTA) per gated position does model initialisation, EM, the G-tests and the
// single-molecule posteriors without leaving the SM.
//
// Same arithmetic as gmm_init_kernel / gmm_em_kernel / gtest_smpred_kernel in
// yc_fit.cu (see there for the reference citations); reorganised after the
// round-1 profile: with ~200 reads per position a 128-thread CTA spends its
// time in block barriers around 2-3 busy threads (M-step, Cholesky).  A warp
// needs no barriers (`__syncwarp`), keeps the model in shared memory across
// the three stages (no global round trip, one launch instead of three), and
// the hardware scheduler load-balances the very uneven EM lengths (2..1001
// E-steps) because every position is its own CTA.
// Eligibility (dispatcher): w <= 5, <= 1024 reads per feature, no sub-sampling
// possible (reads per condition <= max_fit_depth).
#include "yc_fit_common.cuh"
#include "yc_stats.cuh"

namespace yc {

struct FwSmem {
    int xs_off, srow_off, lab_off, buf_off, par_off, mask_off, cnt_off, fstart_off, wtmp_off, red_off, total;
};

__host__ __device__ inline FwSmem fw_layout(int W, int R, int cap) {
    FwSmem s;
    int P = 32;
    while (P < cap) P <<= 1;
    int o = 0;
    s.buf_off = o;    o += (P > 2 * cap ? P : 2 * cap) * (int)sizeof(double);   // MAD sort buffer / EM responsibilities
    s.par_off = o;    o += (4 * W * W + 4 * W + 3 * YC_MAX_SAMPLES + 8) * (int)sizeof(double);
    s.wtmp_off = o;   o += 2 * YC_MAX_SAMPLES * (int)sizeof(double);
    s.red_off = o;    o += (2 * (1 + W + W * (W + 1) / 2) + 1) * 33 * (int)sizeof(double);
    s.mask_off = o;   o += ((cap + 31) / 32 + 1) * (int)sizeof(uint32_t);
    s.cnt_off = o;    o += (W + 1) * R * (int)sizeof(int);
    s.fstart_off = o; o += (YC_MAX_SAMPLES + 2) * (int)sizeof(int);
    s.xs_off = o;     o += W * cap * (int)sizeof(uint16_t);
    o = (o + 3) & ~3;
    s.srow_off = o;   o += cap;
    s.lab_off = o;    o += cap;
    s.total = (o + 15) & ~15;
    return s;
}

template <int W>
__device__ __forceinline__ void fw_row(const uint16_t *xs, int cap, int i, double (&x)[W]) {
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = h2d(xs[j * cap + i]);
}

__device__ __forceinline__ void warp_bitonic1(double *a, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = ((i & k) == 0);
                    const double x = a[i], y = a[l];
                    if ((x > y) == up) { a[i] = y; a[l] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// M-step of one shared Gaussian (reference stats.py:202-205 via pomegranate from_summaries,
// assumptions A3/A4).  Cold, serial code: one copy, rolled loops.
template <int W>
__device__ __noinline__ void em_mstep(EmPar<W> *par, int k, const double *red, double *carry, double step) {
    constexpr int NS = 1 + W + W * (W + 1) / 2;
    double S[NS];
#pragma unroll 1
    for (int q = 0; q < NS; ++q) S[q] = red[(k * NS + q) * 33 + 32] + carry[q];
    if (S[0] < 1e-7) {                                                   // skipped, summaries kept
#pragma unroll 1
        for (int q = 0; q < NS; ++q) carry[q] = S[q];
        return;
    }
#pragma unroll 1
    for (int q = 0; q < NS; ++q) carry[q] = 0.0;
    double u[W];
    const double inv0 = 1.0 / S[0], keep = 1.0 - step;
#pragma unroll 1
    for (int a = 0; a < W; ++a) u[a] = S[1 + a] * inv0;
#pragma unroll 1
    for (int a = 0; a < W; ++a) {
#pragma unroll 1
        for (int c2 = a; c2 < W; ++c2) {
            const double s2 = S[1 + W + a * W - a * (a - 1) / 2 + (c2 - a)];
            const double cv = (s2 - S[1 + a] * u[c2] - u[a] * S[1 + c2] + S[0] * (u[a] * u[c2])) * inv0;
            const double nv = par->cov[k][a * W + c2] * step + cv * keep;
            par->cov[k][a * W + c2] = nv; par->cov[k][c2 * W + a] = nv;
        }
    }
#pragma unroll 1
    for (int a = 0; a < W; ++a) par->mu[k][a] = par->mu[k][a] * step + u[a] * keep;
    double logdet;
    if (!cholesky_inverse<W>(par->cov[k], par->linv[k], logdet)) par->ok = 0;
    par->cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
}

// counts -> G-tests -> fractions of one fitted position (lane 0); returns the combined weights
template <int W>
__device__ __noinline__ void fit_finalize(EmPar<W> *par, const int *fstart, int R, int Rc, double thr,
                                          double *preds_out, double *cw_out, double *stats_out) {
    double cw[3] = {0.0, 0.0, 0.0}, tot = 0.0;
    double pr[YC_MAX_SAMPLES * 3];
    int rows_c = 0, rows_t = 0, q = 0;
#pragma unroll 1
    for (int s = 0; s < R; ++s) {
        const int sz = fstart[s + 1] - fstart[s];
        double *po = preds_out + s * 3;
        if (sz > 0) {
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const double wk = exp(par->logw[s][k]);
                cw[k] += wk * (double)sz;
                const double v = rint(wk * (double)sz);                  // np.round(weights * sizes), stats.py:296
                pr[q * 3 + k] = v; po[k] = v;
            }
            tot += (double)sz;
            ++q;
            if (s < Rc) ++rows_c; else ++rows_t;
        } else { po[0] = po[1] = po[2] = 0.0; }
    }
#pragma unroll 1
    for (int k = 0; k < 3; ++k) { cw[k] /= tot; cw_out[k] = cw[k]; }
    gtest_fractions(pr, rows_c, rows_t, thr, stats_out);
    const double t = cw[0] + cw[1] + cw[2];                               // A7
#pragma unroll 1
    for (int k = 0; k < 3; ++k) par->logw[0][k] = log(cw[k] / t);
}

#ifndef YC_FW_MINBLOCKS
#define YC_FW_MINBLOCKS 1
#endif

template <int W>
__global__ void __launch_bounds__(32, YC_FW_MINBLOCKS)
fit_warp_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
                const int32_t *__restrict__ fit_list, const yc_fit_out_t out,
                uint8_t *__restrict__ status, int8_t *__restrict__ labels_out, const int max_fit_rows,
                double *__restrict__ post, const long long post_capacity,
                unsigned long long *__restrict__ post_cursor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl, lane = threadIdx.x;
    const FwSmem L = fw_layout(W, R, cap);
    double *buf = reinterpret_cast<double *>(smem_raw + L.buf_off);
    EmPar<W> *par = reinterpret_cast<EmPar<W> *>(smem_raw + L.par_off);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + L.mask_off);
    int *cnt = reinterpret_cast<int *>(smem_raw + L.cnt_off);
    int *fstart = reinterpret_cast<int *>(smem_raw + L.fstart_off);
    uint16_t *xs = reinterpret_cast<uint16_t *>(smem_raw + L.xs_off);
    uint8_t *srow = reinterpret_cast<uint8_t *>(smem_raw + L.srow_off);
    int8_t *lab = reinterpret_cast<int8_t *>(smem_raw + L.lab_off);
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT, NP = 2 * NS + 1;
    const uint32_t lt = (1u << lane) - 1u;

    const int slot = blockIdx.x;
    const int64_t g = fit_list[slot];
    if (lane == 0) out.post_off[slot] = -1;

    // ------------------------------------------------ front end: complete rows -> shared memory
    Window<W> w;
    if (!window_setup<W>(b, p, g, w)) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }
    for (int i = lane; i < (W + 1) * R; i += 32) cnt[i] = 0;
    __syncwarp();
    {   // warp-local scan (same as yc_ks_warp.cu)
        const int nr = w.n_reads;
        for (int base = 0; base < nr; base += 32) {
            const int r = base + lane;
            const bool inb = r < nr;
            uint32_t m_all = YC_FULL;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const uint16_t v = inb ? __ldg(w.col[j] + r) : (uint16_t)0x7e00;
                m_all &= __ballot_sync(YC_FULL, inb && !h_isnan(v));
            }
            if (lane == 0) mask[base >> 5] = m_all;
            if (lane == 0) {
                for (int s = 0; s < R; ++s) {
                    const int a = max(w.samp_start[s], base), e = min(w.samp_start[s + 1], base + 32);
                    if (e > a) {
                        const uint32_t seg = (e - a == 32) ? YC_FULL : (((1u << (e - a)) - 1u) << (a - base));
                        cnt[W * R + s] += __popc(m_all & seg);
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        int acc = 0;
        for (int s = 0; s < R; ++s) { fstart[s] = acc; acc += cnt[W * R + s]; }
        fstart[R] = acc;
    }
    __syncwarp();
    const int n = fstart[R];
    if (n < 1 || n > cap) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }
    {   // ordered compaction
        int run = 0;
        const int nr = w.n_reads;
        for (int base = 0; base < nr; base += 32) {
            const uint32_t m = mask[base >> 5];
            if ((m >> lane) & 1u) {
                const int r = base + lane, rank = run + __popc(m & lt);
#pragma unroll
                for (int j = 0; j < W; ++j) xs[j * cap + rank] = __ldg(w.col[j] + r);
                int s = 0;
                while (s + 1 < R && r >= w.samp_start[s + 1]) ++s;
                srow[rank] = (uint8_t)s;
            }
            run += __popc(m);
        }
    }
    __syncwarp();
    const uint64_t pkey = w.pkey;

    // ------------------------------------------------ k-means++ x3 + Lloyd  (assumption A2)
    double best[2 * W], best_total = INFINITY;
    bool have_best = false;
    const int C = (n + 31) / 32;
    const int i0c = min(lane * C, n), i1c = min(i0c + C, n);
    for (int restart = 0; restart < 3; ++restart) {
        double cen[2 * W];
        {
            const double u0 = rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 0);
            const int i0 = min((int)(u0 * (double)n), n - 1);
            double x[W];
            fw_row<W>(xs, cap, i0, x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[j] = x[j];
        }
        {
            double loc = 0.0;
            for (int i = i0c; i < i1c; ++i) { double x[W]; fw_row<W>(xs, cap, i, x); loc = add_rn(loc, sqdist<W>(x, cen)); }
            double inc = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(YC_FULL, inc, o); if (lane >= o) inc = add_rn(inc, t); }
            const double total = __shfl_sync(YC_FULL, inc, 31);
            double run = sub_rn(inc, loc);
            const double target = mul_rn(rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 1), total);
            int below = 0;
            for (int i = i0c; i < i1c; ++i) { double x[W]; fw_row<W>(xs, cap, i, x); run = add_rn(run, sqdist<W>(x, cen)); below += (run <= target); }
            below = warp_sum_i(below);
            double x[W];
            fw_row<W>(xs, cap, min(below, n - 1), x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[W + j] = x[j];
        }
        int iteration = 0;
        double improvement = INFINITY, last = INFINITY, total = INFINITY;
        double acc[2 * W + 3];
        bool have_sums = false;
        while (improvement > 0.5 && iteration < 101) {
            if (have_sums) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (acc[2 * W + 1 + k] > 0.0)
#pragma unroll
                        for (int j = 0; j < W; ++j) cen[k * W + j] = div_rn(acc[k * W + j], acc[2 * W + 1 + k]);
            }
#pragma unroll
            for (int k = 0; k < 2 * W + 3; ++k) acc[k] = 0.0;
            for (int i = lane; i < n; i += 32) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
                const double d0 = sqdist<W>(x, cen), d1 = sqdist<W>(x, cen + W);
                const int k = d1 < d0;
                acc[2 * W] += k ? d1 : d0;
                acc[2 * W + 1 + k] += 1.0;
#pragma unroll
                for (int j = 0; j < W; ++j) { if (k) acc[W + j] += x[j]; else acc[j] += x[j]; }
            }
#pragma unroll
            for (int k = 0; k < 2 * W + 3; ++k) acc[k] = warp_sum(acc[k]);
            have_sums = true;
            total = acc[2 * W];
            if (iteration > 0) improvement = last - total;
            ++iteration;
            last = total;
        }
        if (total < best_total) {
            best_total = total; have_best = true;
#pragma unroll
            for (int j = 0; j < 2 * W; ++j) best[j] = cen[j];
        }
    }
    if (!have_best) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }

    // ------------------------------------------------ MAD outliers (stats.py:114-122, A1)
    int P = 32;
    while (P < n) P <<= 1;
    for (int i = lane; i < P; i += 32) {
        double dmin = INFINITY;
        if (i < n) {
            double x[W];
            fw_row<W>(xs, cap, i, x);
            const double d0 = sqdist<W>(x, best), d1 = sqdist<W>(x, best + W);
            lab[i] = (int8_t)(d1 < d0);
            dmin = d1 < d0 ? d1 : d0;
        }
        buf[i] = dmin;
    }
    __syncwarp();
    warp_bitonic1(buf, P, lane);
    const double med = (n & 1) ? buf[n >> 1] : mul_rn(add_rn(buf[(n >> 1) - 1], buf[n >> 1]), 0.5);
    // MAD = median of |d - med| WITHOUT a second sort: in the sorted distances the deviations are
    // V-shaped, i.e. two ascending runs (left of the median walked outwards, right of it), so the
    // order statistics of their union follow from merge ranks found by binary search.
    double *wsel = reinterpret_cast<double *>(smem_raw + L.wtmp_off);
    const int m_le = upper_bound(buf, n, med);            // elements <= med form the left run
    const int k_hi = n >> 1, k_lo = (n & 1) ? k_hi : k_hi - 1;
    for (int i = lane; i < n; i += 32) {
        const double dv = fabs(sub_rn(buf[i], med));
        int rank;
        if (i < m_le) {                                    // left run, index m_le-1-i; count right-run elements < dv
            int lo = 0, hi = n - m_le;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (fabs(sub_rn(buf[m_le + mid], med)) < dv) lo = mid + 1; else hi = mid; }
            rank = (m_le - 1 - i) + lo;
        } else {                                           // right run, index i-m_le; count left-run elements <= dv
            int lo = 0, hi = m_le;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (fabs(sub_rn(buf[m_le - 1 - mid], med)) <= dv) lo = mid + 1; else hi = mid; }
            rank = (i - m_le) + lo;
        }
        if (rank == k_lo) wsel[0] = dv;
        if (rank == k_hi) wsel[1] = dv;
    }
    __syncwarp();
    const double mad = (n & 1) ? wsel[1] : mul_rn(add_rn(wsel[0], wsel[1]), 0.5);
    const double thr = mul_rn(p.outlier_factor, mad);
    __syncwarp();
    int n_out = 0;
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(xs, cap, i, x);
        const double d0 = sqdist<W>(x, best), d1 = sqdist<W>(x, best + W);
        if ((d1 < d0 ? d1 : d0) > thr) { lab[i] = 2; ++n_out; }
    }
    n_out = warp_sum_i(n_out);
    __syncwarp();
    if (labels_out) for (int i = lane; i < n && i < max_fit_rows; i += 32)
        labels_out[(size_t)slot * max_fit_rows + i] = lab[i];

    // ------------------------------------------------ ML Gaussians, weights, box (stats.py:148-187)
    if (lane == 0) par->ok = 1;
    {
        double S0[NS], S1[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) { S0[q] = 0.0; S1[q] = 0.0; }
        for (int i = lane; i < n; i += 32) {
            const int l = lab[i];
            if (l < 2) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
                const double f0 = l == 0 ? 1.0 : 0.0, f1 = 1.0 - f0;
                S0[0] += f0; S1[0] += f1;
                int q = 1 + W;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    S0[1 + a] += f0 * x[a]; S1[1 + a] += f1 * x[a];
#pragma unroll
                    for (int c2 = a; c2 < W; ++c2) { const double pr = x[a] * x[c2]; S0[q] += f0 * pr; S1[q] += f1 * pr; ++q; }
                }
            }
        }
        double S[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const double t0 = warp_sum(S0[q]), t1 = warp_sum(S1[q]);
            S[q] = lane == 1 ? t1 : t0;
        }
        __syncwarp();
        if (lane < 2) {
            double mu[W], cov[W * W];
            const bool ok = ml_gaussian<W>(S, mu, cov);
            for (int a = 0; a < W; ++a) par->mu[lane][a] = mu[a];
            for (int a = 0; a < W * W; ++a) par->cov[lane][a] = cov[a];
            double logdet;
            if (!ok || !cholesky_inverse<W>(par->cov[lane], par->linv[lane], logdet)) par->ok = 0;
            else par->cst[lane] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
        }
    }
    {
        double lo[W], hi[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { lo[j] = INFINITY; hi[j] = -INFINITY; }
        for (int i = lane; i < n; i += 32) {
            if (n_out == 0 || lab[i] == 2) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
#pragma unroll
                for (int j = 0; j < W; ++j) { lo[j] = fmin(lo[j], x[j]); hi[j] = fmax(hi[j], x[j]); }
            }
        }
        bool box_ok = true;
        double ul = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            lo[j] = warp_min(lo[j]); hi[j] = warp_max(hi[j]);
            box_ok &= hi[j] > lo[j];
            if (lane == 0) { par->lo[j] = lo[j]; par->hi[j] = hi[j]; }
            ul += log(hi[j] - lo[j]);
        }
        __syncwarp();
        if (lane == 0) { par->ulogp = -ul; if (!box_ok) par->ok = 0; }
    }
    for (int s = 0; s < R; ++s) {
        const int a = fstart[s], e = fstart[s + 1];
        int k0 = 0, k1 = 0, k2 = 0;
        for (int i = a + lane; i < e; i += 32) { const int l = lab[i]; k0 += l == 0; k1 += l == 1; k2 += l == 2; }
        k0 = warp_sum_i(k0); k1 = warp_sum_i(k1); k2 = warp_sum_i(k2);
        if (lane == 0) {
            if (e > a) {
                const double den = (double)(e - a) + 1.5;
                par->logw[s][0] = log(((double)k0 + 0.5) / den);
                par->logw[s][1] = log(((double)k1 + 0.5) / den);
                par->logw[s][2] = log(((double)k2 + 0.5) / den);
            } else { par->logw[s][0] = par->logw[s][1] = par->logw[s][2] = 0.0; }
        }
    }
    __syncwarp();
    bool failed = par->ok == 0;

    // ------------------------------------------------ EM (stats.py:194-225)
    double carry[NS];                       // lanes 0/1: summaries kept when S0 < 1e-7 (A3)
#pragma unroll
    for (int q = 0; q < NS; ++q) carry[q] = 0.0;
    double *rb0 = buf, *rb1 = buf + cap;     // responsibilities of components 0 / 1 per fit row
    double *wtmp = reinterpret_cast<double *>(smem_raw + L.wtmp_off);
    double *red = reinterpret_cast<double *>(smem_raw + L.red_off);
    int iteration = 0;
    double improvement = INFINITY, last = 0.0;
    while (!failed && improvement > 0.1 && iteration < 1001) {
        double acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = 0.0;
        double mu0[W], mu1[W], li0[NT], li1[NT], blo[W], bhi[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { mu0[j] = par->mu[0][j]; mu1[j] = par->mu[1][j]; blo[j] = par->lo[j]; bhi[j] = par->hi[j]; }
        {
            int q = 0;
#pragma unroll
            for (int a = 0; a < W; ++a)
#pragma unroll
                for (int bb = 0; bb <= a; ++bb) { li0[q] = par->linv[0][a * W + bb]; li1[q] = par->linv[1][a * W + bb]; ++q; }
        }
        const double c0 = par->cst[0], c1 = par->cst[1], ul = par->ulogp;
        for (int i = lane; i < n; i += 32) {
            double x[W];
            fw_row<W>(xs, cap, i, x);
            const int s = srow[i];
            double q0 = 0.0, q1 = 0.0;
            bool inside = true;
            int q = 0;
#pragma unroll
            for (int a = 0; a < W; ++a) {
                double y0 = 0.0, y1 = 0.0;
#pragma unroll
                for (int bb = 0; bb <= a; ++bb) { y0 = fma(li0[q], x[bb] - mu0[bb], y0); y1 = fma(li1[q], x[bb] - mu1[bb], y1); ++q; }
                q0 = fma(y0, y0, q0); q1 = fma(y1, y1, q1);
                inside &= (x[a] >= blo[a]) && (x[a] <= bhi[a]);
            }
            const double l0 = c0 - 0.5 * q0 + par->logw[s][0];
            const double l1 = c1 - 0.5 * q1 + par->logw[s][1];
            const double l2 = inside ? ul + par->logw[s][2] : -INFINITY;
            const double m = fmax(fmax(l0, l1), l2);
            const double e0 = exp(l0 - m), e1 = exp(l1 - m), e2 = exp(l2 - m);
            const double se = e0 + e1 + e2, inv = 1.0 / se;
            const double r0 = e0 * inv, r1 = e1 * inv;
            rb0[i] = r0; rb1[i] = r1;
            acc[2 * NS] += m + log(se);
            acc[0] += r0; acc[NS] += r1;
            int qq = 1 + W;
#pragma unroll
            for (int a = 0; a < W; ++a) {
                const double xa0 = r0 * x[a], xa1 = r1 * x[a];
                acc[1 + a] += xa0; acc[NS + 1 + a] += xa1;
#pragma unroll
                for (int bb = a; bb < W; ++bb) { acc[qq] = fma(xa0, x[bb], acc[qq]); acc[NS + qq] = fma(xa1, x[bb], acc[NS + qq]); ++qq; }
            }
        }
        // 43 (w=3: 21) partial sums x 32 lanes: park them in shared memory (row stride 33 -> no
        // bank conflicts) and let lane q add up row q; 105 shuffle pairs become 21 stores + 32 loads
#pragma unroll
        for (int q = 0; q < NP; ++q) red[q * 33 + lane] = acc[q];
        __syncwarp();
        for (int q = lane; q < NP; q += 32) {
            double t = 0.0;
#pragma unroll 8
            for (int l2 = 0; l2 < 32; ++l2) t += red[q * 33 + l2];
            red[q * 33 + 32] = t;
        }
        __syncwarp();
        const double ll = red[2 * NS * 33 + 32];
        if (iteration > 0) improvement = ll - last;
        ++iteration;
        last = ll;
        if (!isfinite(ll)) { failed = true; break; }                     // assumption A5
        if (!(improvement > 0.1 && iteration < 1001)) break;             // last E-step's summaries discarded
        __syncwarp();
        // ---- per-sample responsibility sums -> weights (stats.py:206-211); the uniform
        //      component's sum is n_s - sum r0 - sum r1 (responsibilities sum to one)
        for (int s = 0; s < R; ++s) {
            const int a = fstart[s], e = fstart[s + 1];
            double t0 = 0.0, t1 = 0.0;
            for (int i = a + lane; i < e; i += 32) { t0 += rb0[i]; t1 += rb1[i]; }
            t0 = warp_sum(t0); t1 = warp_sum(t1);
            if (lane == 0) { wtmp[2 * s] = t0; wtmp[2 * s + 1] = t1; }
        }
        __syncwarp();
        for (int q = lane; q < 3 * R; q += 32) {        // one lane per (sample, component): logs in parallel
            const int s = q / 3, k = q - 3 * s;
            const int ns = fstart[s + 1] - fstart[s];
            if (ns > 0) {
                const double t0 = wtmp[2 * s], t1 = wtmp[2 * s + 1];
                const double t2 = fmax((double)ns - t0 - t1, 0.0);
                if (t0 + t1 + t2 != 0.0) {
                    const double vk = (k == 0 ? t0 : (k == 1 ? t1 : t2)) + 0.5;
                    par->logw[s][k] = log(vk / ((t0 + 0.5) + (t1 + 0.5) + (t2 + 0.5)));
                }
            }
        }
        // ---- M-step of the shared Gaussians: lane k owns component k (stats.py:202-205, A3/A4)
        const double step = 1.0 - (0.99 * pow(2.0 + (double)iteration, -1e-3));
        if (lane < 2) em_mstep<W>(par, lane, red, carry, step);
        __syncwarp();
        failed = par->ok == 0;
    }

    // ------------------------------------------------ results of the fit (stats.py:250-253, 277-280)
    if (failed) { if (lane == 0) { status[g] = YC_ST_FIT_FAILED; out.iters[slot] = iteration; out.loglik[slot] = last; } return; }
    __syncwarp();
    for (int i = lane; i < 2 * W; i += 32) out.mu[(size_t)slot * 2 * W + i] = par->mu[i / W][i % W];
    for (int i = lane; i < 2 * W * W; i += 32) out.cov[(size_t)slot * 2 * W * W + i] = par->cov[i / (W * W)][i % (W * W)];
    for (int i = lane; i < W; i += 32) { out.box[((size_t)slot * W + i) * 2] = par->lo[i]; out.box[((size_t)slot * W + i) * 2 + 1] = par->hi[i]; }
    for (int i = lane; i < R * 3; i += 32) out.logw[(size_t)slot * R * 3 + i] = par->logw[i / 3][i % 3];
    __syncwarp();                                   // fit_finalize overwrites par->logw[0] with the combined weights
    long long off = -1;
    if (lane == 0) {
        status[g] = YC_ST_TESTED;
        out.iters[slot] = iteration;
        out.loglik[slot] = last;
        fit_finalize<W>(par, fstart, R, Rc, p.fdr_threshold, out.preds + (size_t)slot * R * 3,
                        out.cw + (size_t)slot * 3, out.stats + (size_t)slot * 8);
        if (post != nullptr && out.stats[(size_t)slot * 8 + 2] < p.fdr_threshold) {
            off = (long long)atomicAdd(post_cursor, (unsigned long long)n);
            if (off + n > post_capacity) off = -2;
        }
        out.post_off[slot] = off;
    }
    off = __shfl_sync(YC_FULL, off, 0);
    if (off < 0) return;
    __syncwarp();
    // ------------------------------------------------ posteriors of every complete-window read (stats.py:494-499)
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(xs, cap, i, x);
        double q0 = 0.0, q1 = 0.0;
        bool inside = true;
#pragma unroll
        for (int a = 0; a < W; ++a) {
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int bb = 0; bb <= a; ++bb) {
                y0 = fma(par->linv[0][a * W + bb], x[bb] - par->mu[0][bb], y0);
                y1 = fma(par->linv[1][a * W + bb], x[bb] - par->mu[1][bb], y1);
            }
            q0 = fma(y0, y0, q0); q1 = fma(y1, y1, q1);
            inside &= (x[a] >= par->lo[a]) && (x[a] <= par->hi[a]);
        }
        const double l0 = par->cst[0] - 0.5 * q0 + par->logw[0][0];
        const double l1 = par->cst[1] - 0.5 * q1 + par->logw[0][1];
        const double l2 = inside ? par->ulogp + par->logw[0][2] : -INFINITY;
        const double mx = fmax(fmax(l0, l1), l2);
        const double e0 = exp(l0 - mx), e1 = exp(l1 - mx), e2 = exp(l2 - mx);
        const double inv = 1.0 / (e0 + e1 + e2);
        post[2 * (off + i)] = e1 * inv;
        post[2 * (off + i) + 1] = e2 * inv;
    }
}

static int fw_cap(const yc_batch_t &b) {
    int cap = (b.max_reads + 7) & ~7;
    return cap < 32 ? 32 : cap;
}

bool fit_warp_eligible(const yc_batch_t &b, const yc_params_t &p) {
    return p.window_size <= 5 && b.max_reads <= 1024 && b.max_cond_reads <= p.max_fit_depth && b.max_cond_reads > 0;
}

int fit_warp_smem_bytes(const yc_batch_t &b, const yc_params_t &p) {
    return fw_layout(p.window_size, b.n_samples, fw_cap(b)).total;
}

template <int W>
static int launch_fw(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                     const yc_fit_out_t &out, uint8_t *status, int8_t *labels, int max_fit_rows,
                     double *post, int64_t post_capacity, int64_t *post_cursor, cudaStream_t st) {
    const int cap = fw_cap(b);
    const int smem = fw_layout(W, b.n_samples, cap).total;
    cudaError_t e = cudaFuncSetAttribute(fit_warp_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    fit_warp_kernel<W><<<n_fit, 32, smem, st>>>(b, p, cap, fit_list, out, status, labels, max_fit_rows, post,
                                                (long long)post_capacity,
                                                reinterpret_cast<unsigned long long *>(post_cursor));
    return (int)cudaGetLastError();
}

int fit_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                      const yc_fit_out_t &out, uint8_t *status, int8_t *labels, int max_fit_rows,
                      double *post, int64_t post_capacity, int64_t *post_cursor, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    switch (p.window_size) {
        case 1: return launch_fw<1>(b, p, fit_list, n_fit, out, status, labels, max_fit_rows, post, post_capacity, post_cursor, st);
        case 3: return launch_fw<3>(b, p, fit_list, n_fit, out, status, labels, max_fit_rows, post, post_capacity, post_cursor, st);
        case 5: return launch_fw<5>(b, p, fit_list, n_fit, out, status, labels, max_fit_rows, post, post_capacity, post_cursor, st);
    }
    return YC_EINVAL;
}

}  // namespace yc
