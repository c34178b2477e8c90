// Kernel 2a (model initialisation) and 2b (multi-sample EM), one CTA per
// gated position.
//
// Reference semantics (file:line under /root/reference/yanocomp):
//   stats.py:266-269   per-REPLICATE sub-sampling to max_fit_depth
//   stats.py:81-124    pm.Kmeans(2,'kmeans++',n_init=3).fit(X,100,0.5),
//                      distances to the centroids, MAD outlier labelling
//   stats.py:127-191   ML Gaussians per label, per-sample initial weights,
//                      uniform box from the outlier rows
//   stats.py:194-225   the EM schedule (iteration 0 is E-only; inertia
//                      1-(0.99*(2+it)^-1e-3); pseudocount 0.5; stop when the
//                      log-likelihood gain <= 0.1, NaN, or after 1001 E-steps)
// pomegranate behaviour is restated as documented in oracle/yanocomp_oracle.py
// (ASSUMPTIONS A1-A7).  The initialisation is arithmetic on exact binary16
// values (sums of values and of pairwise products are exact in FP64), so it
// is written with separately rounded operations to reproduce the oracle bit
// for bit; EM runs in FP64 with fused multiply-adds.
#include "yc_common.cuh"
#include "yc_internal.h"
#include "yc_stats.cuh"
#include "yc_fit_common.cuh"

namespace yc {


__host__ __device__ inline int cap_pow2_bytes(int cap) {
    int P = 32;
    while (P < cap) P <<= 1;
    return (P + P / 16) * (int)sizeof(double);          // padded sort layout (YC_PAD16)
}

struct FitSmem {
    int xs_off, srow_off, lab_off, sort_off, mask_off, wordoff_off, cnt_off, fstart_off,
        red_off, hist_off, sel_off, par_off, part_off, wsum_off, total;
};

// which = 0: init kernel (needs sort buffer + labels), 1: EM / gtest kernels
__host__ __device__ inline FitSmem fit_smem_layout(int W, int R, int max_reads, int cap, int which) {
    FitSmem s;
    const int NS = 1 + W + W * (W + 1) / 2;
    int o = 0;
    s.sort_off = o;    o += (which == 0 ? cap_pow2_bytes(cap) : 0);
    s.red_off = o;     o += 64 * YC_NWARPS * (int)sizeof(double);
    s.par_off = o;     o += (2 * (W + 2 * W * W + 2) + 2 * W + 3 * YC_MAX_SAMPLES + 8) * (int)sizeof(double);
    s.part_off = o;    o += (which == 1 ? (YC_NWARPS + 1) * (2 * NS + 1) : 0) * (int)sizeof(double);
    s.wsum_off = o;    o += (which == 1 ? YC_NWARPS * YC_MAX_SAMPLES * 3 : 0) * (int)sizeof(double);
    s.sel_off = o;     o += 4 * (int)sizeof(unsigned long long);
    s.mask_off = o;    o += ((max_reads + 31) / 32 + 1) * (int)sizeof(uint32_t);
    s.wordoff_off = o; o += ((max_reads + 31) / 32 + 1) * (int)sizeof(int);
    s.cnt_off = o;     o += (W + 1) * R * (int)sizeof(int);
    s.fstart_off = o;  o += (YC_MAX_SAMPLES + 2) * (int)sizeof(int);
    s.hist_off = o;    o += 256 * (int)sizeof(int);
    s.xs_off = o;      o += W * cap * (int)sizeof(uint16_t);
    o = (o + 3) & ~3;
    s.srow_off = o;    o += cap;
    s.lab_off = o;     o += (which == 0 ? cap : 0);
    s.total = (o + 15) & ~15;
    return s;
}

// ---------------------------------------------------------------------------
// shared front end: complete rows -> per-sample sub-sampling -> compaction
// ---------------------------------------------------------------------------
template <int W>
struct FitCtx {
    Window<W> w;
    uint16_t *xs;      // [W][cap] binary16 bits of the fit rows, pooled order
    uint8_t *srow;     // [cap] sample of each fit row
    int *fstart;       // [R+1] first fit row of each sample
    int *cnt;          // window_scan counters
    uint32_t *mask;
    int *wordoff;
    int cap, n;        // n = fit rows
    int n_c, n_t;      // complete rows per condition (before sub-sampling)
};

template <int W>
__device__ __forceinline__ bool fit_front_end(const yc_batch_t &b, const yc_params_t &p, int64_t g,
                                              unsigned char *smem, const FitSmem &L, int cap,
                                              bool subsample, FitCtx<W> &c) {
    const int R = b.n_samples, tid = threadIdx.x;
    c.xs = reinterpret_cast<uint16_t *>(smem + L.xs_off);
    c.srow = reinterpret_cast<uint8_t *>(smem + L.srow_off);
    c.fstart = reinterpret_cast<int *>(smem + L.fstart_off);
    c.cnt = reinterpret_cast<int *>(smem + L.cnt_off);
    c.mask = reinterpret_cast<uint32_t *>(smem + L.mask_off);
    c.wordoff = reinterpret_cast<int *>(smem + L.wordoff_off);
    c.cap = cap;
    int *hist = reinterpret_cast<int *>(smem + L.hist_off);
    unsigned long long *sel = reinterpret_cast<unsigned long long *>(smem + L.sel_off);
    int *iscr = reinterpret_cast<int *>(smem + L.red_off);

    if (!window_setup<W>(b, p, g, c.w)) return false;
    for (int i = tid; i < (W + 1) * R; i += YC_BLOCK) c.cnt[i] = 0;
    __syncthreads();
    window_scan<W>(c.w, R, c.mask, c.cnt);
    __syncthreads();
    c.n_c = 0; c.n_t = 0;
    for (int s = 0; s < R; ++s) { if (s < b.n_cntrl) c.n_c += c.cnt[W * R + s]; else c.n_t += c.cnt[W * R + s]; }
    if (subsample) {
        for (int s = 0; s < R; ++s) {
            const int have = c.cnt[W * R + s];
            subsample_rows(c.mask, c.w.samp_start[s], c.w.samp_start[s + 1], have, p.max_fit_depth,
                           c.w.pkey, YC_STREAM_FIT_SAMPLE0 + s, hist, sel);
        }
    }
    if (tid == 0) {
        int acc = 0;
        for (int s = 0; s < R; ++s) {
            c.fstart[s] = acc;
            const int have = c.cnt[W * R + s];
            acc += subsample ? min(have, p.max_fit_depth) : have;
        }
        c.fstart[R] = acc;
    }
    const int nwords = (c.w.n_reads + 31) >> 5;
    c.n = mask_word_prefix(c.mask, nwords, c.wordoff, iscr);
    return true;
}

// copy the selected rows into shared memory (only when they fit: n <= cap)
template <int W>
__device__ __forceinline__ void fit_gather(const FitCtx<W> &c, int R) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwords = (c.w.n_reads + 31) >> 5;
    for (int wd = wid; wd < nwords; wd += YC_NWARPS) {
        const uint32_t m = c.mask[wd];
        if ((m >> lane) & 1u) {
            const int r = wd * 32 + lane;
            const int rank = c.wordoff[wd] + __popc(m & ((1u << lane) - 1u));
#pragma unroll
            for (int j = 0; j < W; ++j) c.xs[j * c.cap + rank] = __ldg(c.w.col[j] + r);
            int s = 0;
            while (s + 1 < R && r >= c.w.samp_start[s + 1]) ++s;
            c.srow[rank] = (uint8_t)s;
        }
    }
    __syncthreads();
}

template <int W>
__device__ __forceinline__ void load_row(const FitCtx<W> &c, int i, double (&x)[W]) {
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = h2d(c.xs[j * c.cap + i]);
}

// Row index of the k-th row (0-based, pooled order) whose label is not 2 -- the mapping from an
// index into X[~outlier_mask] (stats.py:104) back to the fit rows.  Every thread owns the
// contiguous chunk [i0c, i1c); scratch: YC_NWARPS + 1 ints.  All threads return the row.
__device__ __forceinline__ int nth_inlier(const int8_t *lab, int i0c, int i1c, int k, int *scratch) {
    int loc = 0;
    for (int i = i0c; i < i1c; ++i) loc += lab[i] != 2;
    int total;
    const int before = block_excl_scan_i(loc, scratch, total);
    if (k >= before && k < before + loc) {
        int seen = before;
        for (int i = i0c; i < i1c; ++i)
            if (lab[i] != 2) { if (seen == k) { scratch[YC_NWARPS] = i; break; } ++seen; }
    }
    __syncthreads();
    const int row = scratch[YC_NWARPS];
    __syncthreads();
    return row;
}

// ---------------------------------------------------------------------------
// Kernel 2a
// ---------------------------------------------------------------------------
// One Lloyd pass over the fit rows: nearest of the two centroids per row, column sums and size of cluster 1, total
// distance (acc[0..W) | acc[W] | acc[W + 1], per thread).
//  * The nearer centroid is decided from d_k = |x|^2 - 2 x.c_k + |c_k|^2 (|x|^2 precomputed, FMA chains: W + 2
//    operations per centroid instead of 3 W - 1 separately rounded ones); only when the two lie within the error bound
//    of that form (2^-44 of the magnitudes involved is > 64 times the bound) are the NumPy-order distances compared,
//    which is the comparison the oracle makes (A2).  The cluster sums are sums of binary16 values: exact either way.
//  * U rows per step, branch-free (predicates as 0/1 weights, one rare branch for the exact comparison): at 2 CTAs
//    per SM the pass is bound by dependency latency, and rows are the only independent work there is.
#define YC_LLOYD_GUARD 5.6843418860808015e-14     /* 2^-44 */
// (measured on cfg4, init kernel: U = 2 136.6 ms at the time, later U = 3 105.7, U = 4 103.4, U = 6 107.9 ms)
#ifndef YC_LLOYD_U
#define YC_LLOYD_U 4
#endif
template <int W, int U>
__device__ __forceinline__ void lloyd_pass(const FitCtx<W> &c, const double *__restrict__ xx, const int8_t *__restrict__ lab,
                                           const bool db, const int n, const double (&cen)[2 * W], double (&acc)[W + 2]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < W + 2; ++k) acc[k] = 0.0;
    double m0[W], m1[W], cc0 = 0.0, cc1 = 0.0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        m0[j] = -2.0 * cen[j]; m1[j] = -2.0 * cen[W + j];
        cc0 = fma(cen[j], cen[j], cc0); cc1 = fma(cen[W + j], cen[W + j], cc1);
    }
    const double ccmax = fmax(cc0, cc1);
#pragma unroll 1
    for (int i0 = tid; i0 < n; i0 += U * YC_BLOCK) {
        double x[U][W], d0[U], d1[U];
        bool k[U], use[U], need = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * YC_BLOCK;
            use[u] = i < n;
            const int ii = use[u] ? i : i0;
            if (db && lab[ii] == 2) use[u] = false;
            load_row<W>(c, ii, x[u]);
            const double xi = xx[ii];
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int j = 0; j < W; ++j) { t0 = fma(x[u][j], m0[j], t0); t1 = fma(x[u][j], m1[j], t1); }
            d0[u] = (xi + cc0) + t0; d1[u] = (xi + cc1) + t1;
            const double diff = d1[u] - d0[u];
            k[u] = diff < 0.0;
            need |= fabs(diff) <= YC_LLOYD_GUARD * ((xi + ccmax) + (fabs(t0) + fabs(t1)));
        }
        if (need) {
#pragma unroll
            for (int u = 0; u < U; ++u) k[u] = sqdist<W>(x[u], cen + W) < sqdist<W>(x[u], cen);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double w1 = (use[u] && k[u]) ? 1.0 : 0.0;
            const double dsel = k[u] ? d1[u] : d0[u];
            acc[W] += use[u] ? dsel : 0.0;
            acc[W + 1] += w1;
#pragma unroll
            for (int j = 0; j < W; ++j) acc[j] = fma(w1, x[u][j], acc[j]);
        }
    }
}
#ifndef YC_INITC_MINBLOCKS
#define YC_INITC_MINBLOCKS 1
#endif
template <int W>
__global__ void __launch_bounds__(YC_BLOCK, YC_INITC_MINBLOCKS)
gmm_init_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
                const int32_t *__restrict__ fit_list, const int32_t *__restrict__ fit_count, const yc_fit_out_t out,
                uint8_t *__restrict__ status, int8_t *__restrict__ labels_out, const int max_fit_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FitSmem L = fit_smem_layout(W, R, b.max_reads, cap, 0);
    double *sortbuf = reinterpret_cast<double *>(smem_raw + L.sort_off);
    double *red = reinterpret_cast<double *>(smem_raw + L.red_off);
    double *par = reinterpret_cast<double *>(smem_raw + L.par_off);
    int8_t *lab = reinterpret_cast<int8_t *>(smem_raw + L.lab_off);
    int *ired = reinterpret_cast<int *>(red);
    const int slot = blockIdx.x;
    if (fit_count != nullptr && slot >= *fit_count) return;            // the grid is an upper bound: the count lives on the device
    const int64_t g = fit_list[slot];
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT;

    FitCtx<W> c;
    if (!fit_front_end<W>(b, p, g, smem_raw, L, cap, true, c) || c.n > cap || c.n < 1) {
        if (tid == 0) status[g] = YC_ST_FIT_FAILED;
        return;
    }
    fit_gather<W>(c, R);
    const int n = c.n;
    const uint64_t pkey = c.w.pkey;

    // ---------------- --outlier-method dbscan (stats.py:91-104, assumption A11): a row is an outlier when it
    // is neither a core point (>= n/4 rows within eps, itself included) nor within eps of one.  Squared
    // distances of binary16 rows are exact in FP64, so `d2 <= eps^2` is the comparison sklearn makes.
    const bool db = p.add_uniform && p.outlier_method == YC_OUTLIER_DBSCAN;
    int n_in = n;                                          // rows k-means is fitted on
    if (db) {
        const double eps2 = mul_rn(p.dbscan_eps, p.dbscan_eps), min_samples = mul_rn(0.25, (double)n);
        for (int i = tid; i < n; i += YC_BLOCK) {
            double x[W];
            load_row<W>(c, i, x);
            int cnt = 0;
            for (int j = 0; j < n; ++j) {
                double y[W];
                load_row<W>(c, j, y);
                cnt += sqdist<W>(x, y) <= eps2;
            }
            lab[i] = (double)cnt >= min_samples ? 0 : 3;   // 0: core point, 3: not core (resolved below)
        }
        __syncthreads();
        int noise = 0;
        for (int i = tid; i < n; i += YC_BLOCK) {
            if (lab[i] == 3) {
                double x[W];
                load_row<W>(c, i, x);
                bool reach = false;
                for (int j = 0; j < n && !reach; ++j) {
                    if (lab[j] == 0) {                     // lab[j] of other non-core rows may turn 1 meanwhile: never 0
                        double y[W];
                        load_row<W>(c, j, y);
                        reach = sqdist<W>(x, y) <= eps2;
                    }
                }
                lab[i] = reach ? 1 : 2;                    // 1: border point (inlier), 2: outlier
                noise += !reach;
            }
        }
        noise = block_sum_i(noise, ired);                  // also orders the lab[] writes
        n_in = n - noise;
    }

    // column totals of the rows k-means sees (exact), so that Lloyd accumulates one cluster only
    double sumx[W];
    {
#pragma unroll
        for (int j = 0; j < W; ++j) sumx[j] = 0.0;
        for (int i = tid; i < n; i += YC_BLOCK) {
            if (db && lab[i] == 2) continue;
            double x[W];
            load_row<W>(c, i, x);
#pragma unroll
            for (int j = 0; j < W; ++j) sumx[j] += x[j];
        }
        block_sum<W>(sumx, red);
    }

    // |x|^2 of every fit row, parked in the (still unused) MAD sort buffer: Lloyd decides the nearer centroid from
    // d_k = |x|^2 - 2 x.c_k + |c_k|^2 (FMA chains, W + 2 operations per centroid instead of 3 W - 1 separately rounded
    // ones) and falls back to the NumPy-order distances only when the two are closer than the error bound of that form
    double *xx = sortbuf;
    for (int i = tid; i < n; i += YC_BLOCK) {
        double x[W];
        load_row<W>(c, i, x);
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) t = fma(x[j], x[j], t);
        xx[i] = t;
    }
    __syncthreads();

    // ---------------- k-means++ x3 + Lloyd on the inliers  (assumption A2)
    double best[2 * W], best_total = INFINITY;
    bool have_best = false;
    // every thread owns a contiguous chunk of rows (the cumulative sum of the D^2 sampling runs in row order).  The chunk
    // length is rounded up to 2 (mod 4): with binary16 rows in shared memory the lanes of a warp then start 17, 19, ...
    // 32-bit words apart -- distinct banks -- where the natural length for 4 000 rows, 32, put them on two banks
    // (16-way conflicts on every load of the six seeding passes)
    int C = (n + YC_BLOCK - 1) / YC_BLOCK;
    while ((C & 3) != 2) ++C;
    const int i0c = min(tid * C, n), i1c = min(i0c + C, n);
    for (int restart = 0; restart < 3 && n_in > 0; ++restart) {
        double cen[2 * W];
        {
            const double u0 = rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 0);
            int i0 = min((int)(u0 * (double)n_in), n_in - 1);
            if (db) i0 = nth_inlier(lab, i0c, i1c, i0, ired);
            double x[W];
            load_row<W>(c, i0, x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[j] = x[j];
        }
        {   // second centre: D^2 sampling through the cumulative sum
            double loc = 0.0;
#pragma unroll 2
            for (int i = i0c; i < i1c; ++i) {
                if (db && lab[i] == 2) continue;
                double x[W]; load_row<W>(c, i, x); loc = add_rn(loc, sqdist<W>(x, cen));
            }
            // exclusive scan of the chunk sums over the CTA
            double inc = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(YC_FULL, inc, o); if (lane >= o) inc = add_rn(inc, t); }
            if (lane == 31) red[wid] = inc;
            __syncthreads();
            double base = 0.0, total = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < YC_NWARPS; ++w2) { if (w2 < wid) base = add_rn(base, red[w2]); total = add_rn(total, red[w2]); }
            __syncthreads();
            double run = add_rn(base, sub_rn(inc, loc));
            const double u1 = rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 1);
            const double target = mul_rn(u1, total);
            int below = 0;
#pragma unroll 2
            for (int i = i0c; i < i1c; ++i) {
                if (db && lab[i] == 2) continue;
                double x[W]; load_row<W>(c, i, x); run = add_rn(run, sqdist<W>(x, cen)); below += (run <= target);
            }
            below = block_sum_i(below, ired);
            int i1 = min(below, n_in - 1);
            if (db) i1 = nth_inlier(lab, i0c, i1c, i1, ired);
            double x[W];
            load_row<W>(c, i1, x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[W + j] = x[j];
        }
        int iteration = 0;
        double improvement = INFINITY, last = INFINITY, total = INFINITY;
        double acc[W + 2];                                   // column sums of cluster 1, total distance, size of cluster 1
        bool have_sums = false;
        // the centroids of the next pass and the total distance are formed by 2 W + 1 threads between the two barriers
        // of the reduction (every thread used to run the 2 W IEEE divisions: ~15 % of an iteration's instructions)
        double *cs = par;
        while (improvement > 0.5 && iteration < 101) {
            if (have_sums) {
#pragma unroll
                for (int q = 0; q < 2 * W; ++q) cen[q] = cs[q];
            }
            lloyd_pass<W, YC_LLOYD_U>(c, xx, lab, db, n, cen, acc);
            warp_sum_array<W + 2>(acc);
#pragma unroll
            for (int k = 0; k < W + 2; ++k) {
                if (lane == 0) red[k * YC_NWARPS + wid] = acc[k];
            }
            __syncthreads();
            if (tid < 2 * W) {
                // cluster 0 = column totals - cluster 1 (sums of binary16 values: exact either way); an empty cluster
                // keeps its centroid (A2)
                const int j = tid < W ? tid : tid - W;
                double s1 = 0.0, n1 = 0.0, sj = 0.0, keep = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < YC_NWARPS; ++w2) { s1 += red[j * YC_NWARPS + w2]; n1 += red[(W + 1) * YC_NWARPS + w2]; }
#pragma unroll
                for (int q = 0; q < W; ++q) if (q == j) sj = sumx[q];
#pragma unroll
                for (int q = 0; q < 2 * W; ++q) if (q == tid) keep = cen[q];
                const double den = tid < W ? (double)n_in - n1 : n1;
                const double num = tid < W ? sub_rn(sj, s1) : s1;
                cs[tid] = den > 0.0 ? div_rn(num, den) : keep;
            } else if (tid == 2 * W) {
                double t = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < YC_NWARPS; ++w2) t += red[W * YC_NWARPS + w2];
                cs[2 * W] = t;
            }
            __syncthreads();
            acc[W] = cs[2 * W];
            have_sums = true;
            total = acc[W];
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
    if (!have_best && n_in > 0) {
        if (tid == 0) status[g] = YC_ST_FIT_FAILED;
        return;
    }
    // ---------------- nearest-centroid labels; outliers: MAD rule (stats.py:114-122, A1), the DBSCAN
    // mask (stats.py:122), or none with --no-uniform
    int n_out = 0;
    if (db) {
        n_out = n - n_in;
        if (n_in > 0) {                                    // every row an outlier: all labels stay 2 (stats.py:100-101)
            for (int i = tid; i < n; i += YC_BLOCK) {
                if (lab[i] == 2) continue;
                double x[W];
                load_row<W>(c, i, x);
                lab[i] = (int8_t)(sqdist<W>(x, best + W) < sqdist<W>(x, best));
            }
        }
        __syncthreads();
    } else if (!p.add_uniform) {
        for (int i = tid; i < n; i += YC_BLOCK) {
            double x[W];
            load_row<W>(c, i, x);
            lab[i] = (int8_t)(sqdist<W>(x, best + W) < sqdist<W>(x, best));
        }
        __syncthreads();
    } else {
    const int P = max(32, next_pow2(n));
    for (int i = tid; i < P; i += YC_BLOCK) {
        double dmin = INFINITY;
        if (i < n) {
            double x[W];
            load_row<W>(c, i, x);
            const double d0 = sqdist<W>(x, best), d1 = sqdist<W>(x, best + W);
            lab[i] = (int8_t)(d1 < d0);
            dmin = d1 < d0 ? d1 : d0;
        }
        sortbuf[YC_PAD16(i)] = dmin;
    }
    __syncthreads();
    block_sort_segments_regs(sortbuf, P, 1);
    const double med = (n & 1) ? sortbuf[YC_PAD16(n >> 1)] : mul_rn(add_rn(sortbuf[YC_PAD16((n >> 1) - 1)], sortbuf[YC_PAD16(n >> 1)]), 0.5);
    __syncthreads();
    // MAD = median of |d - med| WITHOUT a second sort: in the sorted distances the deviations are V-shaped,
    // i.e. two ascending runs (left of the median walked outwards, right of it), so the order statistics of
    // their union follow from merge ranks found by binary search (same selection as init_warp_kernel)
    double *wsel = par + 4;
    const int m_le = upper_bound_pad(sortbuf, 0, n, med);        // elements <= med form the left run
    const int k_hi = n >> 1, k_lo = (n & 1) ? k_hi : k_hi - 1;
    for (int i = tid; i < n; i += YC_BLOCK) {
        const double dv = fabs(sub_rn(sortbuf[YC_PAD16(i)], med));
        int rank;
        if (i < m_le) {                                    // left run, index m_le-1-i; count right-run elements < dv
            int lo = 0, hi = n - m_le;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (fabs(sub_rn(sortbuf[YC_PAD16(m_le + mid)], med)) < dv) lo = mid + 1; else hi = mid; }
            rank = (m_le - 1 - i) + lo;
        } else {                                           // right run, index i-m_le; count left-run elements <= dv
            int lo = 0, hi = m_le;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (fabs(sub_rn(sortbuf[YC_PAD16(m_le - 1 - mid)], med)) <= dv) lo = mid + 1; else hi = mid; }
            rank = (i - m_le) + lo;
        }
        if (rank == k_lo) wsel[0] = dv;
        if (rank == k_hi) wsel[1] = dv;
    }
    __syncthreads();
    const double mad = (n & 1) ? wsel[1] : mul_rn(add_rn(wsel[0], wsel[1]), 0.5);
    const double thr = mul_rn(p.outlier_factor, mad);
    __syncthreads();
    for (int i = tid; i < n; i += YC_BLOCK) {
        double x[W];
        load_row<W>(c, i, x);
        const double d0 = sqdist<W>(x, best), d1 = sqdist<W>(x, best + W);
        if ((d1 < d0 ? d1 : d0) > thr) { lab[i] = 2; ++n_out; }
    }
    n_out = block_sum_i(n_out, ired);          // also orders the lab[] writes
    }
    if (labels_out) for (int i = tid; i < n && i < max_fit_rows; i += YC_BLOCK)
        labels_out[(size_t)slot * max_fit_rows + i] = lab[i];

    // ---------------- ML Gaussians per label (stats.py:148-152; diag: column mean / population std,
    // N(0, 1) per column for an empty cluster, stats.py:153-166)
    for (int k = 0; k < 2; ++k) {
        double S[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) S[q] = 0.0;
        for (int i = tid; i < n; i += YC_BLOCK) {
            if (lab[i] == k) {
                double x[W];
                load_row<W>(c, i, x);
                S[0] += 1.0;
                int q = 1 + W;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    S[1 + a] += x[a];
#pragma unroll
                    for (int c2 = a; c2 < W; ++c2) S[q++] += x[a] * x[c2];
                }
            }
        }
        block_sum<NS>(S, red);
        if (p.covariance_type == YC_COV_DIAG) {
            // numpy's .std(): sqrt(mean((x - mean)^2)) from the deviations (exactly 0 for a constant column)
            double mean[W], dev[W];
#pragma unroll
            for (int a = 0; a < W; ++a) { mean[a] = S[0] > 0.0 ? div_rn(S[1 + a], S[0]) : 0.0; dev[a] = 0.0; }
            for (int i = tid; i < n; i += YC_BLOCK) {
                if (lab[i] == k) {
                    double x[W];
                    load_row<W>(c, i, x);
#pragma unroll
                    for (int a = 0; a < W; ++a) { const double t = sub_rn(x[a], mean[a]); dev[a] += t * t; }
                }
            }
            block_sum<W>(dev, red);
            if (tid == k * 32) {
                for (int a = 0; a < W; ++a) {
                    const double sg = S[0] > 0.0 ? sqrt(dev[a] / S[0]) : 1.0;
                    out.mu[((size_t)slot * 2 + k) * W + a] = mean[a];
                    for (int c2 = 0; c2 < W; ++c2)
                        out.cov[((size_t)slot * 2 + k) * W * W + a * W + c2] = (a == c2) ? sg * sg : 0.0;
                }
                par[k] = 1.0;                  // no factorisation to fail: sigma = 0 shows up as a non-finite likelihood
            }
        } else if (tid == k * 32) {
            double mu[W], cov[W * W];
            const bool ok = ml_gaussian<W>(S, mu, cov);
            for (int a = 0; a < W; ++a) out.mu[((size_t)slot * 2 + k) * W + a] = mu[a];
            for (int a = 0; a < W * W; ++a) out.cov[((size_t)slot * 2 + k) * W * W + a] = cov[a];
            par[k] = ok ? 1.0 : 0.0;
        }
    }
    // ---------------- uniform box (stats.py:178-187): outlier rows, or all rows if none
    {
        double lo[W], hi[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { lo[j] = INFINITY; hi[j] = -INFINITY; }
        for (int i = tid; i < n; i += YC_BLOCK) {
            if (n_out == 0 || lab[i] == 2) {
                double x[W];
                load_row<W>(c, i, x);
#pragma unroll
                for (int j = 0; j < W; ++j) { lo[j] = fmin(lo[j], x[j]); hi[j] = fmax(hi[j], x[j]); }
            }
        }
        bool box_ok = true;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            lo[j] = block_min1(lo[j], red);
            hi[j] = block_max1(hi[j], red);
            if (p.add_uniform) box_ok &= hi[j] > lo[j];
            if (tid == 0) { out.box[((size_t)slot * W + j) * 2] = lo[j]; out.box[((size_t)slot * W + j) * 2 + 1] = hi[j]; }
        }
        if (tid == 0) par[2] = box_ok ? 1.0 : 0.0;
    }
    // ---------------- per-sample initial weights (stats.py:173-177)
    for (int s = wid; s < R; s += YC_NWARPS) {
        const int a = c.fstart[s], e = c.fstart[s + 1];
        int k0 = 0, k1 = 0, k2 = 0;
        for (int i = a + lane; i < e; i += 32) { const int l = lab[i]; k0 += l == 0; k1 += l == 1; k2 += l == 2; }
        k0 = warp_sum_i(k0); k1 = warp_sum_i(k1); k2 = warp_sum_i(k2);
        if (lane == 0) {
            double *lw = out.logw + ((size_t)slot * R + s) * 3;
            if (e > a) {
                const double den = (double)(e - a) + (p.add_uniform ? 1.5 : 1.0);
                lw[0] = log(((double)k0 + 0.5) / den);
                lw[1] = log(((double)k1 + 0.5) / den);
                lw[2] = p.add_uniform ? log(((double)k2 + 0.5) / den) : -INFINITY;
            } else { lw[0] = lw[1] = lw[2] = 0.0; }
        }
    }
    __syncthreads();
    if (tid == 0 && !(par[0] != 0.0 && par[1] != 0.0 && par[2] != 0.0)) status[g] = YC_ST_FIT_FAILED;
}

// ---------------------------------------------------------------------------
// Kernel 2b
// ---------------------------------------------------------------------------
// rows per thread whose E-steps run side by side in gmm_em_kernel.  Measured on cfg4 (B200): U = 1 101.0 ms, U = 2 at 168
// registers / 3 CTAs per SM 103.0 ms, U = 2 at 240 registers / 2 CTAs per SM 111.9 ms -- the kernel is not latency-bound
#ifndef YC_EM_U
#define YC_EM_U 1
#endif
#ifndef YC_EMC_MINBLOCKS
#define YC_EMC_MINBLOCKS 3
#endif
template <int W>
__global__ void __launch_bounds__(YC_BLOCK, YC_EMC_MINBLOCKS)
gmm_em_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
              const int32_t *__restrict__ fit_list, const int32_t *__restrict__ fit_count, const yc_fit_out_t out,
              uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FitSmem L = fit_smem_layout(W, R, b.max_reads, cap, 1);
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT, NP = 2 * NS + 1;
    EmPar<W> *par = reinterpret_cast<EmPar<W> *>(smem_raw + L.par_off);
    double *part = reinterpret_cast<double *>(smem_raw + L.part_off);    // [NWARPS][NP] + carry [2*NS]
    double *carry = part + YC_NWARPS * NP;
    double *wsum = reinterpret_cast<double *>(smem_raw + L.wsum_off);    // [NWARPS][R][3]
    const int slot = blockIdx.x;
    if (fit_count != nullptr && slot >= *fit_count) return;            // the grid is an upper bound: the count lives on the device
    const int64_t g = fit_list[slot];
    if (status[g] != YC_ST_GATED) return;                                // init failed

    FitCtx<W> c;
    if (!fit_front_end<W>(b, p, g, smem_raw, L, cap, true, c) || c.n > cap) {
        if (tid == 0) status[g] = YC_ST_FIT_FAILED;
        return;
    }
    fit_gather<W>(c, R);
    const int n = c.n;

    // load the initial model
    for (int i = tid; i < 2 * W; i += YC_BLOCK) par->mu[i / W][i % W] = out.mu[(size_t)slot * 2 * W + i];
    for (int i = tid; i < 2 * W * W; i += YC_BLOCK) par->cov[i / (W * W)][i % (W * W)] = out.cov[(size_t)slot * 2 * W * W + i];
    for (int i = tid; i < W; i += YC_BLOCK) { par->lo[i] = out.box[((size_t)slot * W + i) * 2]; par->hi[i] = out.box[((size_t)slot * W + i) * 2 + 1]; }
    for (int i = tid; i < R * 3; i += YC_BLOCK) par->logw[i / 3][i % 3] = out.logw[(size_t)slot * R * 3 + i];
    for (int i = tid; i < 2 * NS; i += YC_BLOCK) carry[i] = 0.0;
    if (tid == 0) par->ok = 1;
    __syncthreads();
    if (tid < 64 && lane == 0) {
        const int k = wid;
        double logdet;
        if (!cholesky_inverse<W>(par->cov[k], par->linv[k], logdet)) par->ok = 0;
        par->cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
    }
    if (tid == 64) {
        double s = 0.0;
        for (int j = 0; j < W; ++j) s += log(par->hi[j] - par->lo[j]);
        par->ulogp = -s;
    }
    __syncthreads();

    // the uniform box does not move during EM (A5): whether a row lies inside it (bounds inclusive) is decided once
    // and kept in bit 7 of the row's sample byte
    for (int i = tid; i < n; i += YC_BLOCK) {
        double x[W];
        load_row<W>(c, i, x);
        bool in = true;
#pragma unroll
        for (int a = 0; a < W; ++a) in &= (x[a] >= par->lo[a]) && (x[a] <= par->hi[a]);
        if (in) c.srow[i] |= 0x80;
    }
    __syncthreads();
    int iteration = 0;
    double improvement = INFINITY, last = 0.0;
    bool failed = par->ok == 0;
    while (!failed && improvement > 0.1 && iteration < 1001) {
        // ---- E-step + sufficient statistics (GeneralMixtureModel.summarize, A6)
        double acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = 0.0;
        for (int q = lane; q < R * 3; q += 32) wsum[wid * YC_MAX_SAMPLES * 3 + q] = 0.0;
        __syncwarp();
        double mu0[W], mu1[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { mu0[j] = par->mu[0][j]; mu1[j] = par->mu[1][j]; }
        double seprod = 1.0;
        int cur_s = -1;
        double ps0 = 0.0, ps1 = 0.0, ps2 = 0.0;
        // YC_EM_U 32-row chunks per warp and step (1 by default, see above): E-steps first, then the statistics row by
        // row in the order of the one-row loop (same results for every U).
        for (int base0 = wid * 32; base0 < n; base0 += YC_EM_U * YC_NWARPS * 32) {
            double x[YC_EM_U][W], r0[YC_EM_U], r1[YC_EM_U], r2[YC_EM_U];
            int sidx[YC_EM_U];
#pragma unroll
            for (int u = 0; u < YC_EM_U; ++u) {
                const int i = base0 + u * YC_NWARPS * 32 + lane;
                const bool valid = i < n;
                const int ii = valid ? i : 0;
                load_row<W>(c, ii, x[u]);
                const int sv = c.srow[ii];
                const int s = sv & 0x7f;
                const bool inside = (sv & 0x80) != 0;
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    double y0 = 0.0, y1 = 0.0;
#pragma unroll
                    for (int bb = 0; bb <= a; ++bb) {
                        y0 = fma(par->linv[0][a * W + bb], x[u][bb] - mu0[bb], y0);
                        y1 = fma(par->linv[1][a * W + bb], x[u][bb] - mu1[bb], y1);
                    }
                    q0 = fma(y0, y0, q0);
                    q1 = fma(y1, y1, q1);
                }
                const double l0 = par->cst[0] - 0.5 * q0 + par->logw[s][0];
                const double l1 = par->cst[1] - 0.5 * q1 + par->logw[s][1];
                const double l2 = (inside && p.add_uniform) ? par->ulogp + par->logw[s][2] : -INFINITY;
                // arg-max by comparisons; exp(l_max - l_max) = 1 exactly, so only the two other terms need an exp
                // (the same evaluation as em_warp_kernel, yc_fit_warp.cu)
                const bool m0 = l0 >= l1 && l0 >= l2, m1 = !m0 && l1 >= l2;
                const double m = m0 ? l0 : (m1 ? l1 : l2);
                double ea, eb;
                exp_nonpos2((m0 ? l1 : l0) - m, ((m0 || m1) ? l2 : l1) - m, ea, eb);
                const double e0 = m0 ? 1.0 : ea, e1 = m0 ? ea : (m1 ? 1.0 : eb), e2 = (m0 || m1) ? eb : 1.0;
                const double se = e0 + e1 + e2;
                const double inv = rcp_1to3(se);
                r0[u] = valid ? e0 * inv : 0.0; r1[u] = valid ? e1 * inv : 0.0; r2[u] = valid ? e2 * inv : 0.0;
                sidx[u] = valid ? s : -1;
                acc[2 * NS] += valid ? m : 0.0;
                seprod *= valid ? se : 1.0;                  // one log per thread: log prod se_i, 1 <= se_i <= 3
            }
#pragma unroll
            for (int u = 0; u < YC_EM_U; ++u) {
                const int base = base0 + u * YC_NWARPS * 32;
                if (base >= n) break;                        // warp-uniform
                acc[0] += r0[u]; acc[NS] += r1[u];
                int q = 1 + W;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    const double xa0 = r0[u] * x[u][a], xa1 = r1[u] * x[u][a];
                    acc[1 + a] += xa0; acc[NS + 1 + a] += xa1;
#pragma unroll
                    for (int bb = a; bb < W; ++bb) { acc[q] = fma(xa0, x[u][bb], acc[q]); acc[NS + q] = fma(xa1, x[u][bb], acc[NS + q]); ++q; }
                }
                // per-sample responsibility sums (gmm.summaries).  Rows are ordered by sample, so nearly every 32-row chunk
                // lies inside one sample: those only add to per-lane partials of the current sample `cur_s` (warp-uniform),
                // reduced over the warp when the sample changes; a chunk that straddles samples is reduced directly.
                const int s_lo = __shfl_sync(YC_FULL, sidx[u], 0);
                const int s_hi = c.srow[min(base + 31, n - 1)] & 0x7f;
                if (s_lo != cur_s || s_hi != cur_s) {
                    if (cur_s >= 0) {
                        const double t0 = warp_sum(ps0), t1 = warp_sum(ps1), t2 = warp_sum(ps2);
                        if (lane == 0) { double *ws = wsum + (wid * YC_MAX_SAMPLES + cur_s) * 3; ws[0] += t0; ws[1] += t1; ws[2] += t2; }
                    }
                    ps0 = ps1 = ps2 = 0.0;
                    cur_s = s_lo == s_hi ? s_lo : -1;
                }
                if (cur_s >= 0) {
                    ps0 += r0[u]; ps1 += r1[u]; ps2 += r2[u];            // rows beyond n carry zeros
                } else {
                    for (int ss = s_lo; ss <= s_hi; ++ss) {
                        const bool mine = sidx[u] == ss;
                        const double t0 = warp_sum(mine ? r0[u] : 0.0), t1 = warp_sum(mine ? r1[u] : 0.0), t2 = warp_sum(mine ? r2[u] : 0.0);
                        if (lane == 0) {
                            double *ws = wsum + (wid * YC_MAX_SAMPLES + ss) * 3;
                            ws[0] += t0; ws[1] += t1; ws[2] += t2;
                        }
                    }
                }
            }
        }
        if (cur_s >= 0) {
            const double t0 = warp_sum(ps0), t1 = warp_sum(ps1), t2 = warp_sum(ps2);
            if (lane == 0) { double *ws = wsum + (wid * YC_MAX_SAMPLES + cur_s) * 3; ws[0] += t0; ws[1] += t1; ws[2] += t2; }
        }
        acc[2 * NS] += log(seprod);
        warp_sum_array<NP>(acc);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (lane == 0) part[wid * NP + q] = acc[q];
        }
        __syncthreads();
        // ---- bookkeeping of em(): likelihood, stop rule
        double ll = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < YC_NWARPS; ++w2) ll += part[w2 * NP + 2 * NS];
        if (iteration > 0) improvement = ll - last;
        ++iteration;
        last = ll;
        if (!isfinite(ll)) { failed = true; break; }                     // assumption A5
        const bool again = improvement > 0.1 && iteration < 1001;
        if (!again) break;            // the last E-step's summaries are discarded (stats.py:224)
        // ---- M-step of the NEXT iteration (stats.py:202-211)
        const double step = 1.0 - (0.99 * pow(2.0 + (double)iteration, -1e-3));
        if (tid < 64) {
            // warp k updates component k.  The summaries S = carried + this pass are formed IN `carry` (one lane per
            // statistic); a skipped update simply leaves them there.  The full-covariance update is spread over the
            // lanes (one per matrix entry / mean coordinate: same expressions, same results as the serial form that
            // kept 126 threads at the barrier); only the W x W factorisation stays on one lane.
            const int k = wid;
            double *S = carry + k * NS;
#pragma unroll 1
            for (int q = lane; q < NS; q += 32) {
                double t = S[q];
                for (int w2 = 0; w2 < YC_NWARPS; ++w2) t += part[w2 * NP + k * NS + q];
                S[q] = t;
            }
            __syncwarp();
            const double S0 = S[0];
            if (p.covariance_type == YC_COV_DIAG) {
                // NormalDistribution.from_summaries per column (A10): skipped below 1e-8, mean AND sigma blended
                if (S0 >= 1e-8 && lane == 0) {
                    double logdet = 0.0;
#pragma unroll 1
                    for (int a = 0; a < W; ++a) {
                        const double s2 = S[1 + W + a * W - a * (a - 1) / 2];
                        const double m_new = S[1 + a] / S0;
                        const double sg_new = sqrt(s2 / S0 - (S[1 + a] * S[1 + a]) / (S0 * S0));
                        const double sg = sqrt(par->cov[k][a * W + a]) * step + sg_new * (1.0 - step);
                        par->mu[k][a] = par->mu[k][a] * step + m_new * (1.0 - step);
                        for (int c2 = 0; c2 < W; ++c2) { par->cov[k][a * W + c2] = 0.0; par->linv[k][a * W + c2] = 0.0; }
                        par->cov[k][a * W + a] = sg * sg;
                        par->linv[k][a * W + a] = 1.0 / sg;
                        logdet += 2.0 * log(sg);
                    }
                    par->cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
                    for (int q = 0; q < NS; ++q) S[q] = 0.0;
                }
            } else if (S0 >= 1e-7) {                                     // A3: below, skipped and the summaries kept
#pragma unroll 1
                for (int e = lane; e < W * W; e += 32) {
                    const int a = e / W, c2 = e - a * W;
                    const int lo = a < c2 ? a : c2, hi = a < c2 ? c2 : a;
                    const double s2 = S[1 + W + lo * W - lo * (lo - 1) / 2 + (hi - lo)];
                    const double ua = S[1 + a] / S0, uc = S[1 + c2] / S0;
                    const double cv = (s2 - S[1 + a] * uc - ua * S[1 + c2] + S0 * (ua * uc)) / S0;
                    par->cov[k][e] = par->cov[k][e] * step + cv * (1.0 - step);
                }
                if (lane < W) par->mu[k][lane] = par->mu[k][lane] * step + (S[1 + lane] / S0) * (1.0 - step);
                __syncwarp();
#pragma unroll 1
                for (int q = lane; q < NS; q += 32) S[q] = 0.0;
                if (lane == 0) {
                    double logdet;
                    if (!cholesky_inverse<W>(par->cov[k], par->linv[k], logdet)) par->ok = 0;   // A4
                    par->cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
                }
            }
        }
        if (tid >= 64 && tid < 64 + R) {
            const int s = tid - 64;
            if (c.fstart[s + 1] > c.fstart[s]) {
                double v[3] = {0.0, 0.0, 0.0};
                for (int w2 = 0; w2 < YC_NWARPS; ++w2)
                    for (int k = 0; k < 3; ++k) v[k] += wsum[(w2 * YC_MAX_SAMPLES + s) * 3 + k];
                if (v[0] + v[1] + v[2] != 0.0) {
                    const int ncomp = p.add_uniform ? 3 : 2;
                    double t = 0.0;
                    for (int k = 0; k < ncomp; ++k) { v[k] += 0.5; t += v[k]; }
                    for (int k = 0; k < 3; ++k) par->logw[s][k] = k < ncomp ? log(v[k] / t) : -INFINITY;
                }
            }
        }
        __syncthreads();
        failed = par->ok == 0;
    }
    // ---- results (fit_multisamp_gmm, stats.py:250-253)
    if (tid == 0) {
        if (failed) {
            status[g] = YC_ST_FIT_FAILED;
        } else {
            status[g] = YC_ST_TESTED;
            double cw[3] = {0.0, 0.0, 0.0}, tot = 0.0;
            for (int s = 0; s < R; ++s) {
                const double ns = (double)(c.fstart[s + 1] - c.fstart[s]);
                if (ns > 0.0) { for (int k = 0; k < 3; ++k) cw[k] += exp(par->logw[s][k]) * ns; tot += ns; }
            }
            for (int k = 0; k < 3; ++k) out.cw[(size_t)slot * 3 + k] = cw[k] / tot;
        }
        out.iters[slot] = iteration;
        out.loglik[slot] = last;
    }
    __syncthreads();
    for (int i = tid; i < 2 * W; i += YC_BLOCK) out.mu[(size_t)slot * 2 * W + i] = par->mu[i / W][i % W];
    for (int i = tid; i < 2 * W * W; i += YC_BLOCK) out.cov[(size_t)slot * 2 * W * W + i] = par->cov[i / (W * W)][i % (W * W)];
    for (int i = tid; i < R * 3; i += YC_BLOCK) out.logw[(size_t)slot * R * 3 + i] = par->logw[i / 3][i % 3];
}

// ---------------------------------------------------------------------------
// Kernel 3
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(YC_BLOCK)
gtest_smpred_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
                    const int32_t *__restrict__ fit_list, const int32_t *__restrict__ fit_count, const yc_fit_out_t out,
                    const uint8_t *__restrict__ status, double *__restrict__ post,
                    const long long post_capacity, unsigned long long *__restrict__ post_cursor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const FitSmem L = fit_smem_layout(W, R, b.max_reads, cap, 1);
    EmPar<W> *par = reinterpret_cast<EmPar<W> *>(smem_raw + L.par_off);
    long long *soff = reinterpret_cast<long long *>(smem_raw + L.sel_off);
    const int slot = blockIdx.x;
    if (fit_count != nullptr && slot >= *fit_count) return;            // the grid is an upper bound: the count lives on the device
    const int64_t g = fit_list[slot];
    if (status[g] != YC_ST_TESTED) {                                     // failed fit: no row
        if (tid == 0) out.post_off[slot] = -1;
        if (tid < 8) out.stats[(size_t)slot * 8 + tid] = nan("");
        return;
    }
    FitCtx<W> c;
    fit_front_end<W>(b, p, g, smem_raw, L, cap, false, c);      // all complete rows, no sub-sampling
    const int n_all = c.n;
    if (tid == 0) {
        // preds = np.round(weights * ORIGINAL sizes)  (stats.py:296), present samples only
        double pr[YC_MAX_SAMPLES * 3];
        int rows_c = 0, rows_t = 0, q = 0;
        for (int s = 0; s < R; ++s) {
            const int sz = c.cnt[W * R + s];
            double *po = out.preds + ((size_t)slot * R + s) * 3;
            if (sz > 0) {
                for (int k = 0; k < 3; ++k) {
                    const double v = rint(exp(out.logw[((size_t)slot * R + s) * 3 + k]) * (double)sz);
                    pr[q * 3 + k] = v; po[k] = v;
                }
                ++q;
                if (s < Rc) ++rows_c; else ++rows_t;
            } else { po[0] = po[1] = po[2] = 0.0; }
        }
        double st[8];
        gtest_fractions(pr, rows_c, rows_t, p.fdr_threshold, st, p.add_uniform ? 3 : 2);
        for (int k = 0; k < 8; ++k) out.stats[(size_t)slot * 8 + k] = st[k];
        long long off = -1;
        if (post != nullptr && st[2] < p.fdr_threshold) {
            off = (long long)atomicAdd(post_cursor, (unsigned long long)n_all);
            if (off + n_all > post_capacity) off = -2;           // caller sized the buffer too small
        }
        out.post_off[slot] = off;
        soff[0] = off;
    }
    __syncthreads();
    const long long off = soff[0];
    if (off < 0) return;
    // ---- posteriors of every complete-window read under the combined weights (stats.py:494-499)
    for (int i = tid; i < 2 * W; i += YC_BLOCK) par->mu[i / W][i % W] = out.mu[(size_t)slot * 2 * W + i];
    for (int i = tid; i < 2 * W * W; i += YC_BLOCK) par->cov[i / (W * W)][i % (W * W)] = out.cov[(size_t)slot * 2 * W * W + i];
    for (int i = tid; i < W; i += YC_BLOCK) { par->lo[i] = out.box[((size_t)slot * W + i) * 2]; par->hi[i] = out.box[((size_t)slot * W + i) * 2 + 1]; }
    __syncthreads();
    if (tid < 64 && lane == 0) {
        const int k = wid;
        double logdet;
        cholesky_inverse<W>(par->cov[k], par->linv[k], logdet);
        par->cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
    }
    if (tid == 64) {
        double s = 0.0;
        for (int j = 0; j < W; ++j) s += log(par->hi[j] - par->lo[j]);
        par->ulogp = -s;
        const double *cw = out.cw + (size_t)slot * 3;
        const double t = cw[0] + cw[1] + cw[2];                  // A7
        for (int k = 0; k < 3; ++k) par->logw[0][k] = log(cw[k] / t);
    }
    __syncthreads();
    const int nwords = (c.w.n_reads + 31) >> 5;
    for (int wd = wid; wd < nwords; wd += YC_NWARPS) {
        const uint32_t m = c.mask[wd];
        if ((m >> lane) & 1u) {
            const int r = wd * 32 + lane;
            const int rank = c.wordoff[wd] + __popc(m & ((1u << lane) - 1u));
            double x[W];
#pragma unroll
            for (int j = 0; j < W; ++j) x[j] = h2d(__ldg(c.w.col[j] + r));
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
            const double l2 = (inside && p.add_uniform) ? par->ulogp + par->logw[0][2] : -INFINITY;
            const double mx = fmax(fmax(l0, l1), l2);
            const double e0 = exp_nonpos(l0 - mx), e1 = exp_nonpos(l1 - mx), e2 = exp_nonpos(l2 - mx);
            const double inv = 1.0 / (e0 + e1 + e2);
            post[2 * (off + rank)] = e1 * inv;
            post[2 * (off + rank) + 1] = e2 * inv;
        }
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
int fit_cap(const yc_batch_t &b, const yc_params_t &p) {
    long long cap = (long long)b.n_samples * p.max_fit_depth;
    if (cap > b.max_reads) cap = b.max_reads;
    cap = (cap + 7) & ~7ll;
    return cap < 32 ? 32 : (int)cap;
}

template <int W_> struct WTag { static constexpr int value = W_; };
template <typename F>
static int dispatch_w(int W, F &&f) {
    switch (W) {
        case 1: return f(WTag<1>());
        case 3: return f(WTag<3>());
        case 5: return f(WTag<5>());
        case 7: return f(WTag<7>());
        case 9: return f(WTag<9>());
    }
    return YC_EINVAL;
}

int fit_smem_bytes(const yc_batch_t &b, const yc_params_t &p, int which) {
    // which 0 = init, 1 = em, 2 = gtest (all complete rows are never staged: cap as em)
    return fit_smem_layout(p.window_size, b.n_samples, b.max_reads, fit_cap(b, p), which == 0 ? 0 : 1).total;
}

int init_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                  const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, int8_t *labels, int max_fit_rows, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    const int cap = fit_cap(b, p);
    const int smem = fit_smem_bytes(b, p, 0);
    return dispatch_w(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        cudaError_t e = cudaFuncSetAttribute(gmm_init_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        gmm_init_kernel<W_><<<n_fit, YC_BLOCK, smem, st>>>(b, p, cap, fit_list, fit_count, out, status, labels, max_fit_rows);
        count_launches(1);
        return (int)cudaGetLastError();
    });
}

int em_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    const int cap = fit_cap(b, p);
    const int smem = fit_smem_bytes(b, p, 1);
    return dispatch_w(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        cudaError_t e = cudaFuncSetAttribute(gmm_em_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        gmm_em_kernel<W_><<<n_fit, YC_BLOCK, smem, st>>>(b, p, cap, fit_list, fit_count, out, status);
        count_launches(1);
        return (int)cudaGetLastError();
    });
}

int gtest_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                   const int32_t *fit_count, const yc_fit_out_t &out, const uint8_t *status, double *post, int64_t post_capacity,
                   int64_t *post_cursor, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    const int cap = fit_cap(b, p);
    const int smem = fit_smem_bytes(b, p, 2);
    return dispatch_w(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        cudaError_t e = cudaFuncSetAttribute(gtest_smpred_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        gtest_smpred_kernel<W_><<<n_fit, YC_BLOCK, smem, st>>>(
            b, p, cap, fit_list, fit_count, out, status, post, (long long)post_capacity,
            reinterpret_cast<unsigned long long *>(post_cursor));
        count_launches(1);
        return (int)cudaGetLastError();
    });
}

}  // namespace yc
