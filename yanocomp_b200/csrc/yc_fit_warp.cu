// Kernels 2a / 2b / 3, small-window fast path: ONE WARP (a 32-thread CTA) per
// gated position.
//
// Same arithmetic as gmm_init_kernel / gmm_em_kernel / gtest_smpred_kernel in
// yc_fit.cu (see there for the reference citations), reorganised after the
// round-1 profiles (profiles/r01_summary.md):
//  * with ~200 reads per position a 128-thread CTA spends its time in block
//    barriers around 2-3 busy threads (M-step, Cholesky); a warp needs only
//    `__syncwarp`, and because every position is its own CTA the hardware
//    scheduler balances the very uneven EM lengths (2..1001 E-steps);
//  * a first fused init+EM+G-test warp kernel was 240 KB of SASS and lost 55 %
//    of its issue slots to instruction-cache misses ("no_instructions"
//    stalls): the three stages are separate kernels again, serial code is
//    __noinline__ with rolled loops, and wide reductions go through shared
//    memory instead of unrolled shuffle trees, so that each kernel's hot loop
//    stays inside the 32 KB instruction cache.
// Eligibility (dispatcher): w <= 5, <= 1024 reads per feature, no sub-sampling
// possible (reads per condition <= max_fit_depth).
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "yc_fit_common.cuh"
#include "yc_internal.h"
#include "yc_stats.cuh"

namespace yc {

// Scheduling (round 2): PERSISTENT warps.  A kernel is launched with as many one-warp CTAs as the device holds at
// once, and every warp fetches the next unprocessed slot of fit_list from a device counter until the list is
// exhausted.  Compared with one CTA per position this pays no CTA launch per position, still balances the very
// uneven EM lengths dynamically, and lets the number of gated positions stay ON THE DEVICE (*fit_count, written by
// yc_ks_batch): the host never waits for it.
// The CTA stays ONE warp, and the fetched slot travels through shared memory, on purpose: the compiler's divergence
// analysis treats an atomic's result (and anything shuffled from it, and threadIdx.x >> 5) as thread-varying, and
// then guards every __shfl_sync / __ballot_sync of the per-position program with a BRA.DIV slow path (a four-warp
// variant of the init kernel grew from 3 352 to 4 560 instructions and ran 11 % slower).  A value loaded from a
// uniform shared-memory address is provably warp-uniform, so the per-position code compiles as it did when the
// slot was blockIdx.x.
#define YC_FW_WARPS 1

__device__ __forceinline__ int fw_next_slot(int *work, int lane) {
    __shared__ int s_slot;
    if (lane == 0) s_slot = atomicAdd(work, 1);
    __syncwarp();
    const int s = s_slot;
    __syncwarp();
    return s;
}
__device__ __forceinline__ int fw_slot_count(int n_fit, const int32_t *fit_count) {
    if (fit_count == nullptr) return n_fit;
    const int c = *fit_count;
    return c < n_fit ? c : n_fit;
}

// EM: row stride of the main reduction rows.  The 2 NS + 1 per-lane partial sums are first folded across lane pairs
// (one shuffle step in registers), so a row holds 16 partials + the total: 17 doubles instead of 33.  That takes the
// EM kernel's shared memory from 10.5 to 8.1 KB per position = 24 instead of 20 resident warps per SM.
constexpr int FW_EM_RS = 17;

struct FwSmem {
    int xs_off, srow_off, lab_off, buf_off, par_off, mask_off, cnt_off, fstart_off, wtmp_off, carry_off, red_off, total;
};

// which: 0 = init (sort buffer, labels, 2 NS reduction rows), 1 = EM (responsibilities, 2 NS + 1 + 2 R
// reduction rows), 2 = G-test / posteriors (rows and the parameter block only)
__host__ __device__ inline FwSmem fw_layout(int W, int R, int cap, int which) {
    FwSmem s;
    int P = 32;
    while (P < cap) P <<= 1;
    const int NS = 1 + W + W * (W + 1) / 2;
    // EM: the responsibilities (2 cap doubles, dead once the statistics pass has read them) and the main
    // reduction rows (2 NS + 1 rows of 33, written after that pass) share one region; only the per-sample
    // sums, flushed DURING the pass, get rows of their own
    const int em_main = (2 * NS + 1) * FW_EM_RS;
    const int buf_n = which == 0 ? P : (which == 1 ? (2 * cap > em_main ? 2 * cap : em_main) : 0);
    const int red_rows = which == 0 ? 2 * NS : (which == 1 ? 2 * R : 3);
    int o = 0;
    s.buf_off = o;    o += buf_n * (int)sizeof(double);                         // MAD sort buffer / EM responsibilities
    s.par_off = o;    o += (4 * W * W + 4 * W + 3 * YC_MAX_SAMPLES + 8) * (int)sizeof(double);
    s.wtmp_off = o;   o += 2 * YC_MAX_SAMPLES * (int)sizeof(double);
    s.carry_off = o;  o += (which == 1 ? 2 * NS : 0) * (int)sizeof(double);          // EM: summaries kept when S0 < 1e-7
    s.red_off = o;    o += red_rows * 33 * (int)sizeof(double);
    s.mask_off = o;   o += ((cap + 31) / 32 + 1) * (int)sizeof(uint32_t);
    s.cnt_off = o;    o += (YC_MAX_SAMPLES + 2) * (int)sizeof(int);
    s.fstart_off = o; o += (YC_MAX_SAMPLES + 2) * (int)sizeof(int);
    s.xs_off = o;     o += W * cap * (int)sizeof(uint16_t);
    o = (o + 3) & ~3;
    s.srow_off = o;   o += cap;
    s.lab_off = o;    o += which == 0 ? cap : 0;
    s.total = (o + 15) & ~15;
    return s;
}

template <int W>
struct FwCtx {
    EmPar<W> *par;
    double *buf, *red, *wtmp;
    uint32_t *mask;
    int *fstart;
    uint16_t *xs;
    uint8_t *srow;
    int8_t *lab;
    int n, cap;
    uint64_t pkey;
};

template <int W>
__device__ __forceinline__ void fw_row(const uint16_t *xs, int cap, int i, double (&x)[W]) {
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = h2d(xs[j * cap + i]);
}

// Sum K per-lane values over the warp through shared memory (row stride 33:
// conflict-free); every lane gets the totals.  Deterministic; ~K stores +
// 32 loads + K broadcast loads instead of 10*K shuffles, and compact code.
template <int K>
__device__ __forceinline__ void warp_sum_smem(double (&v)[K], double *red, int lane) {
    // G lanes share one total: each adds 32 / G of the per-lane partials, then log2 G shuffle steps
    constexpr int KP = K <= 1 ? 1 : (K <= 2 ? 2 : (K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 16 ? 16 : 32))));
    constexpr int G = 32 / KP, CH = 32 / G;
#pragma unroll
    for (int q = 0; q < K; ++q) red[q * 33 + lane] = v[q];
    __syncwarp();
    if (K <= 16) {
        const int q = lane / G, part = lane % G;
        double t = 0.0;
        if (q < K) {
            const double *rq = red + q * 33 + part * CH;
#pragma unroll
            for (int i = 0; i < CH; ++i) t += rq[i];
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(YC_FULL, t, o);
        if (q < K && part == 0) red[q * 33 + 32] = t;
    } else {
#pragma unroll 1
        for (int q = lane; q < K; q += 32) {
            const double *rq = red + q * 33;
            double ta = 0.0, tb = 0.0, tc = 0.0, td = 0.0;             // four interleaved chains: latency / 4
#pragma unroll
            for (int l2 = 0; l2 < 32; l2 += 4) { ta += rq[l2]; tb += rq[l2 + 1]; tc += rq[l2 + 2]; td += rq[l2 + 3]; }
            red[q * 33 + 32] = (ta + tb) + (tc + td);
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < K; ++q) v[q] = red[q * 33 + 32];
    __syncwarp();
}

// ascending sort of buf[0..P), P a power of two >= 32: in registers (P / 32 values per lane) up to
// P = 512, in register chunks of 512 merged through shared memory beyond (warp_sort_long, yc_common.cuh)
template <int E>
__device__ __forceinline__ void warp_sort_buf_regs(double *buf, int lane) {
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = buf[e * 32 + lane];              // any assignment will do before the sort
    __syncwarp();
    warp_sort_regs<E>(v, lane, 32);
#pragma unroll
    for (int e = 0; e < E; ++e) buf[lane * E + e] = v[e];
    __syncwarp();
}
// complete-window rows of position g -> shared memory (binary16, pooled order), per-sample offsets
template <int W>
__device__ __forceinline__ bool fw_front(const yc_batch_t &b, const yc_params_t &p, int64_t g,
                                      unsigned char *smem, const FwSmem &L, int cap, int lane, bool gather,
                                      FwCtx<W> &c) {
    const int R = b.n_samples;
    c.buf = reinterpret_cast<double *>(smem + L.buf_off);
    c.par = reinterpret_cast<EmPar<W> *>(smem + L.par_off);
    c.red = reinterpret_cast<double *>(smem + L.red_off);
    c.wtmp = reinterpret_cast<double *>(smem + L.wtmp_off);
    c.mask = reinterpret_cast<uint32_t *>(smem + L.mask_off);
    int *cnt = reinterpret_cast<int *>(smem + L.cnt_off);
    c.fstart = reinterpret_cast<int *>(smem + L.fstart_off);
    c.xs = reinterpret_cast<uint16_t *>(smem + L.xs_off);
    c.srow = reinterpret_cast<uint8_t *>(smem + L.srow_off);
    c.lab = reinterpret_cast<int8_t *>(smem + L.lab_off);
    c.cap = cap;
    c.n = 0;
    Window<W> w;
    if (!window_setup<W>(b, p, g, w)) return false;
    c.pkey = w.pkey;
    if (lane <= R) cnt[lane] = 0;
    __syncwarp();
    const int nr = w.n_reads;
    // One pass over the rows in blocks of FW_DEPTH x 32: all loads of a block are issued before the first
    // ballot consumes one, so a position pays one memory latency per 256 rows instead of one per 32 rows and
    // pass (the profile had 8-12 % of the EM kernel's time on these loads), and the values go straight from
    // registers into shared memory: no second pass to gather them.
    constexpr int FW_DEPTH = 8;
    const uint32_t lt = (1u << lane) - 1u;
    int run = 0;
#pragma unroll 1
    for (int blk = 0; blk < nr; blk += 32 * FW_DEPTH) {
        uint16_t v[FW_DEPTH][W];
#pragma unroll
        for (int k = 0; k < FW_DEPTH; ++k) {
            const int r = blk + 32 * k + lane;
#pragma unroll
            for (int j = 0; j < W; ++j) v[k][j] = r < nr ? __ldg(w.col[j] + r) : (uint16_t)0x7e00;
        }
#pragma unroll
        for (int k = 0; k < FW_DEPTH; ++k) {
            const int base = blk + 32 * k;
            if (base < nr) {                                // warp-uniform
                const int r = base + lane;
                uint32_t m_all = YC_FULL;
#pragma unroll
                for (int j = 0; j < W; ++j) m_all &= __ballot_sync(YC_FULL, !h_isnan(v[k][j]));
                if (lane == 0) c.mask[base >> 5] = m_all;
                if (lane < R) {                             // lane s counts the complete rows of sample s
                    const int a = max(w.samp_start[lane], base), e = min(w.samp_start[lane + 1], base + 32);
                    if (e > a) {
                        const uint32_t seg = (e - a == 32) ? YC_FULL : (((1u << (e - a)) - 1u) << (a - base));
                        cnt[lane] += __popc(m_all & seg);
                    }
                }
                if (gather && ((m_all >> lane) & 1u)) {
                    const int rank = run + __popc(m_all & lt);       // <= complete rows <= n_reads <= cap
#pragma unroll
                    for (int j = 0; j < W; ++j) c.xs[j * cap + rank] = v[k][j];
                    int s2 = 0;
                    while (s2 + 1 < R && r >= w.samp_start[s2 + 1]) ++s2;
                    c.srow[rank] = (uint8_t)s2;
                }
                run += __popc(m_all);
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        int acc = 0;
#pragma unroll 1
        for (int s = 0; s < R; ++s) { c.fstart[s] = acc; acc += cnt[s]; }
        c.fstart[R] = acc;
    }
    __syncwarp();
    c.n = c.fstart[R];
    if (c.n < 1 || c.n > cap) return false;
    return true;
}

// ===========================================================================
// Kernel 2a (warp): k-means++ x3, MAD outliers, ML Gaussians, weights, box
// ===========================================================================
// unroll factor of the k-means row loops (seeding passes, Lloyd pass) of init_warp_kernel
#ifndef YC_FW_LLOYD_UNROLL
#define YC_FW_LLOYD_UNROLL 1
#endif
#ifndef YC_FW_INIT_MINBLOCKS
#define YC_FW_INIT_MINBLOCKS 20
#endif
// SE = slots per lane of the MAD sort (the batch's row capacity rounded up to a power of two, / 32), 0 = more
// than 512 rows (warp_sort_long).  A template parameter because merely carrying the other sort variants' code
// made the kernel slower (3.49 -> 3.63 ms on cfg2 with the long sort's 2 100 unexecuted instructions).
template <int W, int SE>
__device__ __forceinline__ void
init_warp_position(const yc_batch_t &b, const yc_params_t &p, const int cap, const int slot,
                   const int32_t *__restrict__ fit_list, const yc_fit_out_t &out,
                   uint8_t *__restrict__ status, int8_t *__restrict__ labels_out, const int max_fit_rows,
                   unsigned char *smem_raw, const FwSmem &L, const int lane) {
    const int R = b.n_samples;
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT;
    const int64_t g = fit_list[slot];
    const bool add_uniform = p.add_uniform != 0, want_labels = labels_out != nullptr;
    FwCtx<W> c;
    if (!fw_front<W>(b, p, g, smem_raw, L, cap, lane, true, c)) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }
    const int n = c.n;
    const uint16_t *xs = c.xs;
    double *buf = c.buf, *red = c.red;
    int8_t *lab = c.lab;
    const uint64_t pkey = c.pkey;

    // column totals of the fit rows (exact), so that Lloyd only accumulates one cluster
    double sumx[W];
#pragma unroll
    for (int j = 0; j < W; ++j) sumx[j] = 0.0;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(xs, cap, i, x);
#pragma unroll
        for (int j = 0; j < W; ++j) sumx[j] += x[j];
    }
#pragma unroll
    for (int j = 0; j < W; ++j) sumx[j] = warp_sum(sumx[j]);

    // ---- k-means++ x3 + Lloyd  (assumption A2)
    double best[2 * W], best_total = INFINITY;
    bool have_best = false;
    int C = (n + 31) / 32;
    if ((C & 7) == 0) ++C;                               // a chunk length that is a multiple of 8 rows = 4+-way bank conflicts on the binary16 rows
    const int i0c = min(lane * C, n), i1c = min(i0c + C, n);
#pragma unroll 1
    for (int restart = 0; restart < 3; ++restart) {
        double cen[2 * W];
        {
            const double u0 = rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 0);
            double x[W];
            fw_row<W>(xs, cap, min((int)(u0 * (double)n), n - 1), x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[j] = x[j];
        }
        {   // second centre: D^2 sampling through the (exact) cumulative sum
            double loc = 0.0;
YC_UNROLL(YC_FW_LLOYD_UNROLL)
            for (int i = i0c; i < i1c; ++i) { double x[W]; fw_row<W>(xs, cap, i, x); loc = add_rn(loc, sqdist<W>(x, cen)); }
            double inc = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(YC_FULL, inc, o); if (lane >= o) inc = add_rn(inc, t); }
            const double total = __shfl_sync(YC_FULL, inc, 31);
            double run = sub_rn(inc, loc);
            const double target = mul_rn(rng_uniform(pkey, YC_STREAM_KMEANS0 + restart, 1), total);
            int below = 0;
YC_UNROLL(YC_FW_LLOYD_UNROLL)
            for (int i = i0c; i < i1c; ++i) { double x[W]; fw_row<W>(xs, cap, i, x); run = add_rn(run, sqdist<W>(x, cen)); below += (run <= target); }
            below = warp_sum_i(below);
            double x[W];
            fw_row<W>(xs, cap, min(below, n - 1), x);
#pragma unroll
            for (int j = 0; j < W; ++j) cen[W + j] = x[j];
        }
        int iteration = 0;
        double improvement = INFINITY, last = INFINITY, total = INFINITY;
        double acc[W + 2];                                   // column sums of cluster 1, total distance, size of cluster 1
        bool have_sums = false;
        double *cs = red + (W + 2) * 33;                     // the 2W centroid coordinates, kept in shared memory
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 2 * W; ++q) cs[q] = cen[q];
        }
        __syncwarp();
#pragma unroll 1
        while (improvement > 0.5 && iteration < 101) {
            if (have_sums) {
                // centroid = cluster sum / cluster size from the totals of the previous pass (still in `red`).
                // Cluster 0's sums are the column totals minus cluster 1's (sums of binary16 values: exact either
                // way).  Lane q divides coordinate q (one IEEE division per lane instead of 2W per lane) and stores
                // it; an empty cluster keeps its centroid (A2).
                if (lane < 2 * W) {
                    const int j = lane < W ? lane : lane - W;
                    const double s1 = red[j * 33 + 32], n1d = red[(W + 1) * 33 + 32];
                    double sj = sumx[0];
#pragma unroll
                    for (int jj = 1; jj < W; ++jj) if (j == jj) sj = sumx[jj];
                    const double den = lane < W ? (double)n - n1d : n1d;
                    const double num = lane < W ? sub_rn(sj, s1) : s1;
                    if (den > 0.0) cs[lane] = div_rn(num, den);
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 2 * W; ++q) cen[q] = cs[q];
            }
#pragma unroll
            for (int k = 0; k < W + 2; ++k) acc[k] = 0.0;
YC_UNROLL(YC_FW_LLOYD_UNROLL)
            for (int i = lane; i < n; i += 32) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
                const double d0 = sqdist<W>(x, cen), d1 = sqdist<W>(x, cen + W);
                const bool k = d1 < d0;
                // predicated adds (the select form cost two FSEL per double on top of the DADD)
                if (k) {
                    acc[W] += d1;
                    acc[W + 1] += 1.0;
#pragma unroll
                    for (int j = 0; j < W; ++j) acc[j] += x[j];
                } else {
                    acc[W] += d0;
                }
            }
            warp_sum_smem<W + 2>(acc, red, lane);
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
    if (!have_best) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }

    // ---- nearest-centroid labels; MAD outliers unless --no-uniform (stats.py:114-122, A1)
    int n_out = 0;
    if (!add_uniform) {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
            double x[W];
            fw_row<W>(xs, cap, i, x);
            lab[i] = (int8_t)(sqdist<W>(x, best + W) < sqdist<W>(x, best));
        }
        __syncwarp();
    } else {
    int P = 32 * SE;                                   // fixed by the batch (padding = +inf)
    if (SE == 0) { P = 1024; while (P < cap) P <<= 1; }
#pragma unroll 1
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
    if (SE > 0) warp_sort_buf_regs<(SE > 0 ? SE : 1)>(buf, lane); else warp_sort_long(buf, P, lane);
    const double med = (n & 1) ? buf[n >> 1] : mul_rn(add_rn(buf[(n >> 1) - 1], buf[n >> 1]), 0.5);
    // MAD = median of |d - med| WITHOUT a second sort: in the sorted distances the deviations are
    // V-shaped, i.e. two ascending runs (left of the median walked outwards, right of it), so the
    // order statistics of their union follow from merge ranks found by binary search.
    double *wsel = c.wtmp;
    const int m_le = upper_bound(buf, n, med);            // elements <= med form the left run
    const int k_hi = n >> 1, k_lo = (n & 1) ? k_hi : k_hi - 1;
#pragma unroll 1
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
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(xs, cap, i, x);
        const double d0 = sqdist<W>(x, best), d1 = sqdist<W>(x, best + W);
        if ((d1 < d0 ? d1 : d0) > thr) { lab[i] = 2; ++n_out; }
    }
    n_out = warp_sum_i(n_out);
    __syncwarp();
    }
    if (want_labels)
        for (int i = lane; i < n && i < max_fit_rows; i += 32) labels_out[(size_t)slot * max_fit_rows + i] = lab[i];

    // ---- ML Gaussians per label from exact sufficient statistics (stats.py:148-152)
    bool ok_all = true;
    {
        double S[2 * NS];
#pragma unroll
        for (int q = 0; q < 2 * NS; ++q) S[q] = 0.0;
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
            const int l = lab[i];
            if (l < 2) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
                const double f0 = l == 0 ? 1.0 : 0.0, f1 = 1.0 - f0;
                S[0] += f0; S[NS] += f1;
                int q = 1 + W;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    S[1 + a] += f0 * x[a]; S[NS + 1 + a] += f1 * x[a];
#pragma unroll
                    for (int c2 = a; c2 < W; ++c2) { const double pr = x[a] * x[c2]; S[q] += f0 * pr; S[NS + q] += f1 * pr; ++q; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2 * NS; ++q) red[q * 33 + lane] = S[q];
        __syncwarp();
        double tot_q[2];                                    // lane q sums rows q and q+32
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = lane + 32 * h;
            double t = 0.0;
            if (q < 2 * NS) {
#pragma unroll 8
                for (int l2 = 0; l2 < 32; ++l2) t += red[q * 33 + l2];
            }
            tot_q[h] = t;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) if (lane + 32 * h < 2 * NS) red[lane + 32 * h] = tot_q[h];   // compact totals
        __syncwarp();
        bool ok = true;
        if (lane < 2) {
            double mu[W], cov[W * W];
            ok = ml_gaussian<W>(red + lane * NS, mu, cov);
#pragma unroll 1
            for (int a = 0; a < W; ++a) out.mu[((size_t)slot * 2 + lane) * W + a] = mu[a];
#pragma unroll 1
            for (int a = 0; a < W * W; ++a) out.cov[((size_t)slot * 2 + lane) * W * W + a] = cov[a];
        }
        ok_all = __all_sync(YC_FULL, ok);
    }
    // ---- uniform box (stats.py:178-187): outlier rows, or all rows if none
    {
        double lo[W], hi[W];
#pragma unroll
        for (int j = 0; j < W; ++j) { lo[j] = INFINITY; hi[j] = -INFINITY; }
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
            if (n_out == 0 || lab[i] == 2) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
#pragma unroll
                for (int j = 0; j < W; ++j) { lo[j] = min_nn(lo[j], x[j]); hi[j] = max_nn(hi[j], x[j]); }
            }
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
            lo[j] = warp_min(lo[j]); hi[j] = warp_max(hi[j]);
            if (add_uniform) ok_all &= hi[j] > lo[j];
            if (lane == 0) { out.box[((size_t)slot * W + j) * 2] = lo[j]; out.box[((size_t)slot * W + j) * 2 + 1] = hi[j]; }
        }
    }
    // ---- per-sample initial weights (stats.py:173-177): lane k writes component k
#pragma unroll 1
    for (int s = 0; s < R; ++s) {
        const int a = c.fstart[s], e = c.fstart[s + 1];
        int k0 = 0, k1 = 0;
#pragma unroll 1
        for (int i = a + lane; i < e; i += 32) { const int l = lab[i]; k0 += l == 0; k1 += l == 1; }
        k0 = warp_sum_i(k0); k1 = warp_sum_i(k1);
        if (lane < 3) {
            double v = 0.0;
            if (e > a) {
                const int kk = lane == 0 ? k0 : (lane == 1 ? k1 : (e - a) - k0 - k1);
                v = log(((double)kk + 0.5) / ((double)(e - a) + (add_uniform ? 1.5 : 1.0)));
                if (lane == 2 && !add_uniform) v = -INFINITY;       // two-component model
            }
            out.logw[((size_t)slot * R + s) * 3 + lane] = v;
        }
    }
    if (lane == 0 && !ok_all) status[g] = YC_ST_FIT_FAILED;
}

template <int W, int SE>
__global__ void __launch_bounds__(32 * YC_FW_WARPS, YC_FW_INIT_MINBLOCKS / YC_FW_WARPS)
init_warp_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
                 const int32_t *__restrict__ fit_list, const int n_fit, const int32_t *__restrict__ fit_count,
                 int *__restrict__ work, const yc_fit_out_t out,
                 uint8_t *__restrict__ status, int8_t *__restrict__ labels_out, const int max_fit_rows) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int lane = threadIdx.x;
    const FwSmem L = fw_layout(W, b.n_samples, cap, 0);
    unsigned char *smem_raw = smem_all;
    const int n = fw_slot_count(n_fit, fit_count);
#pragma unroll 1
    for (int slot = fw_next_slot(work, lane); slot < n; slot = fw_next_slot(work, lane)) {
        init_warp_position<W, SE>(b, p, cap, slot, fit_list, out, status, labels_out, max_fit_rows, smem_raw, L, lane);
        __syncwarp();
    }
}

// ===========================================================================
// Kernel 2b (warp): EM
// ===========================================================================
// inertia of the M-step before E-step `iteration` (stats.py:202): 1 - 0.99 (2 + iteration)^-0.001,
// tabulated on the host (the same libm pow the reference's Python float power calls) instead of
// a pow() per position and iteration on two lanes
__constant__ double c_em_step[1008];

static int upload_em_steps() {
    static std::atomic<bool> done[64];                                   // zero-initialised; callers may be on several threads
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 64 && done[dev].load(std::memory_order_acquire)) return 0;
    double h[1008];
    for (int it = 0; it < 1008; ++it) h[it] = 1.0 - (0.99 * pow(2.0 + (double)it, -1e-3));
    e = cudaMemcpyToSymbol(c_em_step, h, sizeof(h));
    if (e != cudaSuccess) return (int)e;
    if (dev < 64) done[dev].store(true, std::memory_order_release);
    return 0;
}

// M-step of the two shared Gaussians (reference stats.py:202-205 via pomegranate from_summaries, assumptions
// A3/A4), spread over the warp (round-2 profile: the M-step ran on 2 of 32 lanes and was a fifth of the kernel's
// issue slots):
//   em_blend   one lane per (component, statistic): totals + carried summaries; one lane per (component, mean
//              coordinate) and per (component, covariance entry) for the ML estimate and the inertia blend;
//   em_factor  one lane per component for what is serial: Cholesky, triangular inverse, product of the pivots.
// The logarithm of the determinant is NOT taken here: the caller takes one log for the two determinants and the
// 3 R sample weights together (a single log sequence per iteration instead of two divergent ones).
// Unlike the initialisation (cholesky_inverse, bit-exact with the oracle) EM is compared with a tolerance.
template <int W>
__device__ __forceinline__ void em_blend(EmPar<W> *par, const double *red, double *carry, double *ubuf, double step,
                                      int lane, bool &skip0, bool &skip1) {
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT;
#pragma unroll 1
    for (int idx = lane; idx < 2 * NS; idx += 32) carry[idx] += red[idx * FW_EM_RS + 16];
    __syncwarp();
    skip0 = carry[0] < 1e-7;                                             // skipped, summaries kept (A3)
    skip1 = carry[NS] < 1e-7;
    const double keep = 1.0 - step;
    if (lane < 2 * W) {
        const int k = lane >= W, a = lane - k * W;
        if (!(k ? skip1 : skip0)) {
            const double u = carry[k * NS + 1 + a] * (1.0 / carry[k * NS]);
            ubuf[lane] = u;
            par->mu[k][a] = par->mu[k][a] * step + u * keep;
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int idx = lane; idx < 2 * NT; idx += 32) {
        const int k = idx >= NT, q = idx - k * NT;
        if (!(k ? skip1 : skip0)) {
            int a = 0, rem = q;                                          // (a, c) of upper-triangle entry q
#pragma unroll
            for (int t = 0; t < W - 1; ++t) if (rem >= W - a) { rem -= W - a; ++a; }
            const int c2 = a + rem;
            const double *S = carry + k * NS, *u = ubuf + k * W;
            const double inv0 = 1.0 / S[0];
            const double cv = (S[1 + W + q] - S[1 + a] * u[c2] - u[a] * S[1 + c2] + S[0] * (u[a] * u[c2])) * inv0;
            const double nv = par->cov[k][a * W + c2] * step + cv * keep;
            par->cov[k][a * W + c2] = nv; par->cov[k][c2 * W + a] = nv;
        }
    }
    __syncwarp();
}

// Cholesky A = L L^T of par->cov[k] (L lower, packed by rows), L^-1 -> par->linv[k]; returns the product of the
// pivots (= |A|) and in `sumlog` the sum of their logs when the product leaves the double range; par->ok = 0 on a
// non-positive / non-finite pivot (A4).  One lane per component: everything in registers, reciprocal square roots
// instead of square roots and divisions.
template <int W>
__device__ __forceinline__ double em_factor(EmPar<W> *par, int k, double &sumlog) {
    constexpr int NT = W * (W + 1) / 2;
    double A[NT];
    {
        int q = 0;
#pragma unroll
        for (int a = 0; a < W; ++a)
#pragma unroll
            for (int c2 = a; c2 < W; ++c2) A[q++] = par->cov[k][a * W + c2];
    }
    double Lp[NT], Li[NT], rinv[W], prod = 1.0;
    sumlog = 0.0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            // A(i,j) with j <= i lives in the upper triangle at (j,i)
            double sacc = A[j * W - j * (j - 1) / 2 + (i - j)];
#pragma unroll
            for (int t = 0; t < j; ++t) sacc = fma(-Lp[i * (i + 1) / 2 + t], Lp[j * (j + 1) / 2 + t], sacc);
            if (i == j) {
                ok &= (sacc > 0.0) && isfinite(sacc);
                prod *= sacc;
                rinv[i] = rsqrt(sacc);                                   // one reciprocal square root instead of sqrt + division
                Lp[i * (i + 1) / 2 + i] = sacc * rinv[i];
            } else {
                Lp[i * (i + 1) / 2 + j] = sacc * rinv[j];
            }
        }
    }
    if (!ok) { par->ok = 0; return 1.0; }                                // A4
    if (!(prod > 1e-290 && prod < 1e290)) {                              // product out of range: sum of logs
        prod = 1.0;
#pragma unroll
        for (int i = 0; i < W; ++i) sumlog += 2.0 * log(Lp[i * (i + 1) / 2 + i]);
    }
#pragma unroll
    for (int c2 = 0; c2 < W; ++c2) {
        Li[c2 * (c2 + 1) / 2 + c2] = rinv[c2];
#pragma unroll
        for (int r = c2 + 1; r < W; ++r) {
            double sacc = 0.0;
#pragma unroll
            for (int t = c2; t < r; ++t) sacc = fma(-Lp[r * (r + 1) / 2 + t], Li[t * (t + 1) / 2 + c2], sacc);
            Li[r * (r + 1) / 2 + c2] = sacc * rinv[r];
        }
    }
#pragma unroll
    for (int r = 0; r < W; ++r)
#pragma unroll
        for (int c2 = 0; c2 <= r; ++c2) par->linv[k][r * W + c2] = Li[r * (r + 1) / 2 + c2];
    return prod;
}

// model of slot -> shared memory parameter block (+ Cholesky factors, uniform log density)
template <int W>
__device__ __forceinline__ void load_model(EmPar<W> *par, const yc_fit_out_t &out, int slot, int lane,
                                        const double *logw_src, int n_logw) {
    for (int i = lane; i < 2 * W; i += 32) par->mu[i / W][i % W] = out.mu[(size_t)slot * 2 * W + i];
    for (int i = lane; i < 2 * W * W; i += 32) par->cov[i / (W * W)][i % (W * W)] = out.cov[(size_t)slot * 2 * W * W + i];
    for (int i = lane; i < W; i += 32) { par->lo[i] = out.box[((size_t)slot * W + i) * 2]; par->hi[i] = out.box[((size_t)slot * W + i) * 2 + 1]; }
    for (int i = lane; i < n_logw; i += 32) par->logw[i / 3][i % 3] = logw_src[i];
    if (lane == 0) par->ok = 1;
    __syncwarp();
    if (lane < 2) {
        double logdet;
        if (!cholesky_inverse<W>(par->cov[lane], par->linv[lane], logdet)) par->ok = 0;
        par->cst[lane] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
    }
    if (lane == 2) {
        double s = 0.0;
#pragma unroll 1
        for (int j = 0; j < W; ++j) s += log(par->hi[j] - par->lo[j]);
        par->ulogp = -s;
    }
    __syncwarp();
}

#ifndef YC_EM_UNROLL1
#define YC_EM_UNROLL1 1
#endif
#ifndef YC_EM_UNROLL2
#define YC_EM_UNROLL2 1
#endif
constexpr int kEmUnroll1 = YC_EM_UNROLL1, kEmUnroll2 = YC_EM_UNROLL2;
#ifndef YC_FW_MINBLOCKS
#define YC_FW_MINBLOCKS 24
#endif

template <int W>
__device__ __forceinline__ void
em_warp_position(const yc_batch_t &b, const yc_params_t &p, const int cap, const int slot,
                 const int32_t *__restrict__ fit_list, const yc_fit_out_t &out, uint8_t *__restrict__ status,
                 unsigned char *smem_raw, const FwSmem &L, const int lane) {
    const int R = b.n_samples;
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT, NP = 2 * NS + 1;
    const int64_t g = fit_list[slot];
    if (status[g] != YC_ST_GATED) return;                               // the initialisation failed
    FwCtx<W> c;
    if (!fw_front<W>(b, p, g, smem_raw, L, cap, lane, true, c)) { if (lane == 0) status[g] = YC_ST_FIT_FAILED; return; }
    const int n = c.n;
    const uint16_t *xs = c.xs;
    uint8_t *srow = c.srow;
    EmPar<W> *par = c.par;
    double *red = c.buf, *ssum = c.red;                 // main reduction rows alias the responsibilities (fw_layout)
    double *rb0 = c.buf, *rb1 = c.buf + cap;            // responsibilities of components 0 / 1 per fit row
    const int *fstart = c.fstart;
    load_model<W>(par, out, slot, lane, out.logw + (size_t)slot * R * 3, R * 3);
    bool failed = par->ok == 0;
    const bool uni = p.add_uniform != 0;

    double *carry = reinterpret_cast<double *>(smem_raw + L.carry_off);   // [2][NS]: summaries kept when S0 < 1e-7 (A3)
    for (int q = lane; q < 2 * NS; q += 32) carry[q] = 0.0;
    // The uniform box does not move during EM (A5), so whether a row lies inside it (bounds inclusive) is decided
    // once and kept in bit 7 of the row's sample byte instead of 2 W comparisons per row and E-step.
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(xs, cap, i, x);
        bool in = true;
#pragma unroll
        for (int a = 0; a < W; ++a) in &= (x[a] >= par->lo[a]) && (x[a] <= par->hi[a]);
        if (in) srow[i] |= 0x80;
    }
    __syncwarp();
    int iteration = 0;
    double improvement = INFINITY, last = 0.0;
#pragma unroll 1
    while (!failed && improvement > 0.1 && iteration < 1001) {
        // ---- E-step (GeneralMixtureModel.summarize, A6) in two passes over the rows so that neither
        //      needs both the model (24 doubles) and the 2 NS + 1 accumulators in registers at once:
        //      pass 1 writes the responsibilities and the log-likelihood, pass 2 the sufficient statistics.
        // log-likelihood = sum_i (m_i + log se_i): the logs are taken once per lane from the running
        // product of the se_i (1 <= se_i <= 3 and at most cap/32 <= 32 rows per lane: no overflow)
        double llsum = 0.0, seprod = 1.0;
        {
            double mu0[W], mu1[W], li0[NT], li1[NT];
#pragma unroll
            for (int j = 0; j < W; ++j) { mu0[j] = par->mu[0][j]; mu1[j] = par->mu[1][j]; }
            {
                int q = 0;
#pragma unroll
                for (int a = 0; a < W; ++a)
#pragma unroll
                    for (int bb = 0; bb <= a; ++bb) { li0[q] = par->linv[0][a * W + bb]; li1[q] = par->linv[1][a * W + bb]; ++q; }
            }
            const double c0 = par->cst[0], c1 = par->cst[1], ul = par->ulogp;
#pragma unroll kEmUnroll1
            for (int i = lane; i < n; i += 32) {
                double x[W];
                fw_row<W>(xs, cap, i, x);
                const int sv = srow[i], s = sv & 0x7f;
                const bool inside = (sv & 0x80) != 0;
                double q0 = 0.0, q1 = 0.0;
                int q = 0;
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    double y0 = 0.0, y1 = 0.0;
#pragma unroll
                    for (int bb = 0; bb <= a; ++bb) { y0 = fma(li0[q], x[bb] - mu0[bb], y0); y1 = fma(li1[q], x[bb] - mu1[bb], y1); ++q; }
                    q0 = fma(y0, y0, q0); q1 = fma(y1, y1, q1);
                }
                const double l0 = c0 - 0.5 * q0 + par->logw[s][0];
                const double l1 = c1 - 0.5 * q1 + par->logw[s][1];
                const double l2 = (inside && uni) ? ul + par->logw[s][2] : -INFINITY;
                // exp(l_max - l_max) = 1 exactly: only the two non-maximal terms need an exp
                // arg-max by three comparisons (fmax is a NaN-aware multi-instruction sequence on doubles); a NaN
                // makes every comparison false, m = l2, and the NaN still reaches the likelihood through ea / eb
                const bool m0 = l0 >= l1 && l0 >= l2, m1 = !m0 && l1 >= l2;
                const double m = m0 ? l0 : (m1 ? l1 : l2);
                double ea, eb;
                exp_nonpos2((m0 ? l1 : l0) - m, ((m0 || m1) ? l2 : l1) - m, ea, eb);
                const double e0 = m0 ? 1.0 : ea, e1 = m0 ? ea : (m1 ? 1.0 : eb), e2 = (m0 || m1) ? eb : 1.0;
                const double se = e0 + e1 + e2, inv = rcp_1to3(se);
                rb0[i] = e0 * inv; rb1[i] = e1 * inv;
                llsum += m; seprod *= se;
            }
        }
        llsum += log(seprod);
        // ---- bookkeeping of em() BEFORE the statistics pass: the summaries of the last E-step are discarded
        //      (stats.py:224), so when this E-step ends the loop the second pass is not run at all
        // the total goes through shared memory: a value loaded from a uniform address is provably warp-uniform, so
        // the loop conditions below are uniform branches and the shuffles / ballots of the loop body need no
        // divergence guards (a shuffled value counts as thread-varying in the compiler's analysis)
        {
            const double t = warp_sum(llsum);
            if (lane == 0) par->ll = t;
        }
        __syncwarp();
        const double ll = par->ll;
        if (iteration > 0) improvement = ll - last;
        ++iteration;
        last = ll;
        if (!isfinite(ll)) { failed = true; break; }                     // assumption A5
        if (!(improvement > 0.1 && iteration < 1001)) break;
        double acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = 0.0;
        // per-sample responsibility sums (gmm.summaries) ride along: fit rows are pooled in sample
        // order, so a lane sees its samples in ascending order and flushes a running pair of sums
        // into its own column of red rows NP + 2 s, NP + 2 s + 1 whenever the sample changes
#pragma unroll 1
        for (int q = 0; q < 2 * R; ++q) ssum[q * 33 + lane] = 0.0;
        int cur = -1;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll kEmUnroll2
        for (int i = lane; i < n; i += 32) {
            double x[W];
            fw_row<W>(xs, cap, i, x);
            const double r0 = rb0[i], r1 = rb1[i];
            const int s = srow[i] & 0x7f;
            if (s != cur) {
                if (cur >= 0) { ssum[(2 * cur) * 33 + lane] = t0; ssum[(2 * cur + 1) * 33 + lane] = t1; }
                cur = s; t0 = 0.0; t1 = 0.0;
            }
            t0 += r0; t1 += r1;
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
        if (cur >= 0) { ssum[(2 * cur) * 33 + lane] = t0; ssum[(2 * cur + 1) * 33 + lane] = t1; }
        __syncwarp();                                   // every lane has read its responsibilities: the region is free
        // per-lane partial sums, folded across lane pairs -> 16 partials per row -> totals in red[q * FW_EM_RS + 16]
        // (per-sample rows: 32 partials, total in [32]); four interleaved chains per total
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const double v = acc[q] + __shfl_xor_sync(YC_FULL, acc[q], 16);
            if (lane < 16) red[q * FW_EM_RS + lane] = v;
        }
        __syncwarp();
#pragma unroll 1
        for (int q = lane; q < NP + 2 * R; q += 32) {
            double *rq = q < NP ? red + q * FW_EM_RS : ssum + (q - NP) * 33;
            const int cnt = q < NP ? 16 : 32;
            double ta = 0.0, tb = 0.0, tc = 0.0, td = 0.0;
#pragma unroll 1
            for (int l2 = 0; l2 < cnt; l2 += 4) { ta += rq[l2]; tb += rq[l2 + 1]; tc += rq[l2 + 2]; td += rq[l2 + 3]; }
            rq[cnt] = (ta + tb) + (tc + td);
        }
        __syncwarp();
        // ---- M-step (stats.py:202-211): blend of the shared Gaussians spread over the lanes, then lanes 0 / 1 factor
        //      their covariance while lanes 2.. prepare the weight ratios of (sample, component) from the per-sample
        //      sums (the uniform component's sum is n_s - sum r0 - sum r1: responsibilities sum to one); ONE
        //      logarithm then serves the two determinants and the weights
        {
            bool skip0, skip1;
            em_blend<W>(par, red, carry, c.wtmp, c_em_step[iteration], lane, skip0, skip1);
            double arg = 1.0, sumlog = 0.0;
            bool have = false;
            int ws = 0, wk = 0;
            if (lane < 2) {
                if (!(lane ? skip1 : skip0)) {
                    arg = em_factor<W>(par, lane, sumlog);
                    have = par->ok != 0;
#pragma unroll 1
                    for (int q = 0; q < NS; ++q) carry[lane * NS + q] = 0.0;
                }
            } else if (lane - 2 < 3 * R) {
                const int q = lane - 2;
                ws = q / 3; wk = q - 3 * ws;
                const int ns = fstart[ws + 1] - fstart[ws];
                if (ns > 0) {
                    const double t0s = ssum[(2 * ws) * 33 + 32], t1s = ssum[(2 * ws + 1) * 33 + 32];
                    const double t2s = uni ? fmax((double)ns - t0s - t1s, 0.0) : 0.0;
                    if (t0s + t1s + t2s != 0.0) {
                        const double vk = (wk == 0 ? t0s : (wk == 1 ? t1s : t2s)) + 0.5;
                        const double den = (t0s + 0.5) + (t1s + 0.5) + (uni ? t2s + 0.5 : 0.0);
                        arg = vk / den;
                        have = true;
                    }
                }
            }
            __syncwarp();
            const double lg = log(arg);                                  // one log sequence for all of them
            if (have) {
                if (lane < 2) par->cst[lane] = -0.5 * ((double)W * YC_LOG_2PI + lg + sumlog);
                else par->logw[ws][wk] = (wk == 2 && !uni) ? -INFINITY : lg;
            }
            // more than 10 samples: the remaining (sample, component) pairs
#pragma unroll 1
            for (int q = lane + 30; q < 3 * R; q += 32) {
                const int s = q / 3, k = q - 3 * s;
                const int ns = fstart[s + 1] - fstart[s];
                if (ns > 0) {
                    const double t0s = ssum[(2 * s) * 33 + 32], t1s = ssum[(2 * s + 1) * 33 + 32];
                    const double t2s = uni ? fmax((double)ns - t0s - t1s, 0.0) : 0.0;
                    if (t0s + t1s + t2s != 0.0) {
                        const double vk = (k == 0 ? t0s : (k == 1 ? t1s : t2s)) + 0.5;
                        const double den = (t0s + 0.5) + (t1s + 0.5) + (uni ? t2s + 0.5 : 0.0);
                        par->logw[s][k] = (k == 2 && !uni) ? -INFINITY : log(vk / den);
                    }
                }
            }
        }
        __syncwarp();
        failed = par->ok == 0;
    }
    // ---- results (fit_multisamp_gmm, stats.py:250-253): size-weighted mean of the sample weights,
    //      one lane per (sample, component) for the exponentials
    if (lane == 0) {
        out.iters[slot] = iteration;
        out.loglik[slot] = last;
        status[g] = failed ? YC_ST_FIT_FAILED : YC_ST_TESTED;
    }
    if (failed) return;
    __syncwarp();
#pragma unroll 1
    for (int q = lane; q < 3 * R; q += 32) {
        const int s = q / 3, k = q - 3 * s;
        const double ns = (double)(fstart[s + 1] - fstart[s]);
        red[q] = ns > 0.0 ? exp(par->logw[s][k]) * ns : 0.0;
    }
    __syncwarp();
    // the per-sample sizes (no sub-sampling on this path: fit rows = complete-window rows) travel to
    // yc_gtest_smpred_batch in preds[.][s][0], which it overwrites with the rounded counts
    for (int s2 = lane; s2 < R; s2 += 32) out.preds[((size_t)slot * R + s2) * 3] = (double)(fstart[s2 + 1] - fstart[s2]);
    if (lane < 3) {
        double cwk = 0.0;
#pragma unroll 1
        for (int s = 0; s < R; ++s) cwk += red[3 * s + lane];
        out.cw[(size_t)slot * 3 + lane] = cwk / (double)n;                // n = sum of the present samples' sizes
    }
    for (int i = lane; i < 2 * W; i += 32) out.mu[(size_t)slot * 2 * W + i] = par->mu[i / W][i % W];
    for (int i = lane; i < 2 * W * W; i += 32) out.cov[(size_t)slot * 2 * W * W + i] = par->cov[i / (W * W)][i % (W * W)];
    for (int i = lane; i < R * 3; i += 32) out.logw[(size_t)slot * R * 3 + i] = par->logw[i / 3][i % 3];
}

template <int W>
__global__ void __launch_bounds__(32 * YC_FW_WARPS, YC_FW_MINBLOCKS / YC_FW_WARPS)
em_warp_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
               const int32_t *__restrict__ fit_list, const int n_fit, const int32_t *__restrict__ fit_count,
               int *__restrict__ work, const yc_fit_out_t out, uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int lane = threadIdx.x;
    const FwSmem L = fw_layout(W, b.n_samples, cap, 1);
    unsigned char *smem_raw = smem_all;
    const int n = fw_slot_count(n_fit, fit_count);
#pragma unroll 1
    for (int slot = fw_next_slot(work, lane); slot < n; slot = fw_next_slot(work, lane)) {
        em_warp_position<W>(b, p, cap, slot, fit_list, out, status, smem_raw, L, lane);
        __syncwarp();
    }
}

// ===========================================================================
// Kernel 3 (warp): rounded counts, G-tests, fractions, posteriors
// ===========================================================================
template <int W>
__device__ __forceinline__ void
gtest_warp_position(const yc_batch_t &b, const yc_params_t &p, const int cap, const int slot,
                    const int32_t *__restrict__ fit_list, const yc_fit_out_t &out,
                    const uint8_t *__restrict__ status, double *__restrict__ post,
                    const long long post_capacity, unsigned long long *__restrict__ post_cursor,
                    unsigned char *smem_raw, const FwSmem &L, const int lane) {
    const int R = b.n_samples, Rc = b.n_cntrl;
    const int64_t g = fit_list[slot];
    const bool want_post = post != nullptr, add_uniform = p.add_uniform != 0;
    if (status[g] != YC_ST_TESTED) {                                     // failed fit: no row
        if (lane == 0) out.post_off[slot] = -1;
        if (lane < 8) out.stats[(size_t)slot * 8 + lane] = nan("");
        return;
    }
    // The rows are only needed for the posteriors; without them the per-sample sizes come from the EM
    // kernel (preds[.][s][0]) and the event data is not touched at all.
    FwCtx<W> c;
    int n = 0;
    if (want_post) { fw_front<W>(b, p, g, smem_raw, L, cap, lane, true, c); n = c.n; }
    double *spr = reinterpret_cast<double *>(smem_raw + L.red_off);     // [3R] rounded counts, then [3R] compacted
    double *cpr = spr + 3 * YC_MAX_SAMPLES;
    // preds = np.round(weights * ORIGINAL sizes)  (stats.py:296): one lane per (sample, component)
    {
        int szq[(3 * YC_MAX_SAMPLES + 31) / 32];
#pragma unroll
        for (int h = 0; h < (3 * YC_MAX_SAMPLES + 31) / 32; ++h) {
            const int q = lane + 32 * h;
            szq[h] = 0;
            if (q < 3 * R) {
                const int s2 = q / 3;
                szq[h] = want_post ? c.fstart[s2 + 1] - c.fstart[s2]
                                         : (int)out.preds[((size_t)slot * R + s2) * 3];
            }
        }
        __syncwarp();                                                    // sizes read before the counts overwrite them
#pragma unroll
        for (int h = 0; h < (3 * YC_MAX_SAMPLES + 31) / 32; ++h) {
            const int q = lane + 32 * h;
            if (q < 3 * R) {
                const double v = szq[h] > 0 ? rint(exp(out.logw[(size_t)slot * R * 3 + q]) * (double)szq[h]) : 0.0;
                out.preds[(size_t)slot * R * 3 + q] = v;
                spr[q] = szq[h] > 0 ? v : -1.0;                          // -1: sample without rows at this position
            }
        }
    }
    __syncwarp();
    long long off = -1;
    if (lane == 0) {
        int rows_c = 0, rows_t = 0, q = 0;                               // present samples only
#pragma unroll 1
        for (int s = 0; s < R; ++s) {
            if (spr[3 * s] >= 0.0) {
#pragma unroll 1
                for (int k = 0; k < 3; ++k) cpr[q * 3 + k] = spr[3 * s + k];
                ++q;
                if (s < Rc) ++rows_c; else ++rows_t;
            }
        }
        double st8[8];
        gtest_fractions(cpr, rows_c, rows_t, p.fdr_threshold, st8, add_uniform ? 3 : 2);
#pragma unroll 1
        for (int k = 0; k < 8; ++k) out.stats[(size_t)slot * 8 + k] = st8[k];
        if (want_post && st8[2] < p.fdr_threshold) {
            off = (long long)atomicAdd(post_cursor, (unsigned long long)n);
            if (off + n > post_capacity) off = -2;
        }
        out.post_off[slot] = off;
    }
    off = __shfl_sync(YC_FULL, off, 0);
    if (off < 0) return;
    // ---- posteriors of every complete-window read under the combined weights (stats.py:494-499)
    EmPar<W> *par = c.par;
    load_model<W>(par, out, slot, lane, out.cw + (size_t)slot * 3, 0);
    if (lane == 0) {
        const double *cw = out.cw + (size_t)slot * 3;
        const double t = cw[0] + cw[1] + cw[2];                          // A7
#pragma unroll 1
        for (int k = 0; k < 3; ++k) par->logw[0][k] = log(cw[k] / t);
    }
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        double x[W];
        fw_row<W>(c.xs, cap, i, x);
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
        const double l2 = (inside && add_uniform) ? par->ulogp + par->logw[0][2] : -INFINITY;
        const double mx = fmax(fmax(l0, l1), l2);
        const double e0 = exp_nonpos(l0 - mx), e1 = exp_nonpos(l1 - mx), e2 = exp_nonpos(l2 - mx);
        const double inv = 1.0 / (e0 + e1 + e2);
        post[2 * (off + i)] = e1 * inv;
        post[2 * (off + i) + 1] = e2 * inv;
    }
}

template <int W>
__global__ void __launch_bounds__(32 * YC_FW_WARPS)
gtest_warp_kernel(const yc_batch_t b, const yc_params_t p, const int cap,
                  const int32_t *__restrict__ fit_list, const int n_fit, const int32_t *__restrict__ fit_count,
                  int *__restrict__ work, const yc_fit_out_t out,
                  const uint8_t *__restrict__ status, double *__restrict__ post,
                  const long long post_capacity, unsigned long long *__restrict__ post_cursor) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int lane = threadIdx.x;
    const FwSmem L = fw_layout(W, b.n_samples, cap, 2);
    unsigned char *smem_raw = smem_all;
    const int n = fw_slot_count(n_fit, fit_count);
#pragma unroll 1
    for (int slot = fw_next_slot(work, lane); slot < n; slot = fw_next_slot(work, lane)) {
        gtest_warp_position<W>(b, p, cap, slot, fit_list, out, status, post, post_capacity, post_cursor, smem_raw, L, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int fw_cap(const yc_batch_t &b) {
    int cap = (b.max_reads + 7) & ~7;
    return cap < 32 ? 32 : cap;
}

bool fit_warp_eligible(const yc_batch_t &b, const yc_params_t &p) {
    return p.window_size <= 5 && b.max_reads <= 1024 && b.max_cond_reads <= p.max_fit_depth && b.max_cond_reads > 0;
}

static int fw_smem(const yc_batch_t &b, const yc_params_t &p, int which) {
    return fw_layout(p.window_size, b.n_samples, fw_cap(b), which).total;
}

int fit_warp_smem_bytes(const yc_batch_t &b, const yc_params_t &p) {       // per CTA, the largest of the three kernels
    return YC_FW_WARPS * std::max(fw_smem(b, p, 0), std::max(fw_smem(b, p, 1), fw_smem(b, p, 2)));
}

template <typename F>
static int dispatch_w135(int W, F &&f) {
    switch (W) {
        case 1: return f(std::integral_constant<int, 1>());
        case 3: return f(std::integral_constant<int, 3>());
        case 5: return f(std::integral_constant<int, 5>());
    }
    return YC_EINVAL;
}

// Persistent launch: the device counter the warps fetch slots from comes stream-ordered from the memory pool
// (one int per launch, zeroed, returned after the kernel), the grid is what the device holds at once.
struct FwLaunch {
    int *work = nullptr;
    int grid = 0;
    int prepare(const void *kernel, int smem, int n_fit, cudaStream_t st) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        int dev = 0, sms = 0, per_sm = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * YC_FW_WARPS, smem)) != cudaSuccess)
            return (int)e;
        if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
        const int want = (n_fit + YC_FW_WARPS - 1) / YC_FW_WARPS;
        grid = std::min(want, sms * per_sm);
        keep_pool_memory();
        if ((e = cudaMallocAsync(reinterpret_cast<void **>(&work), sizeof(int), st)) != cudaSuccess) return (int)e;
        return (int)cudaMemsetAsync(work, 0, sizeof(int), st);
    }
    int finish(cudaStream_t st) {
        const cudaError_t e = cudaGetLastError();
        const cudaError_t e2 = cudaFreeAsync(work, st);
        count_launches(1);
        return (int)(e != cudaSuccess ? e : e2);
    }
};

int init_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                       const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, int8_t *labels,
                       int max_fit_rows, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    const int cap = fw_cap(b), smem = fw_smem(b, p, 0) * YC_FW_WARPS;
    return dispatch_w135(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        void (*kern)(const yc_batch_t, const yc_params_t, const int, const int32_t *, const int, const int32_t *, int *,
                     const yc_fit_out_t, uint8_t *, int8_t *, const int);
        if (cap <= 32) kern = init_warp_kernel<W_, 1>;
        else if (cap <= 64) kern = init_warp_kernel<W_, 2>;
        else if (cap <= 128) kern = init_warp_kernel<W_, 4>;
        else if (cap <= 256) kern = init_warp_kernel<W_, 8>;
        else if (cap <= 512) kern = init_warp_kernel<W_, 16>;
        else kern = init_warp_kernel<W_, 0>;
        FwLaunch l;
        if (int e = l.prepare(reinterpret_cast<const void *>(kern), smem, n_fit, st)) return e;
        kern<<<l.grid, 32 * YC_FW_WARPS, smem, st>>>(b, p, cap, fit_list, n_fit, fit_count, l.work, out, status, labels,
                                                     max_fit_rows);
        return l.finish(st);
    });
}

int em_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                     const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    if (int e = upload_em_steps()) return e;
    const int cap = fw_cap(b), smem = fw_smem(b, p, 1) * YC_FW_WARPS;
    return dispatch_w135(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        FwLaunch l;
        if (int e = l.prepare(reinterpret_cast<const void *>(em_warp_kernel<W_>), smem, n_fit, st)) return e;
        em_warp_kernel<W_><<<l.grid, 32 * YC_FW_WARPS, smem, st>>>(b, p, cap, fit_list, n_fit, fit_count, l.work, out, status);
        return l.finish(st);
    });
}

int gtest_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                        const int32_t *fit_count, const yc_fit_out_t &out, const uint8_t *status, double *post,
                        int64_t post_capacity, int64_t *post_cursor, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    const int cap = fw_cap(b), smem = fw_smem(b, p, 2) * YC_FW_WARPS;
    return dispatch_w135(p.window_size, [&](auto tag) -> int {
        constexpr int W_ = decltype(tag)::value;
        FwLaunch l;
        if (int e = l.prepare(reinterpret_cast<const void *>(gtest_warp_kernel<W_>), smem, n_fit, st)) return e;
        gtest_warp_kernel<W_><<<l.grid, 32 * YC_FW_WARPS, smem, st>>>(
            b, p, cap, fit_list, n_fit, fit_count, l.work, out, status, post, (long long)post_capacity,
            reinterpret_cast<unsigned long long *>(post_cursor));
        return l.finish(st);
    });
}

}  // namespace yc
