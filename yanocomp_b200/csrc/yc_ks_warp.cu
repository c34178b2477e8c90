// Kernel 1, small-window fast path: one WARP per position, 32 consecutive
// positions per warp, no block barriers.
//
// Same arithmetic as ks_kernel (yc_ks.cu) -- see there for the reference
// citations -- reorganised after the round-1 profile (profiles/r01_summary.md):
// in the CTA-per-position kernel 127 threads waited at a barrier for the
// single-thread LAPACK axis replay and the block-wide bitonic sort paid 28
// barriers per position.  Here a warp walks a group of 32 adjacent columns in
// four phases:
//   1. per position (all 32 lanes on one position): complete-row mask, depth
//      counters, exact column sums, centred moments -> G = Z^T Z, the first w
//      rows of Z, means and stds parked in shared memory;
//   2. lane q replays dgesdd for position q -- 32 positions at once, every
//      lane busy on the serial part;
//   3. per position: projections, ordered append (ballot prefix, no atomics),
//      warp-synchronous bitonic sort, D by upper-bound searches;
//   4. lane q finalises position q: Kolmogorov p-value, gate, coalesced
//      stores, one atomic per warp to reserve fit_list slots.
// Adjacent columns share w-1 rows of the position-major matrix, so the warp's
// 32+w-1 rows stay L1-resident across the phases.
// Eligibility (checked by the dispatcher): w <= 5, reads per feature <= 1024
// and no sub-sampling possible (reads per condition <= max_fit_depth).
#include "yc_common.cuh"
#include "yc_internal.h"
#include "yc_stats.cuh"
#include "yc_svdsign.cuh"
#include "yc_ks_axis.cuh"

namespace yc {

#define YC_KW_WARPS 4
// unroll factors of the row loops of ks_moments_kernel / ks_d_kernel (projection pass)
#ifndef YC_KM_UNROLL
#define YC_KM_UNROLL 1
#endif
// c / n of the ECDF difference from per-position tables in shared memory (1) or computed per value by div_int_rn (0)
#ifndef YC_KD_TABLES
#define YC_KD_TABLES 1
#endif
#ifndef YC_KD_UNROLL
#define YC_KD_UNROLL 4
#endif

// ---------------------------------------------------------------------------
// Register-resident sort + ECDF difference (round-1 profile: the shared-memory bitonic sort was
// 35 % of this kernel's instructions and the per-element binary searches another 10 %).
//
// Both samples (each padded with +inf to P = 16 E values) sit in a[0..P) and a[P..2P).  Lane l
// loads E of them -- lanes 0-15 the control sample, lanes 16-31 the treatment sample -- and the
// warp sorts both halves at once with a bitonic network in the "mirror" formulation (first step
// of every merge compares i with i ^ (k-1), the rest i with i ^ j): the lower index always keeps
// the minimum, so there is no direction flag, strides below E are register-to-register DMNMX
// pairs and strides >= E one 64-bit shuffle + one DMNMX per value.  The sorted values go back to
// shared memory (index i at i + i/16: conflict-free 64-bit stores for the blocked layout), and
// every lane then walks ITS E consecutive sorted values: the count in the own sample is the
// index of the last equal value + 1, the count in the other sample comes from one binary search
// for the first value and a linear advance for the rest (counts are monotone).
// Returns max |c1/n1 - c2/n2| over the pooled points (np.searchsorted side='right' semantics),
// with the same separately rounded operations as before: tabA[c] = c / n1, tabB[c] = c / n2.
// Measured alternative (round 2, rejected): merging the two sorted halves with provenance tags (one more bitonic merge
// step, counts = prefix sums of the tags, no walk, no table) is bit-exact too but SLOWER: KS 3.06 -> 3.31 ms on cfg2.
#define YC_KPAD(i) ((i) + ((i) >> 4))

template <int E>
__device__ __forceinline__ double ks_sort_ecdf(double *a, const double *tabA, const double *tabB,
                                            int n1, int n2, int lane) {
    constexpr int P = 16 * E;
    const int half = lane >> 4, l16 = lane & 15;
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = a[half * P + e * 16 + l16];      // any assignment will do before the sort
    __syncwarp();
    warp_sort_regs<E>(v, lane, 16);                                      // two segments of 16 lanes each
    // sorted: global index i = lane * E + e (control sample 0..P-1, treatment sample P..2P-1)
#pragma unroll
    for (int e = 0; e < E; ++e) a[YC_KPAD(lane * E + e)] = v[e];
    __syncwarp();
#if !YC_KD_TABLES
    const double fn1 = (double)n1, fn2 = (double)n2;
    const double r1 = div_rn(1.0, fn1), r2 = div_rn(1.0, fn2);
#endif
    const int own_base = half * P, oth_base = P - own_base;
    const int n_own = half ? n2 : n1, n_oth = half ? n1 : n2;
    const int i0 = l16 * E;
    double dmax = 0.0;
    if (i0 < n_own) {
        int co;                                                          // values <= v in the other sample
        {
            int lo = 0, hi = n_oth;
            const double v0 = v[0];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a[YC_KPAD(oth_base + mid)] <= v0) lo = mid + 1; else hi = mid;
            }
            co = lo;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i < n_own) {
                const double x = v[e];
                int cs = i + 1;                                          // values <= x in the own sample
                while (cs < n_own && a[YC_KPAD(own_base + cs)] == x) ++cs;
                if (e > 0) while (co < n_oth && a[YC_KPAD(oth_base + co)] <= x) ++co;
                const int c1 = half ? co : cs, c2 = half ? cs : co;
#if YC_KD_TABLES
                dmax = max_nn(dmax, fabs(sub_rn(tabA[c1], tabB[c2])));
#else
                dmax = max_nn(dmax, fabs(sub_rn(div_int_rn((double)c1, fn1, r1), div_int_rn((double)c2, fn2, r2))));
#endif
            }
        }
    }
    __syncwarp();
    return warp_max(dmax);
}

// warp-local window scan: per-column non-null masks and the complete-row mask go to shared
// memory; afterwards one lane per (column, sample) counts its depth with popcounts over the
// sample's row range (reads are sorted by sample), instead of a per-chunk loop over samples.
//   colmask[j * nwords + wd], j < W: non-null mask of window column j; mask[wd]: complete rows
template <int W>
__device__ __forceinline__ void warp_window_scan(const Window<W> &w, int R, uint32_t *mask, uint32_t *colmask,
                                                 int nwords_max, int *cnt, int lane) {
    const int n = w.n_reads;
    for (int base = 0; base < n; base += 32) {
        const int r = base + lane;
        const bool inb = r < n;
        uint32_t m_all = YC_FULL;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const uint16_t v = inb ? __ldg(w.col[j] + r) : (uint16_t)0x7e00;
            const uint32_t m = __ballot_sync(YC_FULL, inb && !h_isnan(v));
            m_all &= m;
            if (lane == j) colmask[j * nwords_max + (base >> 5)] = m;
        }
        if (lane == W) mask[base >> 5] = m_all;
    }
    __syncwarp();
    for (int q = lane; q < (W + 1) * R; q += 32) {
        const int j = q / R, s = q - j * R;
        const uint32_t *mm = j < W ? colmask + j * nwords_max : mask;
        const int a = w.samp_start[s], e = w.samp_start[s + 1];
        int c = 0;
        for (int wd = a >> 5; wd * 32 < e; ++wd) {
            const int lo = max(a, wd * 32) - wd * 32, hi = min(e, wd * 32 + 32) - wd * 32;   // bits [lo, hi)
            const uint32_t seg = (hi - lo == 32) ? YC_FULL : (((1u << (hi - lo)) - 1u) << lo);
            c += __popc(mm[wd] & seg);
        }
        cnt[q] = c;
    }
}


// ---------------------------------------------------------------------------
// Kernel 1a: window filters + exact moments.  One CTA = 32 adjacent columns, warp w takes columns
// w, w + 4, ... (adjacent columns share w-1 rows: L1 reuse inside the CTA).  Per position, all
// 32 lanes: complete-row mask, depth counters, raw moments -> means, population stds,
// G = Z^T Z, the first W rows of Z.  Staged in shared memory and written transposed
// (scr[k * ncol + i]) so that kernels 1b / 1c read them coalesced.
// ---------------------------------------------------------------------------
struct KmSmem { int gt_off, st_off, mask_off, cm_off, cnt_off, ridx_off, total; };

__host__ __device__ inline KmSmem km_layout(int W, int R, int max_reads) {
    KmSmem s;
    const int nw = (max_reads + 31) / 32;
    int o = 0;
    s.gt_off = o;   o += 32 * (2 * W * W + 2 * W) * (int)sizeof(double);
    s.st_off = o;   o += 32 * 4 * (int)sizeof(int);
    s.mask_off = o; o += YC_KW_WARPS * nw * (int)sizeof(uint32_t);
    s.cm_off = o;   o += YC_KW_WARPS * W * nw * (int)sizeof(uint32_t);
    s.cnt_off = o;  o += YC_KW_WARPS * (W + 1) * R * (int)sizeof(int);
    s.ridx_off = o; o += YC_KW_WARPS * 16 * (int)sizeof(int);
    s.total = (o + 15) & ~15;
    return s;
}

template <int W>
__global__ void __launch_bounds__(YC_KW_WARPS * 32)
ks_moments_kernel(const yc_batch_t b, const yc_params_t p, const int64_t c0, const int ncol,
                  double *__restrict__ scr, int32_t *__restrict__ n_c_out, int32_t *__restrict__ n_t_out,
                  uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwords_max = (b.max_reads + 31) / 32;
    const KmSmem L = km_layout(W, R, b.max_reads);
    double *gt = reinterpret_cast<double *>(smem_raw + L.gt_off);
    int *st = reinterpret_cast<int *>(smem_raw + L.st_off);            // [32][4]: state, n_c, n_t
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + L.mask_off) + wid * nwords_max;
    uint32_t *colmask = reinterpret_cast<uint32_t *>(smem_raw + L.cm_off) + wid * W * nwords_max;
    int *cnt = reinterpret_cast<int *>(smem_raw + L.cnt_off) + wid * (W + 1) * R;
    int *ridx = reinterpret_cast<int *>(smem_raw + L.ridx_off) + wid * 16;
    constexpr int NT = W * (W + 1) / 2, GT = 2 * W * W + 2 * W;
    const int i0 = blockIdx.x * 32;                                     // first column of the CTA inside the chunk

#pragma unroll 1
    for (int q = wid; q < 32; q += YC_KW_WARPS) {
        const int64_t g = c0 + i0 + q;
        int state = 0, n_c = 0, n_t = 0;                                // state 0: not a window
        Window<W> w;
        if (i0 + q < ncol && window_setup<W>(b, p, g, w)) {
            const int n = w.n_reads;
            warp_window_scan<W>(w, R, mask, colmask, nwords_max, cnt, lane);
            __syncwarp();
            if (window_depth_ok<W>(w, R, Rc, p.min_read_depth, cnt, n_c, n_t))
                state = (n_c == 0 || n_t == 0) ? 1 : 2;                 // 1: degenerate
            if (state == 2) {
                const int ntot = n_c + n_t;
                // one pass: exact raw first and second moments (binary16 inputs: every partial sum is an
                // exactly representable double), centred afterwards: C_ab = S2_ab - S1_a S1_b / n
                double mu[W], cm[NT], sig[W];
#pragma unroll
                for (int j = 0; j < W; ++j) mu[j] = 0.0;
#pragma unroll
                for (int k = 0; k < NT; ++k) cm[k] = 0.0;
                // incomplete rows enter as zeros (no branch: the loads of YC_KM_UNROLL row groups are in flight together)
                YC_UNROLL(YC_KM_UNROLL)
                for (int r = lane; r < n; r += 32) {
                    const bool sel_r = (mask[r >> 5] >> (r & 31)) & 1u;
                    double x[W];
#pragma unroll
                    for (int j = 0; j < W; ++j) { const uint16_t v = __ldg(w.col[j] + r); x[j] = sel_r ? h2d(v) : 0.0; mu[j] += x[j]; }
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < W; ++a)
#pragma unroll
                        for (int c2 = a; c2 < W; ++c2) cm[k++] += x[a] * x[c2];
                }
                warp_sum_array<W>(mu);
                warp_sum_array<NT>(cm);
                // C_ab = S2_ab - (S1_a S1_b) / n, mean_j = S1_j / n, sig_a = sqrt(C_aa / n): the same IEEE operations
                // as before, but one lane per quantity (every lane used to run all 12 divisions), then broadcast
                {
                    double num = 0.0, sub = 0.0;
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < W; ++a)
#pragma unroll
                        for (int c2 = a; c2 < W; ++c2) { if (lane == k) { num = mul_rn(mu[a], mu[c2]); sub = cm[k]; } ++k; }
#pragma unroll
                    for (int j = 0; j < W; ++j) if (lane == NT + j) num = mu[j];
                    const double quo = div_rn(num, (double)ntot);
                    const double mine = lane < NT ? sub_rn(sub, quo) : quo;
#pragma unroll
                    for (int k2 = 0; k2 < NT; ++k2) cm[k2] = __shfl_sync(YC_FULL, mine, k2);
#pragma unroll
                    for (int j = 0; j < W; ++j) mu[j] = __shfl_sync(YC_FULL, mine, NT + j);
                    double sv = 1.0;
                    k = 0;
#pragma unroll
                    for (int a = 0; a < W; ++a) { if (lane == a) sv = cm[k]; k += W - a; }
                    sv = sqrt(sv / (double)ntot);
#pragma unroll
                    for (int a = 0; a < W; ++a) {
                        sig[a] = __shfl_sync(YC_FULL, sv, a);
                        if (!(sig[a] > 0.0)) state = 1;                  // constant column (assumption A9)
                    }
                }
                if (state == 2) {
                    // first W complete rows (pooled order) for the Householder pivots
                    if (lane == 0) {
                        int found = 0;
                        for (int r = 0; r < n && found < W; ++r)
                            if ((mask[r >> 5] >> (r & 31)) & 1u) ridx[found++] = r;
                        for (; found < W; ++found) ridx[found] = -1;
                    }
                    __syncwarp();
                    double *slot = gt + q * GT;                         // G[W*W], T[W*W], mu[W], sig[W]
                    // G = C / (sig sig^T): lane k divides upper-triangle entry k (one division per lane)
                    {
                        double num = 0.0, den = 1.0;
                        int ia = 0, ib = 0, k = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int c2 = a; c2 < W; ++c2) {
                                if (lane == k) { num = cm[k]; den = sig[a] * sig[c2]; ia = a * W + c2; ib = c2 * W + a; }
                                ++k;
                            }
                        if (lane < NT) { const double gv = num / den; slot[ia] = gv; slot[ib] = gv; }
                    }
                    if (lane < W) {
                        double muj = 0.0, sigj = 1.0;
#pragma unroll
                        for (int jj = 0; jj < W; ++jj) if (jj == lane) { muj = mu[jj]; sigj = sig[jj]; }
                        slot[2 * W * W + lane] = muj; slot[2 * W * W + W + lane] = sigj;
                    }
                    if (lane < W * W) {
                        const int k = lane / W, j = lane % W;
                        const int r = ridx[k];
                        double muj = 0.0, sigj = 1.0;
#pragma unroll
                        for (int jj = 0; jj < W; ++jj) if (jj == j) { muj = mu[jj]; sigj = sig[jj]; }
                        const uint16_t *cp = w.col[0];
#pragma unroll
                        for (int jj = 0; jj < W; ++jj) if (jj == j) cp = w.col[jj];
                        slot[W * W + lane] = r >= 0 ? div_rn(sub_rn(h2d(__ldg(cp + r)), muj), sigj) : 0.0;
                    }
                }
            }
        }
        if (lane == 0) { int *s4 = st + q * 4; s4[0] = state; s4[1] = state ? n_c : 0; s4[2] = state ? n_t : 0; }
        __syncwarp();
    }
    __syncthreads();
    // coalesced write-out: per-column bookkeeping, then the moment block transposed
    if (wid == 0 && i0 + lane < ncol) {
        const int64_t g = c0 + i0 + lane;
        const int state = st[lane * 4];
        n_c_out[g] = st[lane * 4 + 1]; n_t_out[g] = st[lane * 4 + 2];
        status[g] = state == 0 ? YC_ST_NOT_A_WINDOW : (state == 1 ? YC_ST_DEGENERATE : YC_ST_KS_PENDING);
    }
    for (int idx = threadIdx.x; idx < 32 * GT; idx += YC_KW_WARPS * 32) {
        const int k = idx >> 5, q = idx & 31;
        if (i0 + q < ncol && st[q * 4] == 2) scr[(size_t)k * ncol + i0 + q] = gt[q * GT + k];
    }
}

// ---------------------------------------------------------------------------
// Kernel 1c: projections, sort, D, p-value, gate.  A warp owns 32 adjacent columns: per pending
// position all lanes project the complete rows (recognised on the fly: no NaN in the W values),
// append them per condition (ballot prefix), sort and difference the ECDFs; then lane q
// finalises position q (Kolmogorov p-value, gate, coalesced stores, one atomic per warp).
// ---------------------------------------------------------------------------
struct KdSmem { int sort_off, tab_off, per_warp; };

__host__ __device__ inline KdSmem kd_layout(int P) {
    KdSmem s;
    int o = 0;
    s.sort_off = o; o += (2 * P + P / 8 + 2) * (int)sizeof(double);     // + 1 pad per 16 (register-sort write-back)
    s.tab_off = o;  o += 2 * (P + 1) * (int)sizeof(double);
    s.per_warp = (o + 15) & ~15;
    return s;
}

#ifndef YC_KD_MINBLOCKS_WIDE
#define YC_KD_MINBLOCKS_WIDE 6      // 16 values per lane in the register sort (or the long path)
#endif
#ifndef YC_KD_MINBLOCKS
#define YC_KD_MINBLOCKS 8
#endif
// SE = values per lane of the register sort (2 P / 32 with P = the batch's per-condition capacity rounded up to
// a power of two), 0 = more than 256 reads per condition.  A template parameter so that the kernel carries one
// sort variant only (instruction footprint).
template <int W, int SE>
__global__ void __launch_bounds__(YC_KW_WARPS * 32, (SE >= 16 || SE == 0) ? YC_KD_MINBLOCKS_WIDE : YC_KD_MINBLOCKS)
ks_d_kernel(const yc_batch_t b, const yc_params_t p, const int P, const int64_t c0, const int ncol,
            const double *__restrict__ scr,
            double *__restrict__ ks_stat, double *__restrict__ ks_pval,
            const int32_t *__restrict__ n_c_in, const int32_t *__restrict__ n_t_in,
            uint8_t *__restrict__ status, int32_t *__restrict__ fit_list, int32_t *__restrict__ fit_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Rc = b.n_cntrl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const KdSmem L = kd_layout(P);
    unsigned char *base_ptr = smem_raw + (size_t)wid * L.per_warp;
    double *sortbuf = reinterpret_cast<double *>(base_ptr + L.sort_off);
    double *tabA = reinterpret_cast<double *>(base_ptr + L.tab_off), *tabB = tabA + (P + 1);
    const int i0 = (blockIdx.x * YC_KW_WARPS + wid) * 32;
    if (i0 >= ncol) return;                                              // no block-level synchronisation below
    const bool mine = i0 + lane < ncol;
    const int64_t g = c0 + i0 + lane;
    const int my_state = mine ? (int)status[g] : YC_ST_NOT_A_WINDOW;
    const int my_n1 = mine ? n_c_in[g] : 0, my_n2 = mine ? n_t_in[g] : 0;
    double my_d = 0.0;
    int tab_n1 = -1, tab_n2 = -1;            // ECDF ratio tables c/n1, c/n2 (IEEE divisions), rebuilt when n changes
    const uint32_t pending = __ballot_sync(YC_FULL, my_state == YC_ST_KS_PENDING);
#pragma unroll 1
    for (uint32_t todo = pending; todo; todo &= todo - 1) {
        const int q = __ffs(todo) - 1;
        const int n1 = __shfl_sync(YC_FULL, my_n1, q), n2 = __shfl_sync(YC_FULL, my_n2, q);
        const int64_t gq = c0 + i0 + q;
        // window columns of position gq (validated by kernel 1a)
        const int f = b.col_feature[gq];
        const int c = (int)(gq - b.col_off[f]);
        const int n = b.n_reads[f];
        const int split = b.samp_start[(size_t)f * (b.n_samples + 1) + Rc];
        const int64_t ld = b.ld[f];
        const uint16_t *fbase = b.events + b.x_off[f];
        const bool rev = b.reverse[f] != 0;
        const uint16_t *col[W];
        double cj[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int cc = rev ? (c + W / 2 - j) : (c - W / 2 + j);
            col[j] = fbase + (int64_t)cc * ld;
            cj[j] = scr[(size_t)(2 * W * W + j) * ncol + i0 + q];
        }
        if (YC_KD_TABLES || SE == 0) {                       // c / n tables: the deep path always, the register path by choice
            if (n1 != tab_n1) { for (int i = lane; i <= n1; i += 32) tabA[i] = div_rn((double)i, (double)n1); tab_n1 = n1; }
            if (n2 != tab_n2) { for (int i = lane; i <= n2; i += 32) tabB[i] = div_rn((double)i, (double)n2); tab_n2 = n2; }
        }
        const int Pq = P;                                  // fixed by the batch (padding = +inf)
        for (int i = lane; i < 2 * Pq; i += 32) sortbuf[i] = INFINITY;
        __syncwarp();
        int ca = 0, cb = 0;
        const uint32_t lt = (1u << lane) - 1u;
        YC_UNROLL(YC_KD_UNROLL)
        for (int base = 0; base < n; base += 32) {
            const int r = base + lane;
            bool bit = r < n;
            const bool cond = r >= split;
            double y = 0.0;
            {
                const int rr = bit ? r : n - 1;             // loads without a branch: those of the next row group overlap
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const uint16_t h = __ldg(col[j] + rr);
                    bit &= !h_isnan(h);
                    y = fma(h2d(h), cj[j], y);
                }
            }
            const uint32_t ma = __ballot_sync(YC_FULL, bit && !cond), mb = __ballot_sync(YC_FULL, bit && cond);
            if (bit) sortbuf[cond ? Pq + cb + __popc(mb & lt) : ca + __popc(ma & lt)] = y;
            ca += __popc(ma); cb += __popc(mb);
        }
        __syncwarp();
        double dmax = 0.0;
        if (SE > 0) {
            // sort + ECDF difference in registers (2 Pq / 32 values per lane)
            dmax = ks_sort_ecdf<(SE > 0 ? SE : 2)>(sortbuf, tabA, tabB, n1, n2, lane);
        } else {
            // deep positions (> 256 reads per condition): each sample sorted through registers in chunks of
            // 512 (warp_sort_smem), one binary search per value
            warp_sort_smem(sortbuf, Pq, lane);
            warp_sort_smem(sortbuf + Pq, Pq, lane);
            const double *A = sortbuf, *B = sortbuf + Pq;
            for (int i = lane; i < n1 + n2; i += 32) {
                int c1, c2;
                if (i < n1) {
                    const double v = A[i];
                    int e = i + 1;
                    while (e < n1 && A[e] == v) ++e;
                    c1 = e; c2 = upper_bound(B, n2, v);
                } else {
                    const double v = B[i - n1];
                    int e = i - n1 + 1;
                    while (e < n2 && B[e] == v) ++e;
                    c2 = e; c1 = upper_bound(A, n1, v);
                }
                dmax = max_nn(dmax, fabs(sub_rn(tabA[c1], tabB[c2])));
            }
            dmax = warp_max(dmax);
        }
        if (lane == q) my_d = dmax;
        __syncwarp();
    }
    // ---- lane q finalises position q
    bool gated = false;
    if (mine) {
        double d = nan(""), pv = nan("");
        uint8_t code = (uint8_t)my_state;
        if (my_state == YC_ST_KS_PENDING) {
            d = my_d;
            const int n1 = my_n1, n2 = my_n2;
            const double en = __dsqrt_rn(div_rn((double)((int64_t)n1 * (int64_t)n2), (double)(n1 + n2)));
            pv = kolmogorov_sf(mul_rn(add_rn(add_rn(en, 0.12), div_rn(0.11, en)), d));
            if (!p.gmm) code = YC_ST_KS_ONLY;
            else if (ks_gate_open(d, pv, p.min_ks, p.fdr_threshold)) { code = YC_ST_GATED; gated = true; }
            else code = YC_ST_KS_FILTERED;
            status[g] = code;
        }
        ks_stat[g] = d; ks_pval[g] = pv;
    }
    (void)gated; (void)fit_list; (void)fit_count;    // fit_list is built afterwards by compact_gated (ordered)
}

static int kw_sort_cap(const yc_batch_t &b) {
    int P = 32;
    while (P < b.max_cond_reads) P <<= 1;
    return P;
}

bool ks_warp_eligible(const yc_batch_t &b, const yc_params_t &p) {
    return p.window_size <= 5 && b.max_reads <= 1024 && b.max_cond_reads <= p.max_fit_depth &&
           b.max_cond_reads > 0;
}

int ks_warp_smem_bytes(const yc_batch_t &b, const yc_params_t &p) {
    const int a = km_layout(p.window_size, b.n_samples, b.max_reads).total;
    const int d = kd_layout(kw_sort_cap(b)).per_warp * YC_KW_WARPS;
    return a > d ? a : d;
}

// Stream-ordered scratch for the moment blocks: (2 W^2 + 2 W) doubles per column of a chunk.
// Chunks bound the scratch for very large batches; the pool keeps the memory between calls.
static const int64_t KW_CHUNK = 1 << 21;

template <int W>
static int launch_kw(const yc_batch_t &b, const yc_params_t &p, double *ks_stat, double *ks_pval,
                     int32_t *n_c, int32_t *n_t, uint8_t *status, int32_t *fit_list, int32_t *fit_count,
                     cudaStream_t st) {
    constexpr int GT = 2 * W * W + 2 * W;
    keep_pool_memory();
    const int P = kw_sort_cap(b);
    const int smem_m = km_layout(W, b.n_samples, b.max_reads).total;
    const int smem_d = kd_layout(P).per_warp * YC_KW_WARPS;
    cudaError_t e = cudaFuncSetAttribute(ks_moments_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_m);
    if (e != cudaSuccess) return (int)e;
    void (*kd)(const yc_batch_t, const yc_params_t, const int, const int64_t, const int, const double *, double *, double *,
               const int32_t *, const int32_t *, uint8_t *, int32_t *, int32_t *);
    switch (P) {
        case 32:  kd = ks_d_kernel<W, 2>; break;
        case 64:  kd = ks_d_kernel<W, 4>; break;
        case 128: kd = ks_d_kernel<W, 8>; break;
        case 256: kd = ks_d_kernel<W, 16>; break;
        default:  kd = ks_d_kernel<W, 0>; break;
    }
    e = cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
    if (e != cudaSuccess) return (int)e;
    const int64_t cap_cols = ks_chunk_override(KW_CHUNK);
    const int64_t chunk = b.total_cols < cap_cols ? b.total_cols : cap_cols;
    double *scr = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void **>(&scr), (size_t)chunk * GT * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    for (int64_t c0 = 0; c0 < b.total_cols; c0 += chunk) {
        const int ncol = (int)(b.total_cols - c0 < chunk ? b.total_cols - c0 : chunk);
        ks_moments_kernel<W><<<(ncol + 31) / 32, YC_KW_WARPS * 32, smem_m, st>>>(b, p, c0, ncol, scr, n_c, n_t, status);
        count_launches(1);
        ks_axis_kernel<W><<<(ncol + 63) / 64, 64, 0, st>>>(c0, ncol, scr, status);
        count_launches(1);
        kd<<<(ncol + 32 * YC_KW_WARPS - 1) / (32 * YC_KW_WARPS), YC_KW_WARPS * 32, smem_d, st>>>(
            b, p, P, c0, ncol, scr, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count);
        count_launches(1);
    }
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(scr, st);
    return (int)(e != cudaSuccess ? e : e2);
}

int ks_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, double *ks_stat, double *ks_pval,
                     int32_t *n_c, int32_t *n_t, uint8_t *status, int32_t *fit_list, int32_t *fit_count,
                     cudaStream_t st) {
    switch (p.window_size) {
        case 1: return launch_kw<1>(b, p, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 3: return launch_kw<3>(b, p, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 5: return launch_kw<5>(b, p, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
    }
    return YC_EINVAL;
}

}  // namespace yc
