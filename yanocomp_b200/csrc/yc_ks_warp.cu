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
#include "yc_stats.cuh"
#include "yc_svdsign.cuh"

namespace yc {

#define YC_KW_WARPS 4

struct KwSmem {
    int sort_off, gt_off, v_off, st_off, mask_off, cnt_off, ridx_off, wc_off, cm_off, tab_off, per_warp;
};

__host__ __device__ inline KwSmem kw_layout(int W, int R, int max_reads, int P) {
    KwSmem s;
    int o = 0;
    s.sort_off = o; o += (2 * P + P / 8 + 2) * (int)sizeof(double);     // + 1 pad per 16 (register-sort write-back)
    s.tab_off = o;  o += 2 * (P + 1) * (int)sizeof(double);
    s.gt_off = o;   o += 32 * (2 * W * W + 2 * W) * (int)sizeof(double);
    s.v_off = o;    o += 32 * W * (int)sizeof(double);
    s.wc_off = o;   o += 32 * W * (int)sizeof(void *);
    s.st_off = o;   o += 32 * 8 * (int)sizeof(int);
    s.mask_off = o; o += 32 * ((max_reads + 31) / 32) * (int)sizeof(uint32_t);
    s.cnt_off = o;  o += (W + 1) * R * (int)sizeof(int);
    s.cm_off = o;   o += W * ((max_reads + 31) / 32) * (int)sizeof(uint32_t);
    s.ridx_off = o; o += 16 * (int)sizeof(int);
    s.per_warp = (o + 15) & ~15;
    return s;
}

__device__ __noinline__ void warp_bitonic(double *a, int P, int nseg, int lane) {
    const int pairs = (P * nseg) >> 1;
#pragma unroll 1
    for (int k = 2; k <= P; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
            for (int t = lane; t < pairs; t += 32) {       // every lane owns one compare-exchange
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (((i & (P - 1)) & k) == 0);
                const double x = a[i], y = a[l];
                if ((x > y) == up) { a[i] = y; a[l] = x; }
            }
            __syncwarp();
        }
    }
}

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
#define YC_KPAD(i) ((i) + ((i) >> 4))

// compare-exchange of two finite-or-inf doubles (no NaNs reach the sort): one DSETP + selects,
// instead of the NaN-aware fmin / fmax sequences
__device__ __forceinline__ void cmpx(double &lo, double &hi) {
    const bool sw = lo > hi;
    const double a = sw ? hi : lo, b = sw ? lo : hi;
    lo = a; hi = b;
}
// the lower lane of a pair keeps the smaller value, the upper lane the larger one
__device__ __forceinline__ double keep_side(double mine, double other, bool lower) {
    const bool take = (other < mine) == lower;          // equal values: either copy will do
    return take ? other : mine;
}

template <int E>
__device__ __noinline__ double ks_sort_ecdf(double *a, const double *tabA, const double *tabB,
                                            int n1, int n2, int lane) {
    constexpr int P = 16 * E;
    const int half = lane >> 4, l16 = lane & 15;
    double v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = a[half * P + e * 16 + l16];      // any assignment will do before the sort
    __syncwarp();
#pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
        // first step of the merge: partner i ^ (k - 1)
        if (k <= E) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int o = e ^ (k - 1);
                if (e < o) cmpx(v[e], v[o]);
            }
        } else {
            const int lm = k / E - 1;                                    // partner lane, values mirrored inside the lane
            const bool lower = (lane & (k / (2 * E))) == 0;
            double t[E];
#pragma unroll
            for (int e = 0; e < E; ++e) t[e] = __shfl_xor_sync(YC_FULL, v[E - 1 - e], lm);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = keep_side(v[e], t[e], lower);
        }
        // remaining steps: partner i ^ j
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j < E) {
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if ((e & j) == 0) cmpx(v[e], v[e | j]);
            } else {
                const int lm = j / E;
                const bool lower = (lane & lm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const double t = __shfl_xor_sync(YC_FULL, v[e], lm);
                    v[e] = keep_side(v[e], t, lower);
                }
            }
        }
    }
    // sorted: global index i = lane * E + e (control sample 0..P-1, treatment sample P..2P-1)
#pragma unroll
    for (int e = 0; e < E; ++e) a[YC_KPAD(lane * E + e)] = v[e];
    __syncwarp();
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
                dmax = fmax(dmax, fabs(sub_rn(tabA[c1], tabB[c2])));
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

template <int W>
__global__ void __launch_bounds__(YC_KW_WARPS * 32)
ks_warp_kernel(const yc_batch_t b, const yc_params_t p, const int P,
               double *__restrict__ ks_stat, double *__restrict__ ks_pval,
               int32_t *__restrict__ n_c_out, int32_t *__restrict__ n_t_out,
               uint8_t *__restrict__ status, int32_t *__restrict__ fit_list,
               int32_t *__restrict__ fit_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const KwSmem L = kw_layout(W, R, b.max_reads, P);
    unsigned char *base_ptr = smem_raw + (size_t)wid * L.per_warp;
    double *sortbuf = reinterpret_cast<double *>(base_ptr + L.sort_off);
    double *gt = reinterpret_cast<double *>(base_ptr + L.gt_off);
    double *vbuf = reinterpret_cast<double *>(base_ptr + L.v_off);
    const uint16_t **wcache = reinterpret_cast<const uint16_t **>(base_ptr + L.wc_off);
    int *st = reinterpret_cast<int *>(base_ptr + L.st_off);       // [32][8]: state, n_c, n_t, n_reads, split
    uint32_t *masks = reinterpret_cast<uint32_t *>(base_ptr + L.mask_off);
    int *cnt = reinterpret_cast<int *>(base_ptr + L.cnt_off);
    int *ridx = reinterpret_cast<int *>(base_ptr + L.ridx_off);
    uint32_t *colmask = reinterpret_cast<uint32_t *>(base_ptr + L.cm_off);
    double *tabA = reinterpret_cast<double *>(base_ptr + L.tab_off), *tabB = tabA + (P + 1);
    const int nwords_max = (b.max_reads + 31) / 32;
    constexpr int NT = W * (W + 1) / 2, GT = 2 * W * W + 2 * W;

    const int64_t g0 = ((int64_t)blockIdx.x * YC_KW_WARPS + wid) * 32;
    if (g0 >= b.total_cols) return;                 // no block-level synchronisation anywhere below

    // ------------------------------------------------------------ phase 1
    for (int q = 0; q < 32; ++q) {
        const int64_t g = g0 + q;
        int state = 0, n_c = 0, n_t = 0, n = 0, split = 0;    // state 0: not a window
        Window<W> w;
        if (g < b.total_cols && window_setup<W>(b, p, g, w)) {
            n = w.n_reads;
            split = w.samp_start[Rc];
            uint32_t *mask = masks + q * nwords_max;
            warp_window_scan<W>(w, R, mask, colmask, nwords_max, cnt, lane);
            __syncwarp();
            if (window_depth_ok<W>(w, R, Rc, p.min_read_depth, cnt, n_c, n_t)) {
                state = (n_c == 0 || n_t == 0) ? 1 : 2;       // 1: degenerate
            }
            if (state == 2) {
                const int ntot = n_c + n_t;
                // one pass: exact raw first and second moments (binary16 inputs: every partial sum is an
                // exactly representable double), centred afterwards: C_ab = S2_ab - S1_a S1_b / n
                double mu[W], cm[NT], sig[W];
#pragma unroll
                for (int j = 0; j < W; ++j) mu[j] = 0.0;
#pragma unroll
                for (int k = 0; k < NT; ++k) cm[k] = 0.0;
                for (int r = lane; r < n; r += 32)
                    if ((mask[r >> 5] >> (r & 31)) & 1u) {
                        double x[W];
#pragma unroll
                        for (int j = 0; j < W; ++j) { x[j] = h2d(__ldg(w.col[j] + r)); mu[j] += x[j]; }
                        int k = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int c2 = a; c2 < W; ++c2) cm[k++] += x[a] * x[c2];
                    }
#pragma unroll
                for (int j = 0; j < W; ++j) mu[j] = warp_sum(mu[j]);
                {
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < W; ++a)
#pragma unroll
                        for (int c2 = a; c2 < W; ++c2) { cm[k] = sub_rn(warp_sum(cm[k]), div_rn(mul_rn(mu[a], mu[c2]), (double)ntot)); ++k; }
                }
#pragma unroll
                for (int j = 0; j < W; ++j) mu[j] = div_rn(mu[j], (double)ntot);
                {
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < W; ++a) {
                        sig[a] = sqrt(cm[k] / (double)ntot);
                        if (!(sig[a] > 0.0)) state = 1;          // constant column (assumption A9)
                        k += W - a;
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
                    double *slot = gt + q * GT;                 // G[W*W], T[W*W], mu[W], sig[W]
                    if (lane == 0) {
                        int k = 0;
                        for (int a = 0; a < W; ++a)
                            for (int c2 = a; c2 < W; ++c2) {
                                const double gv = cm[k++] / (sig[a] * sig[c2]);
                                slot[a * W + c2] = gv; slot[c2 * W + a] = gv;
                            }
                        for (int a = 0; a < W; ++a) { slot[2 * W * W + a] = mu[a]; slot[2 * W * W + W + a] = sig[a]; }
                    }
                    if (lane < W * W) {
                        const int k = lane / W, j = lane % W;
                        const int r = ridx[k];
                        // mu / sig are identical in every lane; select without dynamic indexing
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
#pragma unroll
            for (int j = 0; j < W; ++j) if (lane == j) wcache[q * W + j] = w.col[j];
        }
        if (lane == 0) { int *s = st + q * 8; s[0] = state; s[1] = n_c; s[2] = n_t; s[3] = n; s[4] = split; }
        __syncwarp();
    }
    // ------------------------------------------------------------ phase 2
    if (st[lane * 8] == 2) {
        const double *slot = gt + lane * GT;
        lapack::first_right_singular_vector<W>(slot, slot + W * W, vbuf + lane * W);
    }
    __syncwarp();
    // ------------------------------------------------------------ phase 3
    double my_d = 0.0;
    int tab_n1 = -1, tab_n2 = -1;           // ECDF ratio tables c/n1, c/n2 (IEEE divisions), rebuilt when n changes
    for (int q = 0; q < 32; ++q) {
        const int *s = st + q * 8;
        if (s[0] != 2) continue;                                // warp-uniform
        const int n1 = s[1], n2 = s[2], n = s[3], split = s[4];
        const uint32_t *mask = masks + q * nwords_max;
        const double *slot = gt + q * GT;
        // Only the ORDER of the projections y = sum_j z_j v_j enters D, and
        // sum_j x_j (v_j / sig_j) differs from y by a constant: no centring, no division per read.
        double cj[W];
        const uint16_t *col[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            cj[j] = vbuf[q * W + j] / slot[2 * W * W + W + j];
            col[j] = wcache[q * W + j];
        }
        if (n1 != tab_n1) { for (int i = lane; i <= n1; i += 32) tabA[i] = div_rn((double)i, (double)n1); tab_n1 = n1; }
        if (n2 != tab_n2) { for (int i = lane; i <= n2; i += 32) tabB[i] = div_rn((double)i, (double)n2); tab_n2 = n2; }
        const int Pq = max(32, next_pow2(max(n1, n2)));
        for (int i = lane; i < 2 * Pq; i += 32) sortbuf[i] = INFINITY;
        __syncwarp();
        int ca = 0, cb = 0;
        const uint32_t lt = (1u << lane) - 1u;
        for (int base = 0; base < n; base += 32) {
            const int r = base + lane;
            const bool bit = r < n && ((mask[r >> 5] >> (r & 31)) & 1u);
            const bool cond = r >= split;
            double y = 0.0;
            if (bit) {
#pragma unroll
                for (int j = 0; j < W; ++j) y = fma(h2d(__ldg(col[j] + r)), cj[j], y);
            }
            const uint32_t ma = __ballot_sync(YC_FULL, bit && !cond), mb = __ballot_sync(YC_FULL, bit && cond);
            if (bit) sortbuf[cond ? Pq + cb + __popc(mb & lt) : ca + __popc(ma & lt)] = y;
            ca += __popc(ma); cb += __popc(mb);
        }
        __syncwarp();
        double dmax = 0.0;
        if (Pq <= 256) {
            // sort + ECDF difference in registers (2 Pq / 32 values per lane)
            switch (Pq) {
                case 32:  dmax = ks_sort_ecdf<2>(sortbuf, tabA, tabB, n1, n2, lane); break;
                case 64:  dmax = ks_sort_ecdf<4>(sortbuf, tabA, tabB, n1, n2, lane); break;
                case 128: dmax = ks_sort_ecdf<8>(sortbuf, tabA, tabB, n1, n2, lane); break;
                default:  dmax = ks_sort_ecdf<16>(sortbuf, tabA, tabB, n1, n2, lane); break;
            }
        } else {
            // deep positions (> 256 reads per condition): shared-memory sort, one binary search per value
            warp_bitonic(sortbuf, Pq, 2, lane);
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
                dmax = fmax(dmax, fabs(sub_rn(tabA[c1], tabB[c2])));
            }
            dmax = warp_max(dmax);
        }
        if (lane == q) my_d = dmax;
        __syncwarp();
    }
    // ------------------------------------------------------------ phase 4
    const int64_t g = g0 + lane;
    bool gated = false;
    if (g < b.total_cols) {
        const int *s = st + lane * 8;
        const int state = s[0], n1 = s[1], n2 = s[2];
        double d = nan(""), pv = nan("");
        uint8_t code = YC_ST_NOT_A_WINDOW;
        if (state == 1) code = YC_ST_DEGENERATE;
        else if (state == 2) {
            d = my_d;
            const double en = __dsqrt_rn(div_rn((double)((int64_t)n1 * (int64_t)n2), (double)(n1 + n2)));
            pv = kolmogorov_sf(mul_rn(add_rn(add_rn(en, 0.12), div_rn(0.11, en)), d));
            if (!p.gmm) code = YC_ST_KS_ONLY;
            else if (d >= p.min_ks && pv < p.fdr_threshold) { code = YC_ST_GATED; gated = true; }
            else code = YC_ST_KS_FILTERED;
        }
        ks_stat[g] = d; ks_pval[g] = pv;
        n_c_out[g] = state ? n1 : 0; n_t_out[g] = state ? n2 : 0;
        status[g] = code;
    }
    const uint32_t gm = __ballot_sync(YC_FULL, gated);
    if (gm) {
        int base = 0;
        if (lane == 0) base = atomicAdd(fit_count, __popc(gm));
        base = __shfl_sync(YC_FULL, base, 0);
        if (gated) fit_list[base + __popc(gm & ((1u << lane) - 1u))] = (int32_t)g;
    }
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
    return kw_layout(p.window_size, b.n_samples, b.max_reads, kw_sort_cap(b)).per_warp * YC_KW_WARPS;
}

template <int W>
static int launch_kw(const yc_batch_t &b, const yc_params_t &p, double *ks_stat, double *ks_pval,
                     int32_t *n_c, int32_t *n_t, uint8_t *status, int32_t *fit_list, int32_t *fit_count,
                     cudaStream_t st) {
    const int P = kw_sort_cap(b);
    const int smem = kw_layout(W, b.n_samples, b.max_reads, P).per_warp * YC_KW_WARPS;
    cudaError_t e = cudaFuncSetAttribute(ks_warp_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t groups = (b.total_cols + 31) / 32;
    const unsigned grid = (unsigned)((groups + YC_KW_WARPS - 1) / YC_KW_WARPS);
    ks_warp_kernel<W><<<grid, YC_KW_WARPS * 32, smem, st>>>(b, p, P, ks_stat, ks_pval, n_c, n_t, status,
                                                            fit_list, fit_count);
    return (int)cudaGetLastError();
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
