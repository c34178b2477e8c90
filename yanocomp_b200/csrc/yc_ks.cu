// Kernel 1: window filters + PCA first component + two-sample KS + gate.
// One CTA (128 threads) per candidate position (global column).
//
// Reference semantics reproduced here (file:line under /root/reference/yanocomp):
//   gmmtest.py:24-61   depth filters (per-replicate depth at every window
//                      column, complete-window rows, per-replicate re-check)
//   stats.py:52-59     sub-sampling to max_fit_depth per CONDITION
//   stats.py:20-31     column standardisation (population std) + first
//                      principal component of the pooled window
//   stats.py:34-49     KS D from two sorted samples via upper-bound counts at
//                      every pooled point; p = kstwobign.sf((en+.12+.11/en) d)
//   stats.py:460-465   the gate
//
// D is a maximum over differences of integer-count ratios, so it is
// bit-exact as long as the ORDER of the projections is right; the projections
// are computed in FP64 and only their order is used.
#include "yc_common.cuh"
#include "yc_internal.h"
#include "yc_stats.cuh"
#include "yc_svdsign.cuh"
#include "yc_ks_axis.cuh"

namespace yc {

struct KsSmem {
    // carve-up of the dynamic shared memory block (all offsets in bytes)
    int sort_off, mask_off, wordoff_off, cnt_off, red_off, hist_off, sel_off, misc_off, total;
};

__host__ __device__ inline KsSmem ks_smem_layout(int W, int R, int max_reads, int sort_cap) {
    KsSmem s;
    int o = 0;
    s.sort_off = o; o += 2 * (sort_cap + sort_cap / 16) * (int)sizeof(double);   // padded sort layout (YC_PAD16)
    s.red_off = o;  o += 48 * YC_NWARPS * (int)sizeof(double);
    s.sel_off = o;  o += 4 * (int)sizeof(unsigned long long);
    s.misc_off = o; o += (16 + 2 * YC_MAX_WINDOW * YC_MAX_WINDOW) * (int)sizeof(double);
    s.mask_off = o; o += ((max_reads + 31) / 32 + 1) * (int)sizeof(uint32_t);
    s.wordoff_off = o; o += (sort_cap ? (max_reads + 31) / 32 + 1 : 0) * (int)sizeof(int);
    s.cnt_off = o;  o += (W + 1) * R * (int)sizeof(int);
    s.hist_off = o; o += 256 * (int)sizeof(int);
    s.total = (o + 15) & ~15;
    return s;
}

// The stage runs as three kernels, like the small-window path (yc_ks_warp.cu): the single-thread LAPACK axis
// replay used to hold 127 threads at a barrier for 29 % of this kernel's time (profiles/r01_summary.md).
//   ks_cta_moments_kernel  filters, sub-sampling, exact moments -> G, T, means, stds and the row mask to scratch
//   ks_axis_kernel         one thread per position: axis with numpy's sign -> c_j = v_j / sigma_j   (yc_ks_axis.cuh)
//   ks_cta_d_kernel        projections, sort, D, p-value, gate
template <int W>
__global__ void __launch_bounds__(YC_BLOCK)
ks_cta_moments_kernel(const yc_batch_t b, const yc_params_t p, const int64_t c0, const int ncol,
                      double *__restrict__ scr, uint32_t *__restrict__ mask_scr, const int nwords_max,
                      double *__restrict__ ks_stat, double *__restrict__ ks_pval,
                      int32_t *__restrict__ n_c_out, int32_t *__restrict__ n_t_out, uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl;
    const KsSmem L = ks_smem_layout(W, R, b.max_reads, 0);
    double *red = reinterpret_cast<double *>(smem_raw + L.red_off);
    unsigned long long *sel = reinterpret_cast<unsigned long long *>(smem_raw + L.sel_off);
    double *misc = reinterpret_cast<double *>(smem_raw + L.misc_off);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + L.mask_off);
    int *cnt = reinterpret_cast<int *>(smem_raw + L.cnt_off);
    int *hist = reinterpret_cast<int *>(smem_raw + L.hist_off);
    int *ridx = reinterpret_cast<int *>(misc);            // first W selected rows

    const int tid = threadIdx.x;
    const int i = blockIdx.x;                             // column inside the chunk
    const int64_t g = c0 + i;

    Window<W> w;
    const bool in_range = window_setup<W>(b, p, g, w);
    if (!in_range) {
        if (tid == 0) {
            status[g] = YC_ST_NOT_A_WINDOW;
            ks_stat[g] = nan(""); ks_pval[g] = nan("");
            n_c_out[g] = 0; n_t_out[g] = 0;
        }
        return;
    }
    for (int q = tid; q < (W + 1) * R; q += YC_BLOCK) cnt[q] = 0;
    __syncthreads();
    window_scan<W>(w, R, mask, cnt);
    __syncthreads();
    int n_c, n_t;
    const int is_window = window_depth_ok<W>(w, R, Rc, p.min_read_depth, cnt, n_c, n_t);
    if (!is_window) {
        if (tid == 0) {
            status[g] = YC_ST_NOT_A_WINDOW;
            ks_stat[g] = nan(""); ks_pval[g] = nan("");
            n_c_out[g] = 0; n_t_out[g] = 0;
        }
        return;
    }
    const int n = w.n_reads;
    const int split = w.samp_start[Rc];               // first treatment row
    // sub-sample each condition (stats.py:67-68)
    subsample_rows(mask, 0, split, n_c, p.max_fit_depth, w.pkey, YC_STREAM_KS_CNTRL, hist, sel);
    subsample_rows(mask, split, n, n_t, p.max_fit_depth, w.pkey, YC_STREAM_KS_TREAT, hist, sel);
    const int n1 = min(n_c, p.max_fit_depth), n2 = min(n_t, p.max_fit_depth);
    const int ntot = n1 + n2;
    if (n1 == 0 || n2 == 0) {
        if (tid == 0) {
            status[g] = YC_ST_DEGENERATE;
            ks_stat[g] = nan(""); ks_pval[g] = nan("");
            n_c_out[g] = n_c; n_t_out[g] = n_t;
        }
        return;
    }

    // ---- pass B: exact raw first and second moments in ONE pass (binary16 inputs: every partial sum is an exactly
    // representable double, so neither the order nor the FMA matters), centred afterwards with the same separately
    // rounded operations as ks_moments_kernel (yc_ks_warp.cu): C_ab = S2_ab - (S1_a S1_b) / n, mean_j = S1_j / n.
    // Rows outside the selection enter as zeros (no branch: the loads of four rows are in flight together).
    constexpr int NT = W * (W + 1) / 2;
    double mu[W], cm[NT];
    {
        double sm[W + NT];
#pragma unroll
        for (int k = 0; k < W + NT; ++k) sm[k] = 0.0;
#pragma unroll 4
        for (int r = tid; r < n; r += YC_BLOCK) {
            const bool sel_r = (mask[r >> 5] >> (r & 31)) & 1u;
            double x[W];
#pragma unroll
            for (int j = 0; j < W; ++j) { const uint16_t v = __ldg(w.col[j] + r); x[j] = sel_r ? h2d(v) : 0.0; }
            int k = W;
#pragma unroll
            for (int a = 0; a < W; ++a) {
                sm[a] += x[a];
#pragma unroll
                for (int c2 = a; c2 < W; ++c2) { sm[k] = fma(x[a], x[c2], sm[k]); ++k; }
            }
        }
        block_sum<W + NT>(sm, red);
        int k = 0;
#pragma unroll
        for (int a = 0; a < W; ++a) {
            mu[a] = div_rn(sm[a], (double)ntot);
#pragma unroll
            for (int c2 = a; c2 < W; ++c2) { cm[k] = sub_rn(sm[W + k], div_rn(mul_rn(sm[a], sm[c2]), (double)ntot)); ++k; }
        }
    }
    // population std per column; a zero means a constant column (assumption A9)
    double sig[W];
    bool degenerate = false;
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < W; ++a) {
            sig[a] = sqrt(cm[k] / (double)ntot);
            degenerate |= !(sig[a] > 0.0);
            k += W - a;
        }
    }
    if (degenerate) {
        if (tid == 0) {
            status[g] = YC_ST_DEGENERATE;
            ks_stat[g] = nan(""); ks_pval[g] = nan("");
            n_c_out[g] = n_c; n_t_out[g] = n_t;
        }
        return;
    }
    // ---- hand-over to the axis kernel: G = Z^T Z, the first W selected rows of Z, means, stds; and the
    //      row mask (after sub-sampling) for the D kernel
    if (tid == 0) {
        int found = 0;
        for (int wd = 0; wd * 32 < n && found < W; ++wd) {
            uint32_t m = mask[wd];
            while (m && found < W) { ridx[found++] = wd * 32 + __ffs(m) - 1; m &= m - 1; }
        }
        for (; found < W; ++found) ridx[found] = -1;
        status[g] = YC_ST_KS_PENDING;
        n_c_out[g] = n_c; n_t_out[g] = n_t;
    }
    __syncthreads();
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < W; ++a)
#pragma unroll
            for (int c2 = a; c2 < W; ++c2) {
                if (tid == k) {
                    const double gv = cm[k] / (sig[a] * sig[c2]);
                    scr[(size_t)(a * W + c2) * ncol + i] = gv;
                    scr[(size_t)(c2 * W + a) * ncol + i] = gv;
                }
                ++k;
            }
    }
    if (tid < W * W) {
        const int k = tid / W, j = tid % W;
        const int r = ridx[k];
        double muj = 0.0, sigj = 1.0;
        const uint16_t *cp = w.col[0];
#pragma unroll
        for (int jj = 0; jj < W; ++jj) if (jj == j) { muj = mu[jj]; sigj = sig[jj]; cp = w.col[jj]; }
        scr[(size_t)(W * W + tid) * ncol + i] = r >= 0 ? div_rn(sub_rn(h2d(__ldg(cp + r)), muj), sigj) : 0.0;
    }
    if (tid < W) {
        double muj = 0.0, sigj = 1.0;
#pragma unroll
        for (int jj = 0; jj < W; ++jj) if (jj == tid) { muj = mu[jj]; sigj = sig[jj]; }
        scr[(size_t)(2 * W * W + tid) * ncol + i] = muj;
        scr[(size_t)(2 * W * W + W + tid) * ncol + i] = sigj;
    }
    const int nwords = (n + 31) >> 5;
    for (int wd = tid; wd < nwords; wd += YC_BLOCK) mask_scr[(size_t)i * nwords_max + wd] = mask[wd];
}

template <int W>
__global__ void __launch_bounds__(YC_BLOCK)
ks_cta_d_kernel(const yc_batch_t b, const yc_params_t p, const int sort_cap, const int64_t c0, const int ncol,
                const double *__restrict__ scr, const uint32_t *__restrict__ mask_scr, const int nwords_max,
                double *__restrict__ ks_stat, double *__restrict__ ks_pval,
                const int32_t *__restrict__ n_c_in, const int32_t *__restrict__ n_t_in,
                uint8_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const KsSmem L = ks_smem_layout(W, b.n_samples, b.max_reads, sort_cap);
    double *sortbuf = reinterpret_cast<double *>(smem_raw + L.sort_off);
    double *red = reinterpret_cast<double *>(smem_raw + L.red_off);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + L.mask_off);
    int *wordoff = reinterpret_cast<int *>(smem_raw + L.wordoff_off);
    int *iscr = reinterpret_cast<int *>(red);
    const int tid = threadIdx.x;
    const int i = blockIdx.x;
    const int64_t g = c0 + i;
    if (status[g] != YC_ST_KS_PENDING) return;
    // window columns of position g (validated by the moments kernel)
    const int f = b.col_feature[g];
    const int c = (int)(g - b.col_off[f]);
    const int n = b.n_reads[f];
    const int64_t ld = b.ld[f];
    const uint16_t *fbase = b.events + b.x_off[f];
    const bool rev = b.reverse[f] != 0;
    const uint16_t *col[W];
    // Only the ORDER of the projections y = sum_j z_j v_j enters D, and sum_j x_j (v_j / sig_j) differs from y
    // by a constant: no centring, no division per read (the same arithmetic as ks_d_kernel, yc_ks_warp.cu)
    double cj[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int cc = rev ? (c + W / 2 - j) : (c - W / 2 + j);
        col[j] = fbase + (int64_t)cc * ld;
        cj[j] = scr[(size_t)(2 * W * W + j) * ncol + i];
    }
    const int n1 = min(n_c_in[g], p.max_fit_depth), n2 = min(n_t_in[g], p.max_fit_depth);
    const int nwords = (n + 31) >> 5;
    for (int wd = tid; wd < nwords; wd += YC_BLOCK) mask[wd] = mask_scr[(size_t)i * nwords_max + wd];
    const int P = max(32, next_pow2(max(n1, n2)));
    for (int q = tid; q < 2 * P; q += YC_BLOCK) sortbuf[YC_PAD16(q)] = INFINITY;
    __syncthreads();
    // ---- pass D: projections, appended per condition in row order: the slot of a selected row is its rank among the
    // selected rows (word prefix + popcount; controls come first) -- no atomics, and the loads of several rows overlap
    mask_word_prefix(mask, nwords, wordoff, iscr);
#pragma unroll 4
    for (int r = tid; r < n; r += YC_BLOCK) {
        const uint32_t m = mask[r >> 5];
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) y = fma(h2d(__ldg(col[j] + r)), cj[j], y);
        if ((m >> (r & 31)) & 1u) {
            const int rank = wordoff[r >> 5] + __popc(m & ((1u << (r & 31)) - 1u));
            sortbuf[YC_PAD16(rank < n1 ? rank : P + rank - n1)] = y;
        }
    }
    __syncthreads();
    block_sort_segments_regs(sortbuf, P, 2);
    // ---- D = max |F1 - F2| over the pooled points (stats.py:39-41).  Every thread walks a run of consecutive sorted
    // values of one sample: the count in the own sample is the index after the last equal value, the count in the other
    // one a binary search for the first value of the run and a linear advance afterwards (both counts are monotone).
    // c / n with the separately rounded division of the reference, through div_int_rn.
    double dmax = 0.0;
    const double fn1 = (double)n1, fn2 = (double)n2;
    const double r1 = div_rn(1.0, fn1), r2 = div_rn(1.0, fn2);
    int C = (max(n1, n2) + YC_BLOCK - 1) / YC_BLOCK;
    C |= 1;                                             // odd run length: the lanes' first values lie in distinct banks
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const int own = side ? P : 0, oth = side ? 0 : P;            // logical offsets of the two sorted samples
        const int n_own = side ? n2 : n1, n_oth = side ? n1 : n2;
        const int i0 = tid * C, i1 = min(i0 + C, n_own);
        if (i0 < n_own) {
            int co = upper_bound_pad(sortbuf, oth, n_oth, sortbuf[YC_PAD16(own + i0)]);
#pragma unroll 1
            for (int q = i0; q < i1; ++q) {
                const double x = sortbuf[YC_PAD16(own + q)];
                int cs = q + 1;
                while (cs < n_own && sortbuf[YC_PAD16(own + cs)] == x) ++cs;
                while (co < n_oth && sortbuf[YC_PAD16(oth + co)] <= x) ++co;
                const int c1 = side ? co : cs, c2 = side ? cs : co;
                dmax = fmax(dmax, fabs(sub_rn(div_int_rn((double)c1, fn1, r1), div_int_rn((double)c2, fn2, r2))));
            }
        }
    }
    dmax = block_max1(dmax, red);
    if (tid == 0) {
        // en = sqrt(n1*n2/float(n1+n2)); p = sf((en + 0.12 + 0.11/en) * d)
        const double en = __dsqrt_rn(div_rn((double)((int64_t)n1 * (int64_t)n2), (double)(n1 + n2)));
        const double arg = mul_rn(add_rn(add_rn(en, 0.12), div_rn(0.11, en)), dmax);
        const double pv = kolmogorov_sf(arg);
        ks_stat[g] = dmax; ks_pval[g] = pv;
        uint8_t st;
        if (!p.gmm) st = YC_ST_KS_ONLY;
        else if (ks_gate_open(dmax, pv, p.min_ks, p.fdr_threshold)) {
            st = YC_ST_GATED;                 // fit_list is built afterwards by compact_gated (ordered)
        } else st = YC_ST_KS_FILTERED;
        status[g] = st;
    }
}

// scratch per column of a chunk: the moment block (2 W^2 + 2 W doubles) + the row mask after sub-sampling
static const int64_t KS_CTA_CHUNK = 1 << 18;

template <int W>
static int launch_ks(const yc_batch_t &b, const yc_params_t &p, int sort_cap, int smem_d,
                     double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                     uint8_t *status, int32_t *fit_list, int32_t *fit_count, cudaStream_t st) {
    (void)fit_list; (void)fit_count;
    constexpr int GT = 2 * W * W + 2 * W;
    keep_pool_memory();
    const int smem_m = ks_smem_layout(W, b.n_samples, b.max_reads, 0).total;
    cudaError_t e = cudaFuncSetAttribute(ks_cta_moments_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_m);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_cta_d_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
    if (e != cudaSuccess) return (int)e;
    const int nwords_max = (b.max_reads + 31) / 32;
    const int64_t cap_cols = ks_chunk_override(KS_CTA_CHUNK);
    const int64_t chunk = b.total_cols < cap_cols ? b.total_cols : cap_cols;
    const size_t scr_bytes = (size_t)chunk * GT * sizeof(double);
    unsigned char *raw = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void **>(&raw), scr_bytes + (size_t)chunk * nwords_max * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    double *scr = reinterpret_cast<double *>(raw);
    uint32_t *mask_scr = reinterpret_cast<uint32_t *>(raw + scr_bytes);
    for (int64_t c0 = 0; c0 < b.total_cols; c0 += chunk) {
        const int ncol = (int)(b.total_cols - c0 < chunk ? b.total_cols - c0 : chunk);
        ks_cta_moments_kernel<W><<<ncol, YC_BLOCK, smem_m, st>>>(b, p, c0, ncol, scr, mask_scr, nwords_max,
                                                                ks_stat, ks_pval, n_c, n_t, status);
        ks_axis_kernel<W><<<(ncol + 63) / 64, 64, 0, st>>>(c0, ncol, scr, status);
        ks_cta_d_kernel<W><<<ncol, YC_BLOCK, smem_d, st>>>(b, p, sort_cap, c0, ncol, scr, mask_scr, nwords_max,
                                                          ks_stat, ks_pval, n_c, n_t, status);
        count_launches(3);
    }
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(raw, st);
    return (int)(e != cudaSuccess ? e : e2);
}

int ks_sort_cap(const yc_batch_t &b, const yc_params_t &p) {
    int cap = b.max_reads < p.max_fit_depth ? b.max_reads : p.max_fit_depth;
    int P = 1;
    while (P < cap) P <<= 1;
    return P < 32 ? 32 : P;
}

int ks_smem_bytes(const yc_batch_t &b, const yc_params_t &p) {
    return ks_smem_layout(p.window_size, b.n_samples, b.max_reads, ks_sort_cap(b, p)).total;
}

int ks_dispatch(const yc_batch_t &b, const yc_params_t &p,
                double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                uint8_t *status, int32_t *fit_list, int32_t *fit_count, cudaStream_t st) {
    const int cap = ks_sort_cap(b, p);
    const int smem = ks_smem_layout(p.window_size, b.n_samples, b.max_reads, cap).total;
    switch (p.window_size) {
        case 1: return launch_ks<1>(b, p, cap, smem, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 3: return launch_ks<3>(b, p, cap, smem, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 5: return launch_ks<5>(b, p, cap, smem, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 7: return launch_ks<7>(b, p, cap, smem, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
        case 9: return launch_ks<9>(b, p, cap, smem, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st);
    }
    return YC_EINVAL;
}

}  // namespace yc
