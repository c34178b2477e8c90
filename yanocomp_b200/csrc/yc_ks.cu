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

namespace yc {

struct KsSmem {
    // carve-up of the dynamic shared memory block (all offsets in bytes)
    int sort_off, mask_off, cnt_off, red_off, hist_off, sel_off, misc_off, total;
};

__host__ __device__ inline KsSmem ks_smem_layout(int W, int R, int max_reads, int sort_cap) {
    KsSmem s;
    int o = 0;
    s.sort_off = o; o += 2 * sort_cap * (int)sizeof(double);
    s.red_off = o;  o += 48 * YC_NWARPS * (int)sizeof(double);
    s.sel_off = o;  o += 4 * (int)sizeof(unsigned long long);
    s.misc_off = o; o += (16 + 2 * YC_MAX_WINDOW * YC_MAX_WINDOW) * (int)sizeof(double);
    s.mask_off = o; o += ((max_reads + 31) / 32 + 1) * (int)sizeof(uint32_t);
    s.cnt_off = o;  o += (W + 1) * R * (int)sizeof(int);
    s.hist_off = o; o += 256 * (int)sizeof(int);
    s.total = (o + 15) & ~15;
    return s;
}

template <int W>
__global__ void __launch_bounds__(YC_BLOCK)
ks_kernel(const yc_batch_t b, const yc_params_t p, const int sort_cap,
          double *__restrict__ ks_stat, double *__restrict__ ks_pval,
          int32_t *__restrict__ n_c_out, int32_t *__restrict__ n_t_out,
          uint8_t *__restrict__ status, int32_t *__restrict__ fit_list,
          int32_t *__restrict__ fit_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = b.n_samples, Rc = b.n_cntrl;
    const KsSmem L = ks_smem_layout(W, R, b.max_reads, sort_cap);
    double *sortbuf = reinterpret_cast<double *>(smem_raw + L.sort_off);
    double *red = reinterpret_cast<double *>(smem_raw + L.red_off);
    unsigned long long *sel = reinterpret_cast<unsigned long long *>(smem_raw + L.sel_off);
    double *misc = reinterpret_cast<double *>(smem_raw + L.misc_off);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem_raw + L.mask_off);
    int *cnt = reinterpret_cast<int *>(smem_raw + L.cnt_off);
    int *hist = reinterpret_cast<int *>(smem_raw + L.hist_off);
    int *icnt = reinterpret_cast<int *>(misc + 12);     // two append cursors

    const int tid = threadIdx.x;
    const int64_t g = blockIdx.x;

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
    for (int i = tid; i < (W + 1) * R; i += YC_BLOCK) cnt[i] = 0;
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

    // ---- pass B: column sums (exact for binary16 inputs) -> means
    double mu[W];
    {
        double s[W];
#pragma unroll
        for (int j = 0; j < W; ++j) s[j] = 0.0;
        for (int r = tid; r < n; r += YC_BLOCK) {
            if ((mask[r >> 5] >> (r & 31)) & 1u) {
#pragma unroll
                for (int j = 0; j < W; ++j) s[j] += h2d(__ldg(w.col[j] + r));
            }
        }
        block_sum<W>(s, red);
#pragma unroll
        for (int j = 0; j < W; ++j) mu[j] = div_rn(s[j], (double)ntot);
    }
    // ---- pass C: centred second moments (upper triangle)
    constexpr int NT = W * (W + 1) / 2;
    double cm[NT];
    {
#pragma unroll
        for (int k = 0; k < NT; ++k) cm[k] = 0.0;
        for (int r = tid; r < n; r += YC_BLOCK) {
            if ((mask[r >> 5] >> (r & 31)) & 1u) {
                double dlt[W];
#pragma unroll
                for (int j = 0; j < W; ++j) dlt[j] = h2d(__ldg(w.col[j] + r)) - mu[j];
                int k = 0;
#pragma unroll
                for (int a = 0; a < W; ++a)
#pragma unroll
                    for (int c2 = a; c2 < W; ++c2) cm[k++] += dlt[a] * dlt[c2];
            }
        }
        block_sum<NT>(cm, red);
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
    // ---- first principal axis of the standardised data, oriented as LAPACK's dgesdd
    // orients it (yc_svdsign.cuh): needs G = Z^T Z and the first W pooled rows of Z
    if (tid == 0) {
        // G and T live in shared memory: handing thread-local arrays to the (large) routine
        // was observed to be miscompiled (the callee's zero-initialised workspace aliased T)
        double *G = misc + 16, *T = misc + 16 + W * W;
        double v[W];
        int k = 0;
        for (int a = 0; a < W; ++a)
            for (int c2 = a; c2 < W; ++c2) {
                const double gv = cm[k++] / (sig[a] * sig[c2]);
                G[a * W + c2] = gv; G[c2 * W + a] = gv;
            }
        int found = 0;
        for (int r = 0; r < n && found < W; ++r) {
            if ((mask[r >> 5] >> (r & 31)) & 1u) {
#pragma unroll
                for (int j = 0; j < W; ++j)
                    T[found * W + j] = div_rn(sub_rn(h2d(__ldg(w.col[j] + r)), mu[j]), sig[j]);
                ++found;
            }
        }
        for (; found < W; ++found) for (int j = 0; j < W; ++j) T[found * W + j] = 0.0;
        lapack::first_right_singular_vector<W>(G, T, v);
        for (int a = 0; a < W; ++a) misc[a] = v[a];
        icnt[0] = 0; icnt[1] = 0;
    }
    const int P = max(32, next_pow2(max(n1, n2)));
    for (int i = tid; i < 2 * P; i += YC_BLOCK) sortbuf[i] = INFINITY;
    __syncthreads();
    // Only the ORDER of the projections y = sum_j z_j v_j enters D, and sum_j x_j (v_j / sig_j) differs from y
    // by a constant: no centring, no division per read (the same arithmetic as ks_d_kernel, yc_ks_warp.cu)
    double cj[W];
#pragma unroll
    for (int j = 0; j < W; ++j) cj[j] = misc[j] / sig[j];
    // ---- pass D: projections, appended per condition (order is irrelevant before the sort)
    for (int r = tid; r < n; r += YC_BLOCK) {
        if ((mask[r >> 5] >> (r & 31)) & 1u) {
            double y = 0.0;
#pragma unroll
            for (int j = 0; j < W; ++j) y = fma(h2d(__ldg(w.col[j] + r)), cj[j], y);
            const int cond = r >= split;
            const int slot = atomicAdd(&icnt[cond], 1);
            sortbuf[cond * P + slot] = y;
        }
    }
    __syncthreads();
    block_sort_segments_regs(sortbuf, P, 2);
    // ---- D = max |F1 - F2| over the pooled points (stats.py:39-41)
    const double *A = sortbuf, *B = sortbuf + P;
    double dmax = 0.0;
    const double fn1 = (double)n1, fn2 = (double)n2;
    for (int i = tid; i < ntot; i += YC_BLOCK) {
        const double v = (i < n1) ? A[i] : B[i - n1];
        const int c1 = upper_bound(A, n1, v), c2 = upper_bound(B, n2, v);
        const double df = fabs(sub_rn(div_rn((double)c1, fn1), div_rn((double)c2, fn2)));
        dmax = fmax(dmax, df);
    }
    dmax = block_max1(dmax, red);
    if (tid == 0) {
        // en = sqrt(n1*n2/float(n1+n2)); p = sf((en + 0.12 + 0.11/en) * d)
        const double en = __dsqrt_rn(div_rn((double)((int64_t)n1 * (int64_t)n2), (double)(n1 + n2)));
        const double arg = mul_rn(add_rn(add_rn(en, 0.12), div_rn(0.11, en)), dmax);
        const double pv = kolmogorov_sf(arg);
        ks_stat[g] = dmax; ks_pval[g] = pv;
        n_c_out[g] = n_c; n_t_out[g] = n_t;
        uint8_t st;
        if (!p.gmm) st = YC_ST_KS_ONLY;
        else if (dmax >= p.min_ks && pv < p.fdr_threshold) {
            st = YC_ST_GATED;                 // fit_list is built afterwards by compact_gated (ordered)
        } else st = YC_ST_KS_FILTERED;
        status[g] = st;
    }
}

template <int W>
static int launch_ks(const yc_batch_t &b, const yc_params_t &p, int sort_cap, int smem,
                     double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                     uint8_t *status, int32_t *fit_list, int32_t *fit_count, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(ks_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    ks_kernel<W><<<(unsigned)b.total_cols, YC_BLOCK, smem, st>>>(
        b, p, sort_cap, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count);
    count_launches(1);
    return (int)cudaGetLastError();
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
