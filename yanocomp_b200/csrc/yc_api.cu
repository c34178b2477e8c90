// extern "C" surface of libyanocomp_b200.so (see include/yanocomp_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "yc_internal.h"
#include "yc_common.cuh"
#include "yc_stats.cuh"
#include "yc_svdsign.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(int e, const char *what) {
    if (e == 0) return 0;
    if (e > 0) return fail(e, "%s: %s", what, cudaGetErrorString((cudaError_t)e));
    return fail(e, "%s: invalid argument", what);
}

int check_common(const yc_batch_t *b, const yc_params_t *p) {
    if (!b || !p) return fail(YC_EINVAL, "null batch / params");
    if (b->n_features < 0 || b->total_cols < 0) return fail(YC_EINVAL, "negative sizes");
    if (b->n_samples < 2 || b->n_samples > YC_MAX_SAMPLES)
        return fail(YC_EINVAL, "n_samples=%d outside [2,%d]", b->n_samples, YC_MAX_SAMPLES);
    if (b->n_cntrl < 1 || b->n_cntrl >= b->n_samples)
        return fail(YC_EINVAL, "n_cntrl=%d must leave at least one sample per condition", b->n_cntrl);
    if (p->window_size < 1 || p->window_size > YC_MAX_WINDOW || !(p->window_size & 1))
        return fail(YC_EINVAL, "window_size=%d must be odd and <= %d", p->window_size, YC_MAX_WINDOW);
    if (p->max_fit_depth < 1) return fail(YC_EINVAL, "max_fit_depth must be >= 1");
    if (p->min_read_depth < 1) return fail(YC_EINVAL, "min_read_depth must be >= 1");
    if (b->total_cols > 0x7fffffffll) return fail(YC_EINVAL, "more than 2^31-1 candidate positions in one batch");
    if (p->covariance_type != YC_COV_FULL && p->covariance_type != YC_COV_DIAG)
        return fail(YC_EINVAL, "covariance_type=%d (only full and diag covariances are supported)", p->covariance_type);
    if (p->outlier_method != YC_OUTLIER_MAD && p->outlier_method != YC_OUTLIER_DBSCAN)
        return fail(YC_EINVAL, "outlier_method=%d", p->outlier_method);
    if (p->outlier_method == YC_OUTLIER_DBSCAN && !(p->dbscan_eps > 0.0))
        return fail(YC_EINVAL, "dbscan_eps must be positive");
    return 0;
}

int smem_limit() {
    static int limit = -1;
    if (limit < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            return 0;
        limit = v;
    }
    return limit;
}

int check_smem(int need, const char *kernel) {
    const int lim = smem_limit();
    if (lim == 0) return fail(YC_ENOGPU, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    if (need > lim)
        return fail(YC_ECAPACITY, "%s needs %d B of shared memory per position, the device offers %d B "
                    "(lower --max-fit-depth or split the feature)", kernel, need, lim);
    return 0;
}

// small windows take the warp-per-position kernels (yc_fit_warp.cu); YC_FIT_PATH=cta forces the
// general CTA-per-position kernels (the tests run both)
bool use_warp_fit(const yc_batch_t &b, const yc_params_t &p) {
    const char *force = getenv("YC_FIT_PATH");
    if (force && strcmp(force, "cta") == 0) return false;
    // the non-default model branches (diag covariance, DBSCAN outliers) live in the general kernels only
    if (p.covariance_type != YC_COV_FULL || (p.outlier_method != YC_OUTLIER_MAD && p.add_uniform)) return false;
    return yc::fit_warp_eligible(b, p) && yc::fit_warp_smem_bytes(b, p) <= smem_limit();
}

__global__ void fill_col_feature_kernel(int n_features, const int64_t *__restrict__ col_off,
                                        int32_t *__restrict__ col_feature) {
    for (int f = blockIdx.x; f < n_features; f += gridDim.x) {
        const int64_t a = col_off[f], e = col_off[f + 1];
        for (int64_t g = a + threadIdx.x; g < e; g += blockDim.x) col_feature[g] = f;
    }
}

// ---- ordered compaction of the gated positions: fit_list = ascending global columns with
// status == YC_ST_GATED, *fit_count = their number.  Three small launches (per-block counts,
// one-block exclusive scan, ordered scatter) instead of an atomic append + a sort: the slot
// order of the fit outputs is deterministic and needs no host-side pass.
constexpr int CG_THREADS = 256, CG_ITEMS = 8, CG_TILE = CG_THREADS * CG_ITEMS;

__global__ void __launch_bounds__(CG_THREADS)
compact_count_kernel(const uint8_t *__restrict__ status, long long n, int *__restrict__ block_cnt) {
    __shared__ int wsum[CG_THREADS / 32];
    const long long base = (long long)blockIdx.x * CG_TILE;
    int c = 0;
#pragma unroll
    for (int i = 0; i < CG_ITEMS; ++i) {
        const long long g = base + i * CG_THREADS + threadIdx.x;
        c += (g < n && status[g] == YC_ST_GATED);
    }
    c = yc::warp_sum_i(c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CG_THREADS / 32; ++w) t += wsum[w];
        block_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
compact_scan_kernel(int *__restrict__ block_cnt, int nb, int32_t *__restrict__ fit_count) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? block_cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(YC_FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int t = wtot[lane], s = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(YC_FULL, s, o); if (lane >= o) s += u; }
            wtot[lane] = s - t;                                          // exclusive offsets of the warps
        }
        __syncthreads();
        const int excl = carry_s + wtot[wid] + inc - v;
        if (i < nb) block_cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;                     // running total after this tile
        __syncthreads();
    }
    if (threadIdx.x == 0) *fit_count = carry_s;
}

__global__ void __launch_bounds__(CG_THREADS)
compact_scatter_kernel(const uint8_t *__restrict__ status, long long n, const int *__restrict__ block_off,
                       int32_t *__restrict__ fit_list) {
    __shared__ int wcnt[CG_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * CG_TILE;
    int run = block_off[blockIdx.x];
    for (int i = 0; i < CG_ITEMS; ++i) {
        const long long g = base + i * CG_THREADS + threadIdx.x;
        const bool hit = g < n && status[g] == YC_ST_GATED;
        const uint32_t m = __ballot_sync(YC_FULL, hit);
        if (lane == 0) wcnt[wid] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CG_THREADS / 32; ++w) { const int c = wcnt[w]; if (w < wid) before += c; total += c; }
        if (hit) fit_list[run + before + __popc(m & ((1u << lane) - 1u))] = (int32_t)g;
        run += total;
        __syncthreads();
    }
}

int compact_gated(const uint8_t *status, long long n, int32_t *fit_list, int32_t *fit_count, cudaStream_t st) {
    const int nb = (int)((n + CG_TILE - 1) / CG_TILE);
    int *block_cnt = nullptr;
    yc::keep_pool_memory();
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&block_cnt), (size_t)nb * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    compact_count_kernel<<<nb, CG_THREADS, 0, st>>>(status, n, block_cnt);
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_cnt, nb, fit_count);
    compact_scatter_kernel<<<nb, CG_THREADS, 0, st>>>(status, n, block_cnt, fit_list);
    yc::count_launches(3);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(block_cnt, st);
    return (int)(e != cudaSuccess ? e : e2);
}

// Fixed-width result rows of the gated positions in slot order: everything the host tail reads per fit as ONE
// row of YC_ROW_WIDTH(w) doubles (layout in include/yanocomp_b200.h), so that a shard's results are a single
// contiguous device->host copy and a single fixed-width block in the multi-GPU gather.
__global__ void gather_fit_rows_kernel(const int32_t *__restrict__ fit_list, int n_fit,
                                       const int32_t *__restrict__ fit_count, int w,
                                       const uint8_t *__restrict__ status, const double *__restrict__ ks_stat,
                                       const double *__restrict__ ks_pval,
                                       const int32_t *__restrict__ n_c, const int32_t *__restrict__ n_t,
                                       const double *__restrict__ mu, const double *__restrict__ cov,
                                       const double *__restrict__ stats, double *__restrict__ rows) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_fit || (fit_count != nullptr && slot >= *fit_count)) return;
    const int g = fit_list[slot];
    const int rw = YC_ROW_WIDTH(w);
    double *r = rows + (size_t)slot * rw;
    r[YC_ROW_STATUS] = (double)status[g];
    r[YC_ROW_KS_STAT] = ks_stat[g];
    r[YC_ROW_KS_PVAL] = ks_pval[g];
    r[YC_ROW_N_CNTRL] = (double)n_c[g];
    r[YC_ROW_N_ALL] = (double)(n_c[g] + n_t[g]);
    r[YC_ROW_COLUMN] = (double)g;
    for (int k = 0; k < 8; ++k) r[YC_ROW_STATS + k] = stats[(size_t)slot * 8 + k];
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < w; ++a) {
            r[YC_ROW_MU + k * w + a] = mu[((size_t)slot * 2 + k) * w + a];
            r[YC_ROW_MU + 2 * w + k * w + a] = sqrt(cov[(((size_t)slot * 2 + k) * w + a) * w + a]);
        }
}

template <int D>
__global__ void svd_v1_kernel(const double *__restrict__ G, const double *__restrict__ T,
                              double *__restrict__ v, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) yc::lapack::first_right_singular_vector<D>(G + (size_t)i * D * D, T + (size_t)i * D * D, v + (size_t)i * D);
}

}  // namespace

namespace yc {
int ks_smem_bytes(const yc_batch_t &b, const yc_params_t &p);
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int64_t ks_chunk_override(int64_t builtin) {
    const char *e = getenv("YC_KS_CHUNK");
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (int64_t)v : builtin;
}
void keep_pool_memory() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    done[dev] = true;
}
}

// exhaustive + random check of div_int_rn against the IEEE division (test hook, yc_div_int_rn_check_device)
__global__ void div_int_check_kernel(const int32_t n_max, const int64_t n_random, const uint64_t seed,
                                     unsigned long long *__restrict__ mismatches) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    // every pair 0 <= c <= n <= n_max: pair index q -> (n, c) by rows of the triangle
    const int64_t total = ((int64_t)n_max + 1) * ((int64_t)n_max + 2) / 2 - 1;          // n = 0 excluded
    for (int64_t q = tid; q < total; q += nthreads) {
        int64_t n = (int64_t)((sqrt(8.0 * (double)(q + 1) + 1.0) - 1.0) * 0.5);
        while (n * (n + 1) / 2 > q + 1) --n;
        while ((n + 1) * (n + 2) / 2 <= q + 1) ++n;
        const int64_t c = q + 1 - n * (n + 1) / 2;                                      // 0 <= c <= n, n >= 1
        const double fn = (double)n, fc = (double)c;
        bad += yc::div_int_rn(fc, fn, yc::div_rn(1.0, fn)) != yc::div_rn(fc, fn);
    }
    for (int64_t q = tid; q < n_random; q += nthreads) {
        const uint64_t h = yc::mix64(seed ^ (uint64_t)q), h2 = yc::mix64(h);
        const int64_t n = 1 + (int64_t)(h % 2147483647ull);                             // 1 .. 2^31 - 1
        const int64_t c = (int64_t)(h2 % (uint64_t)(n + 1));
        const double fn = (double)n, fc = (double)c;
        bad += yc::div_int_rn(fc, fn, yc::div_rn(1.0, fn)) != yc::div_rn(fc, fn);
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" {

const char *yc_version(void) { return "yanocomp_b200 0.1.0 (sm_100a)"; }
const char *yc_last_error(void) { return g_err; }
uint64_t yc_kernel_launches(void) { return (uint64_t)yc::g_launches.load(std::memory_order_relaxed); }

int yc_smem_bytes(const yc_batch_t *b, const yc_params_t *p, int64_t out[4]) {
    if (int e = check_common(b, p)) return e;
    out[0] = yc::ks_smem_bytes(*b, *p);
    out[1] = yc::fit_smem_bytes(*b, *p, 0);
    out[2] = yc::fit_smem_bytes(*b, *p, 1);
    out[3] = yc::fit_smem_bytes(*b, *p, 2);
    const int64_t lim = 232448;                 // 227 KB opt-in maximum of sm_100
    for (int k = 0; k < 4; ++k)
        if (out[k] > lim) return fail(YC_ECAPACITY, "kernel %d needs %lld B of shared memory (> %lld)",
                                      k, (long long)out[k], (long long)lim);
    return 0;
}

int yc_fill_col_feature(int32_t n_features, const int64_t *col_off, int32_t *col_feature, void *stream) {
    if (n_features <= 0) return 0;
    const int grid = n_features < 148 * 8 ? n_features : 148 * 8;
    fill_col_feature_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n_features, col_off, col_feature);
    yc::count_launches(1);
    return cuda_fail((int)cudaGetLastError(), "yc_fill_col_feature");
}

int yc_assemble_batch(const uint16_t *raw, const int64_t *tasks, int32_t n_tasks, int32_t max_n_pos,
                      const int32_t *colmap, const int64_t *x_off, const int32_t *ld,
                      uint16_t *events, int64_t n_event_elems, void *stream) {
    if (n_tasks < 0 || max_n_pos < 0 || n_event_elems < 0) return fail(YC_EINVAL, "yc_assemble_batch: negative size");
    if (n_event_elems == 0) return 0;
    if (!events || !x_off || !ld || (n_tasks > 0 && (!raw || !tasks || !colmap)))
        return fail(YC_EINVAL, "yc_assemble_batch: null pointer");
    if ((reinterpret_cast<uintptr_t>(events) & 15u) != 0) return fail(YC_EINVAL, "yc_assemble_batch: events must be 16-byte aligned");
    return cuda_fail(yc::assemble_dispatch(raw, tasks, n_tasks, max_n_pos, colmap, x_off, ld, events, n_event_elems,
                                           (cudaStream_t)stream), "yc_assemble_batch");
}

int yc_ks_batch(const yc_batch_t *b, const yc_params_t *p,
                double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                uint8_t *status, int32_t *fit_list, int32_t *fit_count, void *stream) {
    if (int e = check_common(b, p)) return e;
    if (!ks_stat || !ks_pval || !n_c || !n_t || !status || !fit_list || !fit_count)
        return fail(YC_EINVAL, "yc_ks_batch: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(fit_count, 0, sizeof(int32_t), st);
    if (ce != cudaSuccess) return cuda_fail((int)ce, "yc_ks_batch memset");
    if (b->total_cols == 0) return 0;
    // small windows: one warp per position (yc_ks_warp.cu); otherwise one CTA per position.
    // YC_KS_PATH=cta forces the general kernel (used by the tests to cover both).
    const char *force = getenv("YC_KS_PATH");
    const bool want_cta = force && strcmp(force, "cta") == 0;
    if (!want_cta && yc::ks_warp_eligible(*b, *p) && yc::ks_warp_smem_bytes(*b, *p) <= smem_limit()) {
        if (int e = cuda_fail(yc::ks_warp_dispatch(*b, *p, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st),
                              "yc_ks_batch (warp path)")) return e;
    } else {
        if (int e = check_smem(yc::ks_smem_bytes(*b, *p), "yc_ks_batch")) return e;
        if (int e = cuda_fail(yc::ks_dispatch(*b, *p, ks_stat, ks_pval, n_c, n_t, status, fit_list, fit_count, st),
                              "yc_ks_batch")) return e;
    }
    if (!p->gmm) return 0;                                              // --no-gmm: nothing is gated
    return cuda_fail(compact_gated(status, b->total_cols, fit_list, fit_count, st), "yc_ks_batch (compaction)");
}

int yc_compact_gated(const uint8_t *status, int64_t total_cols, int32_t *fit_list, int32_t *fit_count, void *stream) {
    if (!status || !fit_list || !fit_count || total_cols < 0) return fail(YC_EINVAL, "yc_compact_gated: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(fit_count, 0, sizeof(int32_t), st);
    if (ce != cudaSuccess) return cuda_fail((int)ce, "yc_compact_gated memset");
    if (total_cols == 0) return 0;
    return cuda_fail(compact_gated(status, total_cols, fit_list, fit_count, st), "yc_compact_gated");
}

int yc_gmm_init_batch(const yc_batch_t *b, const yc_params_t *p, const int32_t *fit_list, int32_t n_fit,
                      const int32_t *fit_count, yc_fit_out_t *out, uint8_t *status, int8_t *labels, int32_t max_fit_rows, void *stream) {
    if (int e = check_common(b, p)) return e;
    if (!out || !fit_list || !status) return fail(YC_EINVAL, "yc_gmm_init_batch: null pointer");
    if (n_fit <= 0) return 0;
    if (use_warp_fit(*b, *p))
        return cuda_fail(yc::init_warp_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, labels, max_fit_rows,
                                                (cudaStream_t)stream), "yc_gmm_init_batch (warp path)");
    if (int e = check_smem(yc::fit_smem_bytes(*b, *p, 0), "yc_gmm_init_batch")) return e;
    return cuda_fail(yc::init_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, labels, max_fit_rows,
                                       (cudaStream_t)stream), "yc_gmm_init_batch");
}

int yc_gmm_em_batch(const yc_batch_t *b, const yc_params_t *p, const int32_t *fit_list, int32_t n_fit,
                    const int32_t *fit_count, yc_fit_out_t *out, uint8_t *status, void *stream) {
    if (int e = check_common(b, p)) return e;
    if (!out || !fit_list || !status) return fail(YC_EINVAL, "yc_gmm_em_batch: null pointer");
    if (n_fit <= 0) return 0;
    if (use_warp_fit(*b, *p))
        return cuda_fail(yc::em_warp_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, (cudaStream_t)stream),
                         "yc_gmm_em_batch (warp path)");
    if (int e = check_smem(yc::fit_smem_bytes(*b, *p, 1), "yc_gmm_em_batch")) return e;
    return cuda_fail(yc::em_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, (cudaStream_t)stream),
                     "yc_gmm_em_batch");
}

int yc_gtest_smpred_batch(const yc_batch_t *b, const yc_params_t *p, const int32_t *fit_list, int32_t n_fit,
                          const int32_t *fit_count, yc_fit_out_t *out, const uint8_t *status, double *post, int64_t post_capacity,
                          int64_t *post_cursor, void *stream) {
    if (int e = check_common(b, p)) return e;
    if (!out || !fit_list || !status) return fail(YC_EINVAL, "yc_gtest_smpred_batch: null pointer");
    if (post && !post_cursor) return fail(YC_EINVAL, "yc_gtest_smpred_batch: post without post_cursor");
    cudaStream_t st = (cudaStream_t)stream;
    if (post_cursor) {
        cudaError_t ce = cudaMemsetAsync(post_cursor, 0, sizeof(int64_t), st);
        if (ce != cudaSuccess) return cuda_fail((int)ce, "yc_gtest_smpred_batch memset");
    }
    if (n_fit <= 0) return 0;
    if (use_warp_fit(*b, *p))
        return cuda_fail(yc::gtest_warp_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, post, post_capacity,
                                                 post_cursor, st), "yc_gtest_smpred_batch (warp path)");
    if (int e = check_smem(yc::fit_smem_bytes(*b, *p, 2), "yc_gtest_smpred_batch")) return e;
    return cuda_fail(yc::gtest_dispatch(*b, *p, fit_list, n_fit, fit_count, *out, status, post, post_capacity,
                                        post_cursor, st), "yc_gtest_smpred_batch");
}

int yc_gather_fit_rows(const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count, int32_t window_size,
                       const uint8_t *status, const double *ks_stat, const double *ks_pval,
                       const int32_t *n_c, const int32_t *n_t, const yc_fit_out_t *out, double *rows, void *stream) {
    if (n_fit <= 0) return 0;
    if (!fit_list || !status || !ks_stat || !ks_pval || !n_c || !n_t || !out || !out->mu || !out->cov || !out->stats || !rows)
        return fail(YC_EINVAL, "yc_gather_fit_rows: null pointer");
    gather_fit_rows_kernel<<<(n_fit + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        fit_list, n_fit, fit_count, window_size, status, ks_stat, ks_pval, n_c, n_t, out->mu, out->cov, out->stats, rows);
    yc::count_launches(1);
    return cuda_fail((int)cudaGetLastError(), "yc_gather_fit_rows");
}

int yc_gmm_fit_batch(const yc_batch_t *b, const yc_params_t *p, const int32_t *fit_list, int32_t n_fit,
                     const int32_t *fit_count, yc_fit_out_t *out, uint8_t *status, int8_t *labels, int32_t max_fit_rows,
                     double *post, int64_t post_capacity, int64_t *post_cursor, void *stream) {
    if (int e = yc_gmm_init_batch(b, p, fit_list, n_fit, fit_count, out, status, labels, max_fit_rows, stream)) return e;
    if (int e = yc_gmm_em_batch(b, p, fit_list, n_fit, fit_count, out, status, stream)) return e;
    return yc_gtest_smpred_batch(b, p, fit_list, n_fit, fit_count, out, status, post, post_capacity, post_cursor, stream);
}

// ---- host-side scalar entry points: the same __host__ __device__ code the
// kernels run, callable without a GPU so the CPU test-suite can pin it
// against SciPy (tests/test_stats_host.py).
double yc_kolmogorov_sf_host(double x) { return yc::kolmogorov_sf(x); }
double yc_chi2_sf_1dof_host(double g) { return yc::chi2_sf_1dof(g); }
void yc_gtest_fractions_host(const double *preds, int32_t nrow_c, int32_t nrow_t, double thr, double *out8) {
    yc::gtest_fractions(preds, nrow_c, nrow_t, thr, out8);
}
void yc_gtest_fractions2_host(const double *preds, int32_t nrow_c, int32_t nrow_t, double thr, double *out8) {
    yc::gtest_fractions(preds, nrow_c, nrow_t, thr, out8, 2);
}
uint64_t yc_position_key_host(uint64_t seed, uint64_t fkey, int64_t pos) {
    return yc::position_key(seed, fkey, pos);
}
uint64_t yc_mix64_host(uint64_t z) { return yc::mix64(z); }
int yc_first_right_singular_vector_device(int32_t d, const double *G, const double *T, double *v,
                                          int32_t count, void *stream) {
    if (count <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (count + 63) / 64;
    switch (d) {
        case 1: svd_v1_kernel<1><<<grid, 64, 0, st>>>(G, T, v, count); break;
        case 3: svd_v1_kernel<3><<<grid, 64, 0, st>>>(G, T, v, count); break;
        case 5: svd_v1_kernel<5><<<grid, 64, 0, st>>>(G, T, v, count); break;
        case 7: svd_v1_kernel<7><<<grid, 64, 0, st>>>(G, T, v, count); break;
        case 9: svd_v1_kernel<9><<<grid, 64, 0, st>>>(G, T, v, count); break;
        default: return fail(YC_EINVAL, "window size %d is not supported", d);
    }
    return cuda_fail((int)cudaGetLastError(), "yc_first_right_singular_vector_device");
}
int yc_div_int_rn_check_device(int32_t n_max, int64_t n_random, uint64_t seed, uint64_t *mismatches, void *stream) {
    if (!mismatches || n_max < 1 || n_random < 0) return fail(YC_EINVAL, "yc_div_int_rn_check_device: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return cuda_fail((int)e, "yc_div_int_rn_check_device memset");
    div_int_check_kernel<<<148 * 8, 256, 0, st>>>(n_max, n_random, seed, reinterpret_cast<unsigned long long *>(mismatches));
    return cuda_fail((int)cudaGetLastError(), "yc_div_int_rn_check_device");
}
int yc_first_right_singular_vector_host(int32_t d, const double *G, const double *T, double *v) {
    switch (d) {
        case 1: yc::lapack::first_right_singular_vector<1>(G, T, v); return 0;
        case 3: yc::lapack::first_right_singular_vector<3>(G, T, v); return 0;
        case 5: yc::lapack::first_right_singular_vector<5>(G, T, v); return 0;
        case 7: yc::lapack::first_right_singular_vector<7>(G, T, v); return 0;
        case 9: yc::lapack::first_right_singular_vector<9>(G, T, v); return 0;
    }
    return fail(YC_EINVAL, "window size %d is not supported", d);
}

}  // extern "C"
