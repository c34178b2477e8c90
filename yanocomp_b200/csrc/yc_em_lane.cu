// Kernel 2b, small-window fast path, third form: ONE LANE per gated position.
// EXPERIMENTAL (YC_EM_PATH=lane): parity-tested, but on cfg2 it measured 12.6 ms against 8.0 ms for the
// warp-per-position kernel -- 255 registers leave 8 warps per SM, and with ~3 positions per resident lane
// the slowest fits (up to 75 E-steps against a mean of 8) leave a long tail.  Kept as the starting point for
// a sub-warp (4-8 lanes per position) variant.
//
// Same arithmetic and schedule as em_warp_kernel (yc_fit_warp.cu) / the
// reference's em() (stats.py:194-225).  Why a third form: the warp-per-position
// kernel spends ~45 % of its instructions outside the per-read E-step --
// cross-lane reductions of 21 sufficient statistics, per-sample sums, and an
// M-step that keeps 2 of 32 lanes busy (profiles/r01_summary.md: 17.9-22 active
// lanes per instruction).  With one lane per position there is nothing to
// reduce: a lane walks the reads of ITS position sequentially, keeps the
// sufficient statistics in registers, and runs its own M-step, so all 32 lanes
// do useful work in every phase.
//
//  * Reads are fetched 8 at a time as one 16-byte load per window column (the
//    position-major rows are 16-byte aligned and NaN-padded: YC_BATCH_ROWS_ALIGNED),
//    so a warp touches 32 x w sectors per 8 reads instead of per read.
//  * Positions have very different EM lengths (2..1001 E-steps): a lane that
//    converges writes its result and immediately pulls the next position from
//    a global counter, while its neighbours keep iterating.  All positions of
//    a batch have (almost) the same number of reads, so the lanes of a warp
//    stay in step: read phase (all lanes), then end-of-iteration phase
//    (M-step for the lanes that continue, store + reload for those that stop).
#include "yc_fit_common.cuh"
#include "yc_internal.h"
#include "yc_stats.cuh"

namespace yc {

// M-step of one Gaussian on lane-private data (same maths as em_mstep in yc_fit_warp.cu).
// S = [S0, S1[W], S2 upper] already includes the carried-over summaries.
template <int W>
__device__ __forceinline__ bool lane_mstep(const double *S, double *carry, double *mu, double *cov,
                                        double *li, double &cst, double step) {
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT;
    if (S[0] < 1e-7) {                                                   // skipped, summaries kept (A3)
#pragma unroll
        for (int q = 0; q < NS; ++q) carry[q] = S[q];
        return true;
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) carry[q] = 0.0;
    const double inv0 = 1.0 / S[0], keep = 1.0 - step;
    double u[W], A[NT];
#pragma unroll
    for (int a = 0; a < W; ++a) {
        u[a] = S[1 + a] * inv0;
        mu[a] = mu[a] * step + u[a] * keep;
    }
    {
        int q = 0;
#pragma unroll
        for (int a = 0; a < W; ++a)
#pragma unroll
            for (int c2 = a; c2 < W; ++c2) {
                const double cv = (S[1 + W + q] - S[1 + a] * u[c2] - u[a] * S[1 + c2] + S[0] * (u[a] * u[c2])) * inv0;
                const double nv = cov[q] * step + cv * keep;
                cov[q] = nv;
                A[q++] = nv;
            }
    }
    double Lp[NT], rinv[W], prod = 1.0, sumlog = 0.0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double sacc = A[j * W - j * (j - 1) / 2 + (i - j)];
#pragma unroll
            for (int t = 0; t < j; ++t) sacc = fma(-Lp[i * (i + 1) / 2 + t], Lp[j * (j + 1) / 2 + t], sacc);
            if (i == j) {
                ok &= (sacc > 0.0) && isfinite(sacc);
                prod *= sacc;
                const double d = sqrt(sacc);
                Lp[i * (i + 1) / 2 + i] = d;
                rinv[i] = 1.0 / d;
            } else {
                Lp[i * (i + 1) / 2 + j] = sacc * rinv[j];
            }
        }
    }
    if (!ok) return false;                                               // A4
    if (!(prod > 1e-290 && prod < 1e290)) {
        prod = 1.0;
#pragma unroll
        for (int i = 0; i < W; ++i) sumlog += 2.0 * log(Lp[i * (i + 1) / 2 + i]);
    }
#pragma unroll
    for (int c2 = 0; c2 < W; ++c2) {
        li[c2 * (c2 + 1) / 2 + c2] = rinv[c2];
#pragma unroll
        for (int r = c2 + 1; r < W; ++r) {
            double sacc = 0.0;
#pragma unroll
            for (int t = c2; t < r; ++t) sacc = fma(-Lp[r * (r + 1) / 2 + t], li[t * (t + 1) / 2 + c2], sacc);
            li[r * (r + 1) / 2 + c2] = sacc * rinv[r];
        }
    }
    cst = -0.5 * ((double)W * YC_LOG_2PI + log(prod) + sumlog);
    return true;
}

template <int W>
__global__ void __launch_bounds__(128)
em_lane_kernel(const yc_batch_t b, const yc_params_t p, const int32_t *__restrict__ fit_list, const int n_fit,
               const yc_fit_out_t out, uint8_t *__restrict__ status, int *__restrict__ counter) {
    constexpr int NT = W * (W + 1) / 2, NS = 1 + W + NT, NP = 2 * NS + 1;
    const int R = b.n_samples;
    const bool uni = p.add_uniform != 0;

    // ---- lane-private model and bookkeeping
    double mu[2][W], li[2][NT], cst[2], blo[W], bhi[W], ul = 0.0;
    double cov[2][NT];                                   // upper triangles
    double logw[YC_MAX_SAMPLES * 3];
    double wsum[YC_MAX_SAMPLES * 2];
    int ncnt[YC_MAX_SAMPLES];
    double carry[2][NS];
    const uint16_t *col[W];
    const int32_t *samp_start = nullptr;
    int n_reads = 0, slot = -1, iteration = 0;
    int64_t g = 0;
    double improvement = INFINITY, last = 0.0;
    bool active = false, failed = false, exhausted = false;

    while (true) {
        // ------------------------------------------------ (re)load: lanes without a position pull the next one
        if (!active && !exhausted) {
            while (true) {
                slot = atomicAdd(counter, 1);
                if (slot >= n_fit) { slot = -1; exhausted = true; break; }
                g = fit_list[slot];
                if (status[g] != YC_ST_GATED) continue;                 // the initialisation failed
                Window<W> w;
                if (!window_setup<W>(b, p, g, w)) { status[g] = YC_ST_FIT_FAILED; continue; }
#pragma unroll
                for (int j = 0; j < W; ++j) col[j] = w.col[j];
                samp_start = w.samp_start;
                n_reads = w.n_reads;
                // initial model (yc_gmm_init_batch)
                double full[W * W], linv[W * W], logdet;
                failed = false;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
#pragma unroll
                    for (int a = 0; a < W; ++a) mu[k][a] = out.mu[((size_t)slot * 2 + k) * W + a];
#pragma unroll
                    for (int a = 0; a < W * W; ++a) full[a] = out.cov[((size_t)slot * 2 + k) * W * W + a];
                    {
                        int q = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int c2 = a; c2 < W; ++c2) cov[k][q++] = full[a * W + c2];
                    }
                    if (!cholesky_inverse<W>(full, linv, logdet)) failed = true;
                    {
                        int q = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int bb = 0; bb <= a; ++bb) li[k][q++] = linv[a * W + bb];
                    }
                    cst[k] = -0.5 * ((double)W * YC_LOG_2PI + logdet);
#pragma unroll
                    for (int q = 0; q < NS; ++q) carry[k][q] = 0.0;
                }
                ul = 0.0;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    blo[j] = out.box[((size_t)slot * W + j) * 2]; bhi[j] = out.box[((size_t)slot * W + j) * 2 + 1];
                    ul -= log(bhi[j] - blo[j]);
                }
#pragma unroll 1
                for (int i = 0; i < R * 3; ++i) logw[i] = out.logw[(size_t)slot * R * 3 + i];
                iteration = 0; improvement = INFINITY; last = 0.0;
                if (failed) { status[g] = YC_ST_FIT_FAILED; out.iters[slot] = 0; out.loglik[slot] = 0.0; continue; }
                active = true;
                break;
            }
        }
        if (!__any_sync(YC_FULL, active)) break;

        // ------------------------------------------------ E-step over the lane's reads, 8 per load
        double acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = 0.0;
#pragma unroll 1
        for (int i = 0; i < 2 * R; ++i) wsum[i] = 0.0;
#pragma unroll 1
        for (int i = 0; i < R; ++i) ncnt[i] = 0;
        int s = 0, next_start = 0, cnt_s = 0;
        double t0 = 0.0, t1 = 0.0, lw0 = 0.0, lw1 = 0.0, lw2 = 0.0;
        if (active) { next_start = samp_start[1]; lw0 = logw[0]; lw1 = logw[1]; lw2 = logw[2]; }
        const int nblk = active ? (n_reads + 7) >> 3 : 0;
        int nblk_max = nblk;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nblk_max = max(nblk_max, __shfl_xor_sync(YC_FULL, nblk_max, o));
#pragma unroll 1
        for (int blk = 0; blk < nblk_max; ++blk) {
            if (blk < nblk) {
                uint4 v[W];
#pragma unroll
                for (int j = 0; j < W; ++j) v[j] = __ldg(reinterpret_cast<const uint4 *>(col[j] + blk * 8));
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                    const int r = blk * 8 + i;
                    double x[W];
                    bool complete = true;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const uint16_t h = (uint16_t)(v[j].x & 0xffffu);
                        complete &= !h_isnan(h);
                        x[j] = h2d(h);
                        v[j].x = __funnelshift_r(v[j].x, v[j].y, 16);
                        v[j].y = __funnelshift_r(v[j].y, v[j].z, 16);
                        v[j].z = __funnelshift_r(v[j].z, v[j].w, 16);
                        v[j].w >>= 16;
                    }
                    while (r >= next_start && s + 1 < R) {           // reads are sorted by sample
                        wsum[2 * s] = t0; wsum[2 * s + 1] = t1; ncnt[s] = cnt_s;
                        ++s;
                        t0 = 0.0; t1 = 0.0; cnt_s = 0;
                        next_start = samp_start[s + 1];
                        lw0 = logw[3 * s]; lw1 = logw[3 * s + 1]; lw2 = logw[3 * s + 2];
                    }
                    if (complete) {
                        double q0 = 0.0, q1 = 0.0;
                        bool inside = uni;
                        int q = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a) {
                            double y0 = 0.0, y1 = 0.0;
#pragma unroll
                            for (int bb = 0; bb <= a; ++bb) {
                                y0 = fma(li[0][q], x[bb] - mu[0][bb], y0);
                                y1 = fma(li[1][q], x[bb] - mu[1][bb], y1);
                                ++q;
                            }
                            q0 = fma(y0, y0, q0); q1 = fma(y1, y1, q1);
                            inside &= (x[a] >= blo[a]) && (x[a] <= bhi[a]);
                        }
                        const double l0 = cst[0] - 0.5 * q0 + lw0;
                        const double l1 = cst[1] - 0.5 * q1 + lw1;
                        const double l2 = inside ? ul + lw2 : -INFINITY;
                        const double m = fmax(fmax(l0, l1), l2);
                        const bool m0 = l0 == m, m1 = !m0 && l1 == m;
                        const double ea = exp((m0 ? l1 : l0) - m), eb = exp(((m0 || m1) ? l2 : l1) - m);
                        const double e0 = m0 ? 1.0 : ea, e1 = m0 ? ea : (m1 ? 1.0 : eb), e2 = (m0 || m1) ? eb : 1.0;
                        const double se = e0 + e1 + e2, inv = 1.0 / se;
                        const double r0 = e0 * inv, r1 = e1 * inv;
                        acc[2 * NS] += m + log(se);
                        acc[0] += r0; acc[NS] += r1;
                        t0 += r0; t1 += r1; ++cnt_s;
                        int qq = 1 + W;
#pragma unroll
                        for (int a = 0; a < W; ++a) {
                            const double xa0 = r0 * x[a], xa1 = r1 * x[a];
                            acc[1 + a] += xa0; acc[NS + 1 + a] += xa1;
#pragma unroll
                            for (int bb = a; bb < W; ++bb) {
                                acc[qq] = fma(xa0, x[bb], acc[qq]);
                                acc[NS + qq] = fma(xa1, x[bb], acc[NS + qq]);
                                ++qq;
                            }
                        }
                    }
                }
            }
        }
        // ------------------------------------------------ end of the iteration (stats.py:200-222)
        if (active) {
            wsum[2 * s] = t0; wsum[2 * s + 1] = t1; ncnt[s] = cnt_s;
            const double ll = acc[2 * NS];
            if (iteration > 0) improvement = ll - last;
            ++iteration;
            last = ll;
            bool done = false;
            if (!isfinite(ll)) { failed = true; done = true; }          // A5
            else if (!(improvement > 0.1 && iteration < 1001)) done = true;
            if (!done) {
                // weights (stats.py:206-211): the uniform sum is n_s - sum r0 - sum r1
#pragma unroll 1
                for (int ss = 0; ss < R; ++ss) {
                    const int ns = ncnt[ss];
                    if (ns > 0) {
                        const double a0 = wsum[2 * ss], a1 = wsum[2 * ss + 1];
                        const double a2 = uni ? fmax((double)ns - a0 - a1, 0.0) : 0.0;
                        if (a0 + a1 + a2 != 0.0) {
                            const double den = (a0 + 0.5) + (a1 + 0.5) + (uni ? a2 + 0.5 : 0.0);
                            logw[3 * ss] = log((a0 + 0.5) / den);
                            logw[3 * ss + 1] = log((a1 + 0.5) / den);
                            logw[3 * ss + 2] = uni ? log((a2 + 0.5) / den) : -INFINITY;
                        }
                    }
                }
                const double step = 1.0 - (0.99 * pow(2.0 + (double)iteration, -1e-3));
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    double S[NS];
#pragma unroll
                    for (int q = 0; q < NS; ++q) S[q] = (k == 0 ? acc[q] : acc[NS + q]) + carry[k][q];
                    if (!lane_mstep<W>(S, carry[k], mu[k], cov[k], li[k], cst[k], step)) { failed = true; done = true; }
                }
            }
            if (done) {
                out.iters[slot] = iteration;
                out.loglik[slot] = last;
                if (failed) status[g] = YC_ST_FIT_FAILED;
                else {
                    status[g] = YC_ST_TESTED;
                    double cw[3] = {0.0, 0.0, 0.0}, tot = 0.0;
#pragma unroll 1
                    for (int ss = 0; ss < R; ++ss) {
                        const double ns = (double)ncnt[ss];
                        if (ns > 0.0) {
#pragma unroll 1
                            for (int k = 0; k < 3; ++k) cw[k] += exp(logw[3 * ss + k]) * ns;
                            tot += ns;
                        }
                    }
#pragma unroll 1
                    for (int k = 0; k < 3; ++k) out.cw[(size_t)slot * 3 + k] = cw[k] / tot;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
#pragma unroll
                        for (int a = 0; a < W; ++a) out.mu[((size_t)slot * 2 + k) * W + a] = mu[k][a];
                        int q = 0;
#pragma unroll
                        for (int a = 0; a < W; ++a)
#pragma unroll
                            for (int c2 = a; c2 < W; ++c2) {
                                out.cov[((size_t)slot * 2 + k) * W * W + a * W + c2] = cov[k][q];
                                out.cov[((size_t)slot * 2 + k) * W * W + c2 * W + a] = cov[k][q];
                                ++q;
                            }
                    }
#pragma unroll 1
                    for (int i = 0; i < R * 3; ++i) out.logw[(size_t)slot * R * 3 + i] = logw[i];
                }
                active = false;
            }
        }
    }
}

bool em_lane_eligible(const yc_batch_t &b, const yc_params_t &p) {
    return p.window_size <= 5 && b.max_reads <= 1024 && b.max_cond_reads <= p.max_fit_depth &&
           b.max_cond_reads > 0 && (b.flags & YC_BATCH_ROWS_ALIGNED);
}

int em_lane_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                     const yc_fit_out_t &out, uint8_t *status, cudaStream_t st) {
    if (n_fit <= 0) return 0;
    static int *counters[64] = {nullptr};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return YC_EINVAL;
    if (!counters[dev]) {
        e = cudaMalloc(&counters[dev], sizeof(int));
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaMemsetAsync(counters[dev], 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    // enough resident lanes to fill the machine (they pull positions from the counter), never more than positions
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&](auto kernel) -> int {
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
        if (per_sm < 1) per_sm = 1;
        long long lanes = (long long)sms * per_sm * 128;
        if (lanes > n_fit) lanes = n_fit;
        const int grid = (int)((lanes + 127) / 128);
        kernel<<<grid, 128, 0, st>>>(b, p, fit_list, n_fit, out, status, counters[dev]);
        count_launches(1);
        return (int)cudaGetLastError();
    };
    switch (p.window_size) {
        case 1: return launch(em_lane_kernel<1>);
        case 3: return launch(em_lane_kernel<3>);
        case 5: return launch(em_lane_kernel<5>);
    }
    return YC_EINVAL;
}

}  // namespace yc
