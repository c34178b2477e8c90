// Shared by the two KS paths (yc_ks_warp.cu, yc_ks.cu): the transient status between their kernels and
// the axis kernel.  Scratch layout: scr[k * ncol + i], k < 2 W^2 + 2 W: G (W x W), T = first W rows of Z
// (W x W), column means (W; overwritten with c_j = v_j / sigma_j), column stds (W) of column i of the chunk.
#pragma once
#include "yc_common.cuh"
#include "yc_svdsign.cuh"

// transient status between the KS kernels: passed the filters, moments stored, D pending
#define YC_ST_KS_PENDING 7

namespace yc {
namespace {

// ---------------------------------------------------------------------------
// Kernel 1b: one THREAD per position replays dgesdd on (G, T) -> first right singular vector with
// numpy's sign; stores c_j = v_j / sig_j over the mean rows (only the ORDER of the projections
// enters D, and sum_j x_j c_j differs from the standardised projection by a constant).
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(64)
ks_axis_kernel(const int64_t c0, const int ncol, double *__restrict__ scr, const uint8_t *__restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncol || status[c0 + i] != YC_ST_KS_PENDING) return;
    double GT_[2 * W * W], v[W];
#pragma unroll
    for (int k = 0; k < 2 * W * W; ++k) GT_[k] = scr[(size_t)k * ncol + i];
    lapack::first_right_singular_vector<W>(GT_, GT_ + W * W, v);
#pragma unroll
    for (int j = 0; j < W; ++j) scr[(size_t)(2 * W * W + j) * ncol + i] = v[j] / scr[(size_t)(2 * W * W + W + j) * ncol + i];
}


}  // namespace
}  // namespace yc
