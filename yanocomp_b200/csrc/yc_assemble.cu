// yc_assemble_batch: per-sample event matrices (as stored in the dense sample files, yanocomp_b200/dense.py)
// -> the per-feature position-major matrices of the packed batch (include/yanocomp_b200.h).
//
// This is the cross-sample half of the reference's pivot (yanocomp/io.py:329-341: unstack per file, concat over
// files): every sample's [n_pos][ld] block is copied to rows [row0, row0 + n_reads) of its feature, column by
// column, through the sample's column map when its positions are not the feature's.  Pure data movement: the
// host only builds the small per-(feature, sample) task table.
#include <cstdint>
#include "yc_internal.h"

namespace yc {
namespace {

constexpr int AS_WARPS = 8, AS_POS_PER_CTA = 64;

__global__ void __launch_bounds__(256)
fill_nan_kernel(uint16_t *__restrict__ events, const long long n) {
    // binary16 quiet NaN = "no event"; 8 elements per thread and step
    const long long n8 = n >> 3;
    uint4 *e8 = reinterpret_cast<uint4 *>(events);
    const uint4 v = make_uint4(0x7e007e00u, 0x7e007e00u, 0x7e007e00u, 0x7e007e00u);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) e8[i] = v;
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) events[i] = 0x7e00;
}

// tasks[t] = {feature, src offset (elements of raw), unused, src ld, n_pos, n_reads, first row, colmap offset | -1}
__global__ void __launch_bounds__(32 * AS_WARPS)
assemble_kernel(const uint16_t *__restrict__ raw, const long long *__restrict__ tasks,
                const int32_t *__restrict__ colmap, const int64_t *__restrict__ x_off,
                const int32_t *__restrict__ ld, uint16_t *__restrict__ events) {
    const long long *t = tasks + (size_t)blockIdx.x * 8;
    const int n_pos = (int)t[4];
    const int p0 = blockIdx.y * AS_POS_PER_CTA;
    if (p0 >= n_pos) return;
    const int f = (int)t[0], src_ld = (int)t[3], n_reads = (int)t[5], row0 = (int)t[6];
    const long long cm_off = t[7];
    const uint16_t *src = raw + t[1];
    uint16_t *dst = events + x_off[f] + row0;
    const long long dld = ld[f];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p1 = min(n_pos, p0 + AS_POS_PER_CTA);
    for (int p = p0 + wid; p < p1; p += AS_WARPS) {
        const int c = cm_off < 0 ? p : colmap[cm_off + p];
        const uint16_t *s = src + (long long)p * src_ld;
        uint16_t *d = dst + c * dld;
        for (int r = lane; r < n_reads; r += 32) d[r] = s[r];
    }
}

}  // namespace

int assemble_dispatch(const uint16_t *raw, const int64_t *tasks, int n_tasks, int max_n_pos, const int32_t *colmap,
                      const int64_t *x_off, const int32_t *ld, uint16_t *events, int64_t n_event_elems, cudaStream_t st) {
    if (n_event_elems > 0) {
        long long blocks = (n_event_elems / 8 + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        fill_nan_kernel<<<(int)blocks, 256, 0, st>>>(events, (long long)n_event_elems);
        count_launches(1);
    }
    if (n_tasks > 0 && max_n_pos > 0) {
        const dim3 grid(n_tasks, (max_n_pos + AS_POS_PER_CTA - 1) / AS_POS_PER_CTA);
        assemble_kernel<<<grid, 32 * AS_WARPS, 0, st>>>(raw, reinterpret_cast<const long long *>(tasks), colmap, x_off, ld, events);
        count_launches(1);
    }
    return (int)cudaGetLastError();
}

}  // namespace yc
