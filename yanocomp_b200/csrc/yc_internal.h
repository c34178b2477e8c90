// internal launcher prototypes shared by the translation units of libyanocomp_b200
#pragma once
#include <cuda_runtime.h>
#include "../../include/yanocomp_b200.h"

namespace yc {
// kernel launches made by this library in this process (yc_kernel_launches in the C ABI)
void count_launches(int n);
// the scratch of yc_ks_batch comes from the device's default memory pool (cudaMallocAsync); keep freed
// memory in the pool instead of returning it to the driver at every synchronisation
void keep_pool_memory();
// columns per scratch chunk of yc_ks_batch: YC_KS_CHUNK in the environment overrides the built-in value
// (tests use a tiny one to cover the multi-chunk path)
int64_t ks_chunk_override(int64_t builtin);
int ks_sort_cap(const yc_batch_t &b, const yc_params_t &p);
int ks_smem_bytes(const yc_batch_t &b, const yc_params_t &p);
int ks_dispatch(const yc_batch_t &b, const yc_params_t &p,
                double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                uint8_t *status, int32_t *fit_list, int32_t *fit_count, cudaStream_t st);
bool ks_warp_eligible(const yc_batch_t &b, const yc_params_t &p);
int ks_warp_smem_bytes(const yc_batch_t &b, const yc_params_t &p);
int ks_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, double *ks_stat, double *ks_pval,
                     int32_t *n_c, int32_t *n_t, uint8_t *status, int32_t *fit_list, int32_t *fit_count,
                     cudaStream_t st);
bool fit_warp_eligible(const yc_batch_t &b, const yc_params_t &p);
int fit_warp_smem_bytes(const yc_batch_t &b, const yc_params_t &p);
int init_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                       const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, int8_t *labels, int max_fit_rows, cudaStream_t st);
int em_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                     const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, cudaStream_t st);
int gtest_warp_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                        const int32_t *fit_count, const yc_fit_out_t &out, const uint8_t *status, double *post, int64_t post_capacity,
                        int64_t *post_cursor, cudaStream_t st);
int fit_cap(const yc_batch_t &b, const yc_params_t &p);
int fit_smem_bytes(const yc_batch_t &b, const yc_params_t &p, int which);
int init_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                  const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, int8_t *labels, int max_fit_rows, cudaStream_t st);
int em_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                const int32_t *fit_count, const yc_fit_out_t &out, uint8_t *status, cudaStream_t st);
int gtest_dispatch(const yc_batch_t &b, const yc_params_t &p, const int32_t *fit_list, int n_fit,
                   const int32_t *fit_count, const yc_fit_out_t &out, const uint8_t *status, double *post, int64_t post_capacity,
                   int64_t *post_cursor, cudaStream_t st);
int assemble_dispatch(const uint16_t *raw, const int64_t *tasks, int n_tasks, int max_n_pos, const int32_t *colmap,
                      const int64_t *x_off, const int32_t *ld, uint16_t *events, int64_t n_event_elems, cudaStream_t st);
}  // namespace yc
