/*
 * yanocomp_b200 -- C ABI of the B200 (sm_100a) implementation of the
 * `yanocomp gmmtest` per-position hot path.
 *
 * The reference (bartongroup/yanocomp) is pure Python and has no FFI layer;
 * its boundary for this path is the Python function
 *     parallel_test(opts) -> (DataFrame, sm_preds)      yanocomp/gmmtest.py:182
 * which loops test_chunk (gmmtest.py:139) -> iter_positions (gmmtest.py:81)
 * -> position_stats (yanocomp/stats.py:445) once per (feature, position).
 * The entry points below are what a ctypes binding placed inside
 * `test_chunk` would call instead of that loop.  Every
 * function cites the reference code it replaces.
 *
 * Conventions
 *   - plain C, no exceptions across the ABI; every call returns 0 on success
 *     or a negative YC_E* code / positive cudaError_t, and
 *     yc_last_error() gives a message for the calling thread;
 *   - unless a name ends in `_host`, every pointer is a DEVICE pointer owned
 *     by the caller, and `stream` is a cudaStream_t passed as void* (NULL =
 *     the legacy default stream).  Calls are asynchronous on that stream;
 *   - per-position arrays are structure-of-arrays indexed by the GLOBAL
 *     COLUMN index g = col_off[f] + c  (feature f, column c);
 *   - per-fit arrays are indexed by the SLOT of the position in `fit_list`.
 *
 * Packed ragged batch (what the reworked io.py builds from the prep HDF5,
 * reference yanocomp/io.py:296-343):
 *   feature f owns a position-major fp16 matrix
 *       events[x_off[f] + c*ld[f] + r]     c in [0,n_cols[f]), r in [0,n_reads[f])
 *   NaN = the read has no event at that position.  Column c holds transcript /
 *   genome position col_pos[col_off[f] + c]; positions are strictly increasing
 *   inside a feature but need not be consecutive (columns without any event
 *   may be dropped); a window is valid only over w consecutive positions.
 *   Reads are sorted by sample: sample s owns
 *   rows [samp_start[f*(R+1)+s], samp_start[f*(R+1)+s+1]); samples
 *   0..n_cntrl-1 are the control replicates (`-c` files in order), the rest
 *   the treatment replicates (`-t` files in order).
 */
#ifndef YANOCOMP_B200_H
#define YANOCOMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YC_MAX_WINDOW   9      /* window sizes 1,3,5,7,9 */
#define YC_MAX_SAMPLES  16     /* replicates over both conditions */
#define YC_NCOMP        3      /* two Gaussians + uniform outlier component */

#define YC_COV_FULL        0   /* MultivariateGaussianDistribution                  stats.py:148-152 */
#define YC_COV_DIAG        1   /* IndependentComponentsDistribution of Normals      stats.py:153-166 */
#define YC_OUTLIER_MAD     0   /* distance-to-centroid MAD rule                     stats.py:116-120 */
#define YC_OUTLIER_DBSCAN  1   /* sklearn DBSCAN(eps, min_samples = n/4) noise      stats.py:91-104 */

/* yc_batch_t.flags: every ld[f] and x_off[f] is a multiple of 8 elements and the padding
 * elements events[x_off + c*ld + r], n_reads <= r < ld, are NaN -> rows can be read 16 bytes at a time */
#define YC_BATCH_ROWS_ALIGNED 1

/* error codes (negative; positive values are cudaError_t) */
#define YC_OK            0
#define YC_EINVAL       -1     /* bad argument (see yc_last_error) */
#define YC_ECAPACITY    -2     /* a position needs more shared memory than an SM has */
#define YC_ENOGPU       -3     /* no sm_100 device */

/* per-position status, written by yc_ks_batch / yc_gmm_*_batch */
#define YC_ST_NOT_A_WINDOW 0   /* fails the depth / window filters   gmmtest.py:24-61 */
#define YC_ST_KS_FILTERED  1   /* entered position_stats, failed the gate  stats.py:465 */
#define YC_ST_TESTED       2   /* fitted and tested (was_tested = True)    stats.py:511 */
#define YC_ST_FIT_FAILED   3   /* the LinAlgError path                     stats.py:478-479 */
#define YC_ST_DEGENERATE   4   /* constant column: np.linalg.svd would raise, stats.py:26 */
#define YC_ST_KS_ONLY      5   /* --no-gmm                                 stats.py:460-462 */
#define YC_ST_GATED        6   /* transient: passed the gate, waiting for the fit */

typedef struct yc_batch {
    int32_t  n_features;         /* F */
    int32_t  n_samples;          /* R  (<= YC_MAX_SAMPLES) */
    int32_t  n_cntrl;            /* Rc: samples [0,Rc) are control */
    int32_t  max_reads;          /* max over features of n_reads (sizes shared memory) */
    int32_t  max_cond_reads;     /* max over features of max(control rows, treatment rows) */
    int32_t  flags;              /* YC_BATCH_* */
    int64_t  total_cols;         /* sum of n_cols = number of candidate positions */
    const uint16_t *events;      /* IEEE binary16 bits, layout above */
    const int64_t  *x_off;       /* [F]   element offset of each feature matrix */
    const int32_t  *n_cols;      /* [F] */
    const int32_t  *n_reads;     /* [F] */
    const int32_t  *ld;          /* [F]   row stride in elements (>= n_reads) */
    const int64_t  *col_off;     /* [F+1] prefix sum of n_cols */
    const int32_t  *samp_start;  /* [F*(R+1)] */
    const uint64_t *feature_key; /* [F]   FNV-1a 64 of the feature id (RNG key) */
    const int64_t  *col_pos;     /* [total_cols] position of each global column */
    const uint8_t  *reverse;     /* [F]   1 = '-' strand: window columns reversed, gmmtest.py:37-38 */
    const int32_t  *col_feature; /* [total_cols] feature of each global column (yc_fill_col_feature) */
} yc_batch_t;

typedef struct yc_params {
    int32_t  window_size;        /* -w, gmmtest.py:283-285 (odd, <= YC_MAX_WINDOW) */
    int32_t  min_read_depth;     /* -n, gmmtest.py:245-251, :301-303 */
    int32_t  max_fit_depth;      /* -d, gmmtest.py:304-305 */
    int32_t  gmm;                /* --gmm/--no-gmm, gmmtest.py:310-311 */
    int32_t  add_uniform;        /* -u/--add-uniform / --no-uniform, gmmtest.py:290-292 (1 = 3-component GMM) */
    int32_t  covariance_type;    /* -C, gmmtest.py:286-289: YC_COV_FULL / YC_COV_DIAG */
    double   min_ks;             /* -k, gmmtest.py:306-307 */
    double   fdr_threshold;      /* -f, gmmtest.py:308-309 */
    double   outlier_factor;     /* -e, gmmtest.py:296-298 */
    uint64_t seed;               /* --random-seed, gmmtest.py:315 */
    double   dbscan_eps;         /* -E, gmmtest.py:299-300 */
    int32_t  outlier_method;     /* -D, gmmtest.py:293-295: YC_OUTLIER_MAD / YC_OUTLIER_DBSCAN */
    int32_t  reserved1;
} yc_params_t;

/* fitted model + test results of one gated position (slot-indexed SoA) */
typedef struct yc_fit_out {
    double  *mu;        /* [n_fit*2*w]     component means            stats.py:277-280 */
    double  *cov;       /* [n_fit*2*w*w]   component covariances      stats.py:387-392 */
    double  *box;       /* [n_fit*w*2]     uniform (lo,hi) per column stats.py:183-186 */
    double  *logw;      /* [n_fit*R*3]     per-sample log weights     stats.py:207-211 */
    double  *cw;        /* [n_fit*3]       combined weights           stats.py:251-253 */
    int32_t *iters;     /* [n_fit]         E-steps performed          stats.py:200-222 */
    double  *loglik;    /* [n_fit]         last log-likelihood sum */
    /* yc_gtest_smpred_batch */
    double  *preds;     /* [n_fit*R*3]     np.round(weights*sizes)    stats.py:296 */
    double  *stats;     /* [n_fit*8]       g, hom_g, p, cntrl_frac, treat_frac,
                                           log_odds, lower_ci, upper_ci  stats.py:316-355;
                                           NaN for slots that were not tested */
    int64_t *post_off;  /* [n_fit]         offset (in reads) into `post`, -1 = none */
} yc_fit_out_t;

const char *yc_version(void);
const char *yc_last_error(void);
/* number of kernels this library has launched in this process (monotonic; callers take differences) */
uint64_t    yc_kernel_launches(void);

/* how many bytes of dynamic shared memory the kernels need for this batch
 * (so a caller can reject a batch up front); returns YC_ECAPACITY if any
 * kernel cannot run. out[0..3] = ks, init, em, gtest. */
int yc_smem_bytes(const yc_batch_t *batch_host_copy, const yc_params_t *p, int64_t out[4]);

/* col_feature[g] = f for g in [col_off[f], col_off[f+1])  (helper, one launch) */
int yc_fill_col_feature(int32_t n_features, const int64_t *col_off,
                        int32_t *col_feature, void *stream);

/*
 * Ingest -- the cross-sample half of the reference's pivot (load_gene_events, yanocomp/io.py:296-343:
 * `unstack` per file, `concat` over the files of a condition) on the device.  The per-FILE half
 * (keep-first de-duplication, dwell filter, reads in read-id order) is done once, when a sample file is
 * converted to the dense container (yanocomp_b200/dense.py); its [n_pos][ld] binary16 blocks are uploaded
 * as they are (`raw`) and copied here to rows [row0, row0 + n_reads) of their feature, column by column.
 *   tasks [n_tasks * 8] int64, one row per (feature, sample) block:
 *       {feature, offset of the block in raw (elements), 0, block ld, n_pos, n_reads, row0,
 *        offset into colmap of the block's column map or -1 (block positions == feature positions)}
 *   colmap  int32: feature column of every block position (for the blocks that need one)
 *   x_off, ld  per feature, as in yc_batch_t;  events [n_event_elems] receives the batch's event buffer
 *       (NaN = no event; 16-byte aligned);  max_n_pos = the largest n_pos over the tasks.
 */
int yc_assemble_batch(const uint16_t *raw, const int64_t *tasks, int32_t n_tasks, int32_t max_n_pos,
                      const int32_t *colmap, const int64_t *x_off, const int32_t *ld,
                      uint16_t *events, int64_t n_event_elems, void *stream);

/*
 * Stage 1 -- window filters + PCA + two-sample KS over every candidate position
 * (three kernels: moments, LAPACK axis replay one thread per position, sort + D;
 * one warp or one CTA per position; then an ordered
 * compaction of the gated positions).  Scratch comes stream-ordered from the
 * device's memory pool and is returned before the call ends.  Replaces, per position: get_valid_pos / get_valid_windows /
 * index_pos_range / test_depth (gmmtest.py:24-61), subsample (stats.py:52-59),
 * pca_comp1 (stats.py:20-31), _ks_d / ks_2samp / pca_kstest (stats.py:34-73)
 * and the gate of position_stats (stats.py:460-465).
 *   ks_stat, ks_pval [total_cols]; n_c, n_t [total_cols] = complete-window rows
 *   per condition (before sub-sampling); status [total_cols];
 *   fit_list [total_cols] receives the global column index of every position
 *   that passed the gate, in ascending order (an ordered compaction pass, so
 *   the slot order of the fit outputs is deterministic), *fit_count their
 *   number (device int32).
 */
int yc_ks_batch(const yc_batch_t *b, const yc_params_t *p,
                double *ks_stat, double *ks_pval, int32_t *n_c, int32_t *n_t,
                uint8_t *status, int32_t *fit_list, int32_t *fit_count,
                void *stream);

/*
 * Ordered compaction of the gated columns of `status` ([total_cols], as yc_ks_batch leaves it) into fit_list (ascending
 * column index) and the device counter fit_count -- what yc_ks_batch does at its end for its own batch.  For callers
 * that run yc_ks_batch over several feature RANGES of one resident batch while its events are still arriving (each call
 * on a sub-batch whose pointers are offsets into the same arrays, outputs at the range's column offset) and then fit the
 * whole batch with ONE submission: gmmtest.py:156-178 has no such step, the loop is per position there.
 */
int yc_compact_gated(const uint8_t *status, int64_t total_cols, int32_t *fit_list, int32_t *fit_count, void *stream);

/*
 * The fit stages take the gated positions as (fit_list, n_fit, fit_count):
 * slots [0, min(n_fit, *fit_count)) are processed.  `fit_count` is the DEVICE
 * counter yc_ks_batch wrote, so a caller can queue the three stages right behind
 * yc_ks_batch without waiting for the count on the host: n_fit is then only an
 * upper bound (the capacity of the slot-indexed outputs, at most total_cols).
 * fit_count = NULL means "exactly n_fit slots".  The small-window kernels are
 * persistent (every warp of the grid fetches the next slot from a device
 * counter), so an upper bound costs nothing; the general kernels launch n_fit
 * CTAs of which the surplus exits at once.
 *
 * Stage 2a -- model initialisation, one warp or CTA per gated position.  Replaces
 * the per-replicate subsample of fit_gmm (stats.py:266-269),
 * kmeans_init_clusters (stats.py:81-124: pm.Kmeans k-means++ x3, MAD or DBSCAN
 * outlier labelling) and initialise_gmms (stats.py:127-191, full or diagonal
 * covariance).  Writes mu, cov, box,
 * logw; status[g] becomes YC_ST_FIT_FAILED on a failed Cholesky / zero-width
 * box.  labels (optional, may be NULL): [n_fit*max_fit_rows] int8 initial labels
 * in pooled fit order, for parity tests.
 */
int yc_gmm_init_batch(const yc_batch_t *b, const yc_params_t *p,
                      const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count,
                      yc_fit_out_t *out, uint8_t *status,
                      int8_t *labels, int32_t max_fit_rows, void *stream);

/*
 * Stage 2b -- multi-sample EM, one warp or CTA per gated position.  Replaces em
 * (stats.py:194-225) and the pomegranate calls under it (from_summaries,
 * GeneralMixtureModel.summarize) plus the weight bookkeeping of
 * fit_multisamp_gmm (stats.py:250-253).  Updates mu, cov, logw in place,
 * writes cw, iters, loglik; status[g] = YC_ST_TESTED or YC_ST_FIT_FAILED.
 * It also leaves every sample's complete-window row count in preds[.][s][0]
 * for yc_gtest_smpred_batch, which replaces it with the rounded counts.
 */
int yc_gmm_em_batch(const yc_batch_t *b, const yc_params_t *p,
                    const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count,
                    yc_fit_out_t *out, uint8_t *status, void *stream);

/*
 * Stage 3 -- per-sample modified-fraction counts, G-tests, log odds and the
 * single-molecule posteriors, one warp or CTA per gated position; call it after
 * yc_gmm_em_batch on the same outputs.  Replaces the
 * rounding of fit_gmm (stats.py:296-297), gmm_g_test (stats.py:316-338),
 * calculate_fractional_stats (stats.py:341-355) and predict_proba
 * (stats.py:494-499).  When `post` is not NULL, positions with p < fdr
 * threshold append 2 doubles per complete-window read (P(dist2), P(outlier),
 * pooled read order) at post[2*post_off[slot]...]; *post_cursor (device
 * int64, zeroed by the call) counts reads written; post_capacity is in reads.
 */
int yc_gtest_smpred_batch(const yc_batch_t *b, const yc_params_t *p,
                          const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count,
                          yc_fit_out_t *out, const uint8_t *status,
                          double *post, int64_t post_capacity,
                          int64_t *post_cursor, void *stream);

/*
 * Fixed-width result rows in slot order (one launch): everything the host
 * tail (GMMTestResults, stats.py:423-442; the records of gmmtest.py:164-172)
 * reads per gated position as one row of YC_ROW_WIDTH(w) doubles, so that a
 * shard's results are ONE contiguous device->host copy and one fixed-width
 * block in the multi-GPU gather.  Row layout:
 *   [YC_ROW_STATUS]   final status (YC_ST_TESTED / YC_ST_FIT_FAILED)
 *   [YC_ROW_KS_STAT]  KS D (GMMTestResults.ks_stat)
 *   [YC_ROW_KS_PVAL]  device KS p-value: serves the gate only; the host
 *                     re-decides values within YC_KS_P_GUARD of the threshold
 *                     with SciPy (stats.py:465) from the two counts below
 *   [YC_ROW_N_CNTRL]  complete-window control rows, [YC_ROW_N_ALL] all rows
 *   [YC_ROW_COLUMN]   global column index (= fit_list[slot])
 *   [YC_ROW_STATS..+8] g, hom_g, p, cntrl_frac, treat_frac, log_odds, lower, upper
 *   [YC_ROW_MU..+2w]   dist1_means, dist2_means          (stats.py:277-280)
 *   [YC_ROW_MU+2w..+2w] dist1_stds, dist2_stds = sqrt(diag(cov))
 * rows [n_fit * YC_ROW_WIDTH(w)]; slots [0, min(n_fit, *fit_count)) are written.
 */
#define YC_ROW_STATUS   0
#define YC_ROW_KS_STAT  1
#define YC_ROW_KS_PVAL  2
#define YC_ROW_N_CNTRL  3
#define YC_ROW_N_ALL    4
#define YC_ROW_COLUMN   5
#define YC_ROW_STATS    6
#define YC_ROW_MU       14
#define YC_ROW_WIDTH(w) (14 + 4 * (w))
int yc_gather_fit_rows(const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count, int32_t window_size,
                       const uint8_t *status, const double *ks_stat, const double *ks_pval,
                       const int32_t *n_c, const int32_t *n_t, const yc_fit_out_t *out,
                       double *rows, void *stream);

/*
 * Stages 2a + 2b + 3 back to back on one stream (convenience; same outputs
 * as the three calls above).
 *
 * Each of the three stages has two implementations chosen per batch: for
 * small windows (w <= 5, <= 1024 reads per feature, no sub-sampling possible)
 * one WARP per gated position (csrc/yc_fit_warp.cu), otherwise one 128-thread
 * CTA per gated position (csrc/yc_fit.cu).  YC_FIT_PATH=cta / YC_KS_PATH=cta in
 * the environment force the general kernels.
 */
int yc_gmm_fit_batch(const yc_batch_t *b, const yc_params_t *p,
                     const int32_t *fit_list, int32_t n_fit, const int32_t *fit_count,
                     yc_fit_out_t *out, uint8_t *status,
                     int8_t *labels, int32_t max_fit_rows,
                     double *post, int64_t post_capacity,
                     int64_t *post_cursor, void *stream);

/*
 * Host-side scalar entry points (no GPU needed): the same code the kernels
 * run for scipy.stats.kstwobign.sf (stats.py:48), chi2.sf at 1 dof and the
 * G-test / fraction block (stats.py:302-355), exported so CPU tests can pin
 * them against SciPy; plus the RNG key functions shared with the oracle.
 */
double   yc_kolmogorov_sf_host(double x);
double   yc_chi2_sf_1dof_host(double g);
void     yc_gtest_fractions_host(const double *preds, int32_t nrow_c, int32_t nrow_t,
                                 double thr, double *out8);
/* the same for the two-component model of --no-uniform: the third column of preds is ignored */
void     yc_gtest_fractions2_host(const double *preds, int32_t nrow_c, int32_t nrow_t,
                                  double thr, double *out8);
uint64_t yc_position_key_host(uint64_t seed, uint64_t fkey, int64_t pos);
uint64_t yc_mix64_host(uint64_t z);
/* first right singular vector of Z with numpy.linalg.svd's (LAPACK dgesdd) sign,
 * from G = Z^T Z (d x d) and T = Z[0:d, :] (d x d), both row-major; stats.py:26-29 */
int      yc_first_right_singular_vector_host(int32_t d, const double *G, const double *T, double *v);
/* the same routine run on the device for `count` independent problems (device pointers) */
int      yc_first_right_singular_vector_device(int32_t d, const double *G, const double *T, double *v,
                                               int32_t count, void *stream);

/* test hook: the KS kernels form c / n (stats.py:39-41: counts over sample sizes) as a reciprocal multiplication with two
 * residual corrections instead of an IEEE division; this checks that routine against the division on the device for every
 * pair 0 <= c <= n <= n_max and for n_random pseudo-random pairs with n < 2^31, and writes the number of differing
 * quotients to *mismatches (device pointer) */
int      yc_div_int_rn_check_device(int32_t n_max, int64_t n_random, uint64_t seed, uint64_t *mismatches, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* YANOCOMP_B200_H */
