/* ennemi_b200 — C ABI of the B200 (sm_100a) k-NN mutual-information library.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference (polsys/ennemi 1.5.0)
 * has no FFI of its own: the seam is the five private Python estimators imported at
 * ennemi/_driver.py:18-21 and called at _driver.py:222-226 and :815-832.  Each eb2_* estimator
 * below replaces one of them; the arithmetic it replaces is the scipy.spatial.cKDTree calls at
 * ennemi/_entropy_estimators.py:39,108-110,142,152-154,194-196,240-245 and _psi at :327-350.
 *
 * Conventions
 *  - plain pointers and sizes only; every array is contiguous fp64 / int32 / int64.
 *  - `coords` is DIMENSION-MAJOR: d blocks of n doubles, block t = coordinate t of every
 *    observation (the host transposes the reference's (n, d) row-major arrays once).
 *  - buffers are HOST pointers owned by the caller for the duration of the call, unless
 *    EB2_FLAG_DEVICE_INPUT is set in `flags`, in which case `coords`/`cls` are device pointers on
 *    device `dev` (used to measure the kernels with inputs already resident in HBM).
 *  - optional outputs (eps_out, count outputs) may be NULL; when given they are host arrays of
 *    length n in the caller's original row order.
 *  - `partial` is an EB2_P_LEN-double block of raw sums for the rows [row_lo, row_hi) (see EB2_P_*),
 *    so that query rows can be sharded over GPUs/ranks and combined with one sum-allreduce;
 *    eb2_*_finish turns the (summed) block into the estimate.
 *  - return value: 0 on success, an EB2_ERR_* code otherwise; eb2_last_error() gives the
 *    thread-local message.  There is no CPU fallback: without a usable CUDA device every
 *    estimator returns EB2_ERR_CUDA.
 *  - `dev` = CUDA device ordinal in bits 0-7, optionally a stream "lane" (0..3) in bits 8-15:
 *    each (device, lane) has its own stream, events and staging buffers, so small independent
 *    tasks issued from different host threads run concurrently on one GPU.
 *  - thread safety: all entry points may be called concurrently; calls on the same (device, lane)
 *    serialise on that lane's stream and workspace.
 *  - repeated eb2_ksg_mi_rows / eb2_ksg_mi calls with EB2_FLAG_DEVICE_INPUT and the same arguments are captured into a
 *    CUDA graph on their third occurrence and replayed afterwards (results are bit-identical; the buffer is read
 *    anew on every call).  Environment variable EB2_GRAPH=0 disables this.
 */
#ifndef ENNEMI_B200_H
#define ENNEMI_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define EB2_API __attribute__((visibility("default")))
#else
#define EB2_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EB2_OK 0
#define EB2_ERR_CUDA 1        /* CUDA runtime error / no device */
#define EB2_ERR_ARG 2         /* bad argument (n, k, d, pointers) */
#define EB2_ERR_NONFINITE 3   /* non-finite coordinate: mirrors cKDTree's ValueError */
#define EB2_ERR_UNSUPPORTED 4 /* dimension above EB2_MAX_DIM */
#define EB2_ERR_CONSTANT 5    /* EB2_FLAG_DEVICE_STATS: a window turned out constant; nothing was estimated */

#define EB2_MAX_DIM 32           /* up to 12 dimensions: specialised kernels; 13..32: generic brute-force kernels */

/* flags */
#define EB2_FLAG_DEVICE_INPUT 1u /* coords / cls are device pointers on `dev` */
#define EB2_FLAG_BRUTE_COUNT 2u  /* count 1-D marginals with the tiled all-pairs kernel instead of sort+search */
#define EB2_FLAG_NO_PRUNE 4u     /* visit every candidate tile (pure brute force); default is exact sorted-window pruning */
#define EB2_FLAG_DEVICE_STATS 16u /* *_cols calls (implies SINGLE_USE: no prepared variables are cached): a descriptor with
                                    mean = NaN and std != 0 gets its window's mean and std computed on the device inside the
                                    call (NumPy's pairwise association, as eb2_cache_stats) - no host round trip before the
                                    task.  A window that turns out constant (std < 1e-20, ennemi/_driver.py:879) fails the
                                    call with EB2_ERR_CONSTANT so that the caller can take the reference's warning path */
#define EB2_FLAG_SINGLE_USE 8u   /* *_cols calls: the descriptors are used by this task only - do not cache prepared variables */

/* layout of the 8-double partial block */
#define EB2_P_SUM 0   /* sum over rows of the per-row term (psi combination, or log(dist)) */
#define EB2_P_ZERO_A 1 /* number of rows whose first  count is 0 (psi(0) = inf in the reference) */
#define EB2_P_ZERO_B 2 /* ... second count */
#define EB2_P_ZERO_C 3 /* ... third count */
#define EB2_P_ROWS 4  /* rows reduced */
#define EB2_P_PAIRS 5 /* point pairs evaluated by the all-pairs kernels (work counter for the roofline) */
#define EB2_P_FIX0 8  /* [8..11]: the bivariate pipeline's digamma sum as an exact 128-bit fixed-point integer (units of
                         2^-48), four 32-bit limbs, least significant first, the last one signed: integer-valued doubles
                         add exactly, so a sum of partial blocks over shards / ranks is independent of their number */
#define EB2_P_FIXED 12 /* > 0 when the limbs are present (then they, not EB2_P_SUM, define the sum) */
#define EB2_P_LEN 16

EB2_API int eb2_init(void);            /* optional; creates the per-device contexts eagerly */
EB2_API int eb2_shutdown(void);        /* frees every device workspace */
EB2_API int eb2_device_count(void);    /* usable CUDA devices (0 when there is none) */
EB2_API const char* eb2_last_error(void);
EB2_API const char* eb2_version(void);

/* a1  _estimate_single_mi(x, y, k)            _entropy_estimators.py:69-113
 * coords = [x ; y] (2 x n).  value = psi(n) + psi(k) - mean(psi(nx) + psi(ny)). */
EB2_API int eb2_ksg_mi(int dev, const double* coords, int64_t n, int k, uint32_t flags,
               double* value, double* eps_out, int64_t* nx_out, int64_t* ny_out);
EB2_API int eb2_ksg_mi_rows(int dev, const double* coords, int64_t n, int k, uint32_t flags,
                    int64_t row_lo, int64_t row_hi, double* partial,
                    double* eps_out, int64_t* nx_out, int64_t* ny_out);
EB2_API int eb2_ksg_mi_finish(const double* partial, int64_t n, int k, double* value);

/* a2  _estimate_conditional_mi(x, y, cond, k)  _entropy_estimators.py:116-156
 * coords = [x ; y ; z_0 .. z_{c-1}] ((2+c) x n).
 * value = psi(k) - mean(psi(nxz) + psi(nyz) - psi(nz)). */
EB2_API int eb2_cmi(int dev, const double* coords, int64_t n, int c, int k, uint32_t flags,
            double* value, double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out);
EB2_API int eb2_cmi_rows(int dev, const double* coords, int64_t n, int c, int k, uint32_t flags,
                 int64_t row_lo, int64_t row_hi, double* partial,
                 double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out);
EB2_API int eb2_cmi_finish(const double* partial, int64_t n, int k, double* value);

/* a3  _estimate_semidiscrete_mi(x, y, k)       _entropy_estimators.py:159-200
 * coords = [x] (1 x n); cls[i] in [0, ncls) = index of y_i in np.unique(y).
 * value = psi(n) + psi(k) - mean(psi(n_full)) - sum_c psi(n_c) n_c / n. */
EB2_API int eb2_ross_mi(int dev, const double* coords, const int32_t* cls, int64_t n, int ncls, int k, uint32_t flags,
                double* value, double* eps_out, int64_t* nfull_out);

/* a4  _estimate_conditional_semidiscrete_mi(x, y, cond, k)   _entropy_estimators.py:203-247
 * coords = [x ; z_0 .. z_{c-1}] ((1+c) x n); cls as above.
 * value = psi(k) - mean(psi(nxz) + psi(nyz) - psi(nz)). */
EB2_API int eb2_ross_cmi(int dev, const double* coords, const int32_t* cls, int64_t n, int c, int ncls, int k,
                 uint32_t flags, double* value, double* eps_out,
                 int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out);

/* a5  _estimate_single_entropy(x, k)           _entropy_estimators.py:21-42
 * coords = [x_0 .. x_{m-1}] (m x n).
 * value = psi(n) - psi(k) + m (mean(log dist) + log 2). */
EB2_API int eb2_entropy(int dev, const double* coords, int64_t n, int m, int k, uint32_t flags,
                double* value, double* dist_out);
EB2_API int eb2_entropy_rows(int dev, const double* coords, int64_t n, int m, int k, uint32_t flags,
                     int64_t row_lo, int64_t row_hi, double* partial, double* dist_out);
EB2_API int eb2_entropy_finish(const double* partial, int64_t n, int m, int k, double* value);

/* a6  _psi(x)                                  _entropy_estimators.py:327-350
 * digamma of non-negative integers evaluated ON THE DEVICE with the reference's expansion;
 * out[i] = +inf where counts[i] == 0. */
EB2_API int eb2_psi(int dev, const int64_t* counts, int64_t n, double* out);

/* the two primitives on their own (parity tests, building blocks):
 * kth:   out[i] = (k+1)-th smallest Chebyshev distance from row i to all rows (self included);
 *        with cls != NULL only rows of the same class are candidates.
 * count: out[i] = #{j : max_t |a_it - a_jt| <= radius_i}; with cls != NULL and within_class != 0
 *        only rows of the same class are candidates. */
EB2_API int eb2_kth_distance(int dev, const double* coords, const int32_t* cls, int64_t n, int d, int ncls, int k,
                     uint32_t flags, double* out);
EB2_API int eb2_ball_count(int dev, const double* coords, const int32_t* cls, int64_t n, int d, int ncls,
                   int within_class, const double* radius, uint32_t flags, int64_t* out);

/* timing of the last estimator call on `dev`, from CUDA events on the library's stream:
 * ms[0] total device span (first H2D to last D2H), ms[1] k-NN kernel, ms[2] marginal counting,
 * ms[3] digamma/reduction, ms[4] sort/permutation.  launches = kernels launched by that call. */
EB2_API int eb2_last_timing(int dev, double* ms, int* launches);
/* which implementation the last estimator call on `dev` ran on: 1 = the bivariate pipeline (eb2_ksg_mi*), 2 = the
 * three-level grid (eb2_entropy* in 3 and 4 dimensions from 200,000 rows on; eb2_cmi* only with EB2_G3_CMI set), 0 = the
 * general path (size / k / flags outside the grids, or a bucket overflow on heavily tied data), -1 without a context */
EB2_API int eb2_last_pipeline(int dev);

/* ---- device-resident columns (lag sweeps and pairwise_mi upload every
 * variable once instead of once per task; the reference's per-task _rescale_data,
 * ennemi/_driver.py:871-902, runs on the device with host-computed mean/std and the host's
 * fixed-seed noise vector, bit-identically) ------------------------------------------------------- */
EB2_API int eb2_cache_put(int dev, uint64_t key, const double* host, int64_t n);  /* key != 0; replaces */
EB2_API int eb2_cache_drop(int dev, uint64_t key);                               /* key == 0: everything */
/* ncols columns at once from a row-major (n x ncols) host block with row stride `ld` elements (ld >= ncols): ONE
 * host-to-device copy, then a de-interleave kernel on the device.  Column j of the block is cached under keys[j].
 * Replaces the per-column strided gathers the host would otherwise do for the (n, nvar) arrays the reference's
 * pairwise_mi / estimate_mi take (ennemi/_driver.py:477-483, 703-707 slice them column by column). */
EB2_API int eb2_cache_put_block(int dev, const uint64_t* keys, int ncols, const double* host, int64_t n, int64_t ld);
/* ... from a block that already is in DEVICE memory on `dev` (multi-GPU jobs: every rank uploads 1/G of the rows and
 * the slices are all-gathered over NVLink before this call) */
EB2_API int eb2_cache_put_block_dev(int dev, const uint64_t* keys, int ncols, const double* dev_block, int64_t n, int64_t ld);

/* mean and standard deviation (ddof = 0) of the window column[key][off + i*stride], i in [0, n), computed on the
 * device with NumPy's pairwise-summation association, i.e. the same bits as ndarray.mean() / ndarray.std()
 * (the values _rescale_data uses, ennemi/_driver.py:878-882). */
EB2_API int eb2_cache_stats(int dev, uint64_t key, int64_t off, int64_t stride, int64_t n, double* mean, double* std);
/* ... for nwin windows of one length in one call: window w = column[keys[w]][offs[w] + i*stride] */
EB2_API int eb2_cache_stats_many(int dev, const uint64_t* keys, const int64_t* offs, int nwin, int64_t stride, int64_t n,
                                 double* means, double* stds);

/* coordinate t of the joint space, for i in [0, n):
 *   v_i = column[key][off + i * stride]
 *   if (std != 0)  v_i = (v_i - mean) / std  and, if nkey != 0,  v_i += column[nkey][noff + i * nstride] */
typedef struct eb2_col {
  uint64_t key;
  int64_t off, stride;
  double mean, std;
  uint64_t nkey;
  int64_t noff, nstride;
} eb2_col_t;

/* a1 / a2 on cached columns: cols[0] = x, cols[1] = y, cols[2..] = condition.  On EB2_ERR_NONFINITE,
 * eb2_last_data_flags() tells NaN input (bit0) from otherwise non-finite prepared data (bit1). */
EB2_API int eb2_ksg_mi_cols(int dev, const eb2_col_t* cols, int64_t n, int k, uint32_t flags, double* value);
EB2_API int eb2_cmi_cols(int dev, const eb2_col_t* cols, int64_t n, int c, int k, uint32_t flags, double* value);
/* k-NN entropy of m device-resident columns (a5 on cached columns): the conditional entropy
 * H(X | C) = H(X, C) - H(C) of ennemi/_driver.py:202-220 uploads X and C once and names the columns of C in both terms
 * (SURVEY.md 8 f4).  std = 0 in a descriptor passes the values through unscaled, as estimate_entropy does. */
EB2_API int eb2_entropy_cols(int dev, const eb2_col_t* cols, int64_t n, int m, int k, uint32_t flags, double* value);
/* ... and with query rows sharded (partial block as in eb2_*_rows; finish with eb2_*_finish) */
EB2_API int eb2_ksg_mi_cols_rows(int dev, const eb2_col_t* cols, int64_t n, int k, uint32_t flags,
                                 int64_t row_lo, int64_t row_hi, double* partial);
EB2_API int eb2_cmi_cols_rows(int dev, const eb2_col_t* cols, int64_t n, int c, int k, uint32_t flags,
                              int64_t row_lo, int64_t row_hi, double* partial);
EB2_API int eb2_last_data_flags(void);
/* ntasks tasks of one shape (c = 0: a1, c > 0: a2) in one call; task t uses cols[t*(2+c) .. ).  status[t] is 0
 * or EB2_ERR_* | data_flags << 8 (values[t] = NaN).  Replaces a host loop of per-task calls. */
EB2_API int eb2_mi_cols_batch(int dev, const eb2_col_t* cols, int64_t ntasks, int c, int64_t n, int k, uint32_t flags,
                              double* values, int* status);

/* BASELINE.json configs[1] from ONE process: a single estimate with its query rows sharded over the first `ngpu` GPUs
 * of the box (host buffers; every GPU receives the point set).  The partial blocks are summed on the host; their
 * digamma sums are exact integers, so the value is bit-identical for every ngpu.  The one-process-per-GPU form (NCCL
 * all-reduce of the same partial blocks) is ennemi_b200/distributed.py::sharded_ksg_mi. */
EB2_API int eb2_sharded_ksg_mi(int ngpu, const double* coords, int64_t n, int k, uint32_t flags, double* value);

/* a7 for pairwise_mi (the task list of ennemi/_driver.py:703-707 in ONE call): npairs bivariate estimates over a set
 * of nvar prepared variables; pair t is (x = cols[pairs[2t]], y = cols[pairs[2t+1]]).  Every variable is rescaled and
 * sorted once, the pairs run through the estimator in batches (one launch per stage for a whole batch).  values[t] /
 * status[t] as in eb2_mi_cols_batch. */
EB2_API int eb2_ksg_mi_pairs(int dev, const eb2_col_t* cols, int nvar, const int32_t* pairs, int64_t npairs, int64_t n, int k,
                             uint32_t flags, double* values, int* status);

/* roofline denominator: FP64 (DADD) instructions per second this device retires, in 10^12/s,
 * measured with a register-resident kernel (best of 5).  The all-pairs kernels are bound by it. */
EB2_API int eb2_measure_fp64_peak(int dev, double* tera_instr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ENNEMI_B200_H */
