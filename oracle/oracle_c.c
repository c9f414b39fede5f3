/* CPU oracle, C backend.  TEST INFRASTRUCTURE ONLY — never linked into the product.
 *
 * All-pairs restatement of the two scipy.spatial.cKDTree calls the reference makes on its
 * k-NN path (/root/reference/ennemi/_entropy_estimators.py:39,108-110,142,152-154,194-196,
 * 240-245); semantics per SURVEY.md Appendix A.  Points are row-major (n, d) fp64.
 * Parity status: pinned — tests/test_oracle_golden.py checks this backend bit for bit
 * against fixtures produced by the unmodified reference (oracle/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* max-norm distance, one rounded subtraction per dimension */
static inline double cheb(const double *a, const double *b, int d) {
  double m = 0.0;
  for (int t = 0; t < d; ++t) {
    double v = fabs(a[t] - b[t]);
    if (v > m) m = v;
  }
  return m;
}

/* (k+1)-th smallest distance from each query to the candidate set (self is just another
 * candidate at distance 0); inf when there are fewer than k+1 candidates. */
int orc_kth_distance(const double *cand, int64_t n_cand, const double *query, int64_t n_query, int d, int k,
                     double *out) {
  const int k1 = k + 1;
  int fail = 0;
#pragma omp parallel
  {
    double *best = (double *)malloc(sizeof(double) * (size_t)k1);
    if (!best) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(static)
      for (int64_t i = 0; i < n_query; ++i) {
        for (int t = 0; t < k1; ++t) best[t] = INFINITY;      /* sorted ascending */
        const double *q = query + i * d;
        for (int64_t j = 0; j < n_cand; ++j) {
          double v = cheb(q, cand + j * d, d);
          if (v < best[k1 - 1]) {
            int t = k1 - 1;
            while (t > 0 && best[t - 1] > v) { best[t] = best[t - 1]; --t; }
            best[t] = v;
          }
        }
        out[i] = best[k1 - 1];
      }
      free(best);
    }
  }
  return fail;
}

/* #{j : dist(query_i, cand_j) <= radius_i}, inclusive */
int orc_ball_count(const double *cand, int64_t n_cand, const double *query, int64_t n_query, int d,
                   const double *radius, int64_t *out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_query; ++i) {
    const double *q = query + i * d;
    const double r = radius[i];
    int64_t c = 0;
    for (int64_t j = 0; j < n_cand; ++j) c += (cheb(q, cand + j * d, d) <= r);
    out[i] = c;
  }
  return 0;
}
