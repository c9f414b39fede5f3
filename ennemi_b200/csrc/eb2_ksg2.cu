// ennemi_b200 — kernels of the bivariate KSG pipeline (sm_100a): sort-free adaptive grid.  Data layout: eb2_ksg2.h.
//
// What each stage replaces in the reference (ennemi/_entropy_estimators.py):
//   colgrid + layout   the three cKDTree builds (:100-102)
//   knn                grid.query(xy, k=[k+1], p=inf)                         (:108)
//   count_psi          x_grid / y_grid.query_ball_point(.., eps - 1e-12, p=inf, return_length=True) and the
//                      digamma terms of the mean                              (:109-110, :113, :327-350)
// Bit-exactness rules are those of eb2_kernels.cuh: one rounded fp64 subtraction per coordinate, exact
// comparisons; cell ranges are conservative (thresholds widened by 2^-50 relative) and every value in a boundary
// cell is decided by the exact predicate.
#include <algorithm>
#include <cstdlib>

#include "eb2_kernels.cuh"
#include "eb2_ksg2.h"

namespace eb2 {
namespace k2 {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSplitThreads = 512;
constexpr int kSplitItems = 8;                 // kSplitThreads * kSplitItems = 4,096 samples at most
constexpr int kHistRows = 2048;                // rows per CTA of the bucket histogram / scatter kernels
constexpr int kThreadsB = 256;                 // threads per CTA everywhere else
constexpr double kSlack = 8.881784197001252e-16;   // 2^-50

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// floor((v - lo) * sc) clamped to [0, ncell - 1]: monotone non-decreasing in v (rounded subtraction, multiplication
// by a non-negative scale, truncation and clamping all are).  The one function cells are built AND queried with.
__device__ __forceinline__ int lin_cell(double v, double lo, double sc, int ncell) {
  const double t = (v - lo) * sc;
  if (!(t > 0.0)) return 0;                    // (also 0 * inf = NaN when the map is degenerate: one cell)
  if (t >= (double)ncell) return ncell - 1;
  return (int)t;
}

// bucket of v: quantile stretch c = number of splitters below v, cut linearly into kSub parts
struct BucketFn {
  const double* split;
  const double* clo;
  const double* csc;
  int Bc;
};
__device__ __forceinline__ int bucket_fn(const BucketFn& f, double v) {
  int lo = 0, hi = f.Bc - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (f.split[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;                                   // (kSub == 1: the buckets are the quantile stretches themselves)
}

// in-place exclusive prefix sum of a[0 .. m) in shared memory, m <= 16 * blockDim.x, by the whole CTA;
// returns the total.  red: 32 ints of shared scratch.
__device__ __forceinline__ int block_excl_scan(int* a, int m, int* red) {
  const int per = (m + blockDim.x - 1) / blockDim.x;
  const int lo = min(m, (int)threadIdx.x * per), hi = min(m, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 31) red[w] = inc;
  __syncthreads();
  int base = 0, total = 0;
  for (int i = 0; i < nw; ++i) {
    if (i < w) base += red[i];
    total += red[i];
  }
  int run = base + inc - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// block-wide minimum and maximum of finite doubles (every thread gets both); red: 64 doubles of shared scratch
__device__ __forceinline__ void block_minmax(double& mn, double& mx, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(kFull, mn, o));
    mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[w] = mn; red[32 + w] = mx; }
  __syncthreads();
  mn = red[0]; mx = red[32];
  for (int i = 1; i < nw; ++i) { mn = fmin(mn, red[i]); mx = fmax(mx, red[32 + i]); }
  __syncthreads();
}

// block-wide sums of three doubles (every thread gets them); red: 64 doubles of shared scratch, at most 16 warps
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    b += __shfl_xor_sync(kFull, b, o);
    c += __shfl_xor_sync(kFull, c, o);
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[w] = a; red[16 + w] = b; red[32 + w] = c; }
  __syncthreads();
  a = red[0]; b = red[16]; c = red[32];
  for (int i = 1; i < nw; ++i) { a += red[i]; b += red[16 + i]; c += red[32 + i]; }
  __syncthreads();
}

// ---- colgrid 1: the bucket function from a sorted regular sample ------------------------------------------------
// The S <= 4,096 sample values are ranked by counting, spread over the GPU: a CTA owns 32 samples (one per lane), its
// eight warps each compare them with an eighth of all samples (every lane reads the same shared-memory word: a
// broadcast), and the sample is written to its rank.  Ties are broken by sample index, so the ranks are a permutation.
constexpr int kRankThreads = 256;
__global__ void __launch_bounds__(kRankThreads) sample_gather_kernel(const Col* cols, long long n, int S) {
  const Col c = cols[blockIdx.y];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= S) return;
  // sample g sits in the middle of the g-th of S equal stretches of the column
  double v = c.vals[(long long)(((2 * (long long)g + 1) * n) / (2 * (long long)S))];
  if (!(fabs(v) < d_inf())) v = d_inf();        // non-finite input: reported by the histogram kernel
  c.sraw[g] = v;
}

__global__ void __launch_bounds__(kRankThreads) sample_rank_kernel(const Col* cols, int S) {
  __shared__ double sk[kSplitThreads * kSplitItems];
  __shared__ int part[kRankThreads / 32][32];
  const Col c = cols[blockIdx.y];
  for (int g = threadIdx.x; g < S; g += blockDim.x) sk[g] = c.sraw[g];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + lane;
  const double v = g < S ? sk[g] : d_inf();
  const int per = (S + 7) / 8;
  const int j0 = warp * per, j1 = min(S, j0 + per);
  int cnt = 0;
  for (int j = j0; j < j1; ++j) {
    const double u = sk[j];
    cnt += (int)(u < v || (u == v && j < g));
  }
  part[warp][lane] = cnt;
  __syncthreads();
  if (warp == 0 && g < S) {
    int r = 0;
#pragma unroll
    for (int w = 0; w < kRankThreads / 32; ++w) r += part[w][lane];
    c.ssort[r] = v;
  }
}

__global__ void __launch_bounds__(kSplitThreads) split_tables_kernel(const Col* cols, int Bc, int over, int NB) {
  const Col c = cols[blockIdx.x];
  const int S = Bc * over;
  const double* sk = c.ssort;
  for (int q = threadIdx.x; q < Bc; q += blockDim.x) {
    const double lo = q == 0 ? sk[0] : sk[q * over - 1];
    const double hi = q == Bc - 1 ? sk[S - 1] : sk[(q + 1) * over - 1];
    c.clo[q] = lo;
    c.csc[q] = (hi > lo && hi - lo < d_inf()) ? (double)kSub / (hi - lo) : 0.0;
    if (q < Bc - 1) c.split[q] = hi;
    // bounds of the bucket's values for the search kernels: the values of bucket q lie in (vlo, vhi]
    c.vlo[q] = q == 0 ? -d_inf() : lo;
    c.vhi[q] = q == Bc - 1 ? d_inf() : hi;
  }
  for (int b = threadIdx.x; b <= NB; b += blockDim.x) { c.count[b] = 0; c.fill[b] = 0; }
}

// ---- colgrid 2: bucket of every row, rows per bucket ---------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsB) bucket_hist_kernel(const Col* cols, long long n, int Bc, int NB) {
  __shared__ double s_split[kMaxCoarse];
  __shared__ int s_hist[kMaxBuckets];
  const Col c = cols[blockIdx.y];
  for (int q = threadIdx.x; q < Bc; q += blockDim.x) s_split[q] = q < Bc - 1 ? c.split[q] : d_inf();
  for (int b = threadIdx.x; b < NB; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  const BucketFn fn{s_split, nullptr, nullptr, Bc};
  const long long r0 = (long long)blockIdx.x * kHistRows;
  bool bad = false;
#pragma unroll 4
  for (int i = 0; i < kHistRows / kThreadsB; ++i) {
    const long long r = r0 + i * kThreadsB + threadIdx.x;
    if (r < n) {
      const double v = c.vals[r];
      if (!(fabs(v) < d_inf())) bad = true;
      const int b = bucket_fn(fn, v);
      c.bkt[r] = (unsigned short)b;
      atomicAdd(&s_hist[b], 1);
    }
  }
  if (bad) atomicOr(c.flag, kFlagNonFinite);
  __syncthreads();
  for (int b = threadIdx.x; b < NB; b += blockDim.x)
    if (s_hist[b]) atomicAdd(&c.count[b], s_hist[b]);
}

// ---- colgrid 3: rows and values grouped by bucket ------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsB) bucket_scatter_kernel(const Col* cols, long long n, int NB) {
  __shared__ int s_boff[kMaxBuckets];
  __shared__ int s_hist[kMaxBuckets];
  __shared__ int s_base[kMaxBuckets];
  __shared__ int red[32];
  const Col c = cols[blockIdx.y];
  for (int b = threadIdx.x; b < NB; b += blockDim.x) { s_boff[b] = c.count[b]; s_hist[b] = 0; }
  __syncthreads();
  block_excl_scan(s_boff, NB, red);
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < NB; b += blockDim.x) c.boff[b] = s_boff[b];
    if (threadIdx.x == 0) c.boff[NB] = (int)n;
  }
  const long long r0 = (long long)blockIdx.x * kHistRows;
  for (int i = 0; i < kHistRows / kThreadsB; ++i) {
    const long long r = r0 + i * kThreadsB + threadIdx.x;
    if (r < n) atomicAdd(&s_hist[c.bkt[r]], 1);
  }
  __syncthreads();
  // one reservation per (CTA, bucket); the order of the rows inside a bucket is irrelevant downstream
  for (int b = threadIdx.x; b < NB; b += blockDim.x) {
    const int h = s_hist[b];
    if (h) s_base[b] = s_boff[b] + atomicAdd(&c.fill[b], h);
    s_hist[b] = 0;
  }
  __syncthreads();
  for (int i = 0; i < kHistRows / kThreadsB; ++i) {
    const long long r = r0 + i * kThreadsB + threadIdx.x;
    if (r < n) {
      const int b = c.bkt[r];
      const int pos = s_base[b] + atomicAdd(&s_hist[b], 1);
      c.srow[pos] = (int)r;
      c.sval[pos] = c.vals[r];
    }
  }
}

// ---- colgrid 4: the values of every bucket grouped into fine cells (about one per cell) ----------------------------
// The cell map of bucket b is linear over the stretch the bucket function assigns to it (split[b-1], split[b]] (the
// outermost sample values at the two ends; values beyond them clamp into the end cells).
static_assert(kSub == 1, "fine_cells_kernel takes the bucket bounds from the quantile stretches; bucket_fn returns the stretch");
__global__ void __launch_bounds__(kThreadsB) fine_cells_kernel(const Col* cols, long long n, int NB) {
  __shared__ int s_hist[kMaxCells];
  __shared__ int red[32];
  const Col c = cols[blockIdx.y];
  const int b = blockIdx.x;
  const int len = c.count[b], off = c.boff[b];
  if (b == NB - 1 && threadIdx.x == 0) { c.fstart[n] = (int)n; c.fstart[n + 1] = (int)n; }
  if (len > kBucketCap) {
    if (threadIdx.x == 0) atomicOr(c.flag, kFlagOverflow);
    return;
  }
  if (*c.flag & (kFlagNonFinite | kFlagNaN)) return;
  const double lo = c.clo[b];
  const int G = min(max(len, 1), kMaxCells);
  const double sc = c.csc[b] * (double)G;                   // csc = 1 / width of the stretch (0: degenerate)
  if (threadIdx.x == 0) {
    c.fsc[b] = sc;
    c.ncell[b] = len ? G : 0;
    c.fg[b] = FineGrid{lo, sc, off, len ? G : 0};
  }
  if (len == 0) return;
  for (int g = threadIdx.x; g < G; g += blockDim.x) s_hist[g] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) atomicAdd(&s_hist[lin_cell(c.sval[off + i], lo, sc, G)], 1);
  __syncthreads();
  block_excl_scan(s_hist, G, red);
  for (int g = threadIdx.x; g < len; g += blockDim.x) c.fstart[off + g] = g < G ? off + s_hist[g] : off + len;
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const double v = c.sval[off + i];
    const int pos = atomicAdd(&s_hist[lin_cell(v, lo, sc, G)], 1);
    c.fval[off + pos] = v;
  }
}

// ---- layout: the rows of every x-bucket grouped into cells of the bucket's own y range (about one per cell) ---------
constexpr int kLayoutThreads = 512;
constexpr int kLayoutItems = 8;       // rows per thread kept in registers: buckets of up to 4,096 rows gather y once

__global__ void __launch_bounds__(kLayoutThreads, 2) layout_kernel(const Col* cols, const Prob* probs, long long n, int NB) {
  __shared__ int s_hist[kMaxCells];
  __shared__ double redd[64];
  __shared__ int red[32];
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  const int b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0) {
    *pr.left_count = 0u;
    pr.acc[0] = 0ull; pr.acc[1] = 0ull; pr.acc[2] = 0ull; pr.acc[3] = 0ull;
  }
  if ((*cx.flag | *cy.flag) != 0) return;
  const int len = cx.count[b], off = cx.boff[b];
  if (b == NB - 1 && threadIdx.x == 0) { pr.cstart[n] = (int)n; pr.cstart[n + 1] = (int)n; }
  if (len == 0) {
    if (threadIdx.x == 0) { pr.bymin[b] = 0.0; pr.bysc[b] = 0.0; }
    return;
  }
  const bool fast = len <= kLayoutItems * kLayoutThreads;
  double yv[kLayoutItems];
  int rv[kLayoutItems], cv[kLayoutItems];
  double mn = d_inf(), mx = -d_inf();
  if (fast) {
#pragma unroll
    for (int u = 0; u < kLayoutItems; ++u) {
      const int i = threadIdx.x + u * kLayoutThreads;
      rv[u] = i < len ? cx.srow[off + i] : 0;
    }
#pragma unroll
    for (int u = 0; u < kLayoutItems; ++u) {
      const int i = threadIdx.x + u * kLayoutThreads;
      yv[u] = cy.vals[rv[u]];
      if (i < len) { mn = fmin(mn, yv[u]); mx = fmax(mx, yv[u]); }
    }
  } else {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const double y = cy.vals[cx.srow[off + i]];
      mn = fmin(mn, y);
      mx = fmax(mx, y);
    }
  }
  block_minmax(mn, mx, redd);
  const int F = min(len, kMaxCells);
  const double sc = (mx > mn && mx - mn < d_inf()) ? (double)F / (mx - mn) : 0.0;
  if (threadIdx.x == 0) { pr.bymin[b] = mn; pr.bysc[b] = sc; }
  for (int f = threadIdx.x; f < F; f += blockDim.x) s_hist[f] = 0;
  __syncthreads();
  if (fast) {
#pragma unroll
    for (int u = 0; u < kLayoutItems; ++u) {
      const int i = threadIdx.x + u * kLayoutThreads;
      cv[u] = lin_cell(yv[u], mn, sc, F);
      if (i < len) atomicAdd(&s_hist[cv[u]], 1);
    }
  } else {
    for (int i = threadIdx.x; i < len; i += blockDim.x)
      atomicAdd(&s_hist[lin_cell(cy.vals[cx.srow[off + i]], mn, sc, F)], 1);
  }
  __syncthreads();
  block_excl_scan(s_hist, F, red);
  for (int f = threadIdx.x; f < len; f += blockDim.x) pr.cstart[off + f] = f < F ? off + s_hist[f] : off + len;
  __syncthreads();
  if (fast) {
#pragma unroll
    for (int u = 0; u < kLayoutItems; ++u) {
      const int i = threadIdx.x + u * kLayoutThreads;
      if (i < len) {
        const int pos = off + atomicAdd(&s_hist[cv[u]], 1);
        pr.px[pos] = cx.sval[off + i];
        pr.py[pos] = yv[u];
        pr.prow[pos] = rv[u];
        pr.pbkt[pos] = (unsigned short)b;
      }
    }
  } else {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const int r = cx.srow[off + i];
      const double y = cy.vals[r];
      const int pos = off + atomicAdd(&s_hist[lin_cell(y, mn, sc, F)], 1);
      pr.px[pos] = cx.sval[off + i];
      pr.py[pos] = y;
      pr.prow[pos] = r;
      pr.pbkt[pos] = (unsigned short)b;
    }
  }
}

// ---- knn ---------------------------------------------------------------------------------------------------------------
constexpr int kHeavy = 64;            // windows of more slots than this are scanned by a warp (leftover kernel), not a thread

// window of x-bucket [off, off + len) whose y can lie within thr of qy: the slots of the cells the widened window touches
__device__ __forceinline__ void bucket_window(const Prob& pr, int off, int len, double ymin, double sc, double qy, double thr,
                                              int& a, int& e) {
  a = off;
  e = off + len;
  if (thr < d_inf()) {
    const int F = min(len, kMaxCells);
    const double lo_v = (qy - thr) - (fabs(qy) + thr) * kSlack;
    const double hi_v = (qy + thr) + (fabs(qy) + thr) * kSlack;
    const int f_lo = lin_cell(lo_v, ymin, sc, F), f_hi = lin_cell(hi_v, ymin, sc, F);
    a = pr.cstart[off + f_lo];
    if (f_hi + 1 < F) e = pr.cstart[off + f_hi + 1];
  }
}

// Slots [a, e) against one query, in groups of four (eight independent loads in flight, one branch per group); a group
// with a slot that passes the exact test is gone through slot by slot.  SKIP: slots [skip_a, skip_e) (the seeds) are
// jumped over.
template <int K1T, bool SKIP>
__device__ __forceinline__ void scan_range(const double* __restrict__ px, const double* __restrict__ py, int a, int e,
                                           double qx, double qy, double (&best)[K1T], double& thr, unsigned long long& np,
                                           int skip_a = 0, int skip_e = 0) {
  if (e > a) np += (unsigned long long)(e - a);
  if (SKIP && a >= skip_a && a < skip_e) a = skip_e;
#pragma unroll 1
  for (int s = a; s < e; s += 4) {
    if (SKIP && s >= skip_a && s < skip_e) s = skip_e;   // (a group that straddles the start of the seeds masks them below)
    double dx[4], dy[4];
    bool any = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = min(s + u, e - 1);           // the last group repeats its final slot; repeats are masked below
      dx[u] = fabs(qx - px[t]);
      dy[u] = fabs(qy - py[t]);
      if (SKIP && t >= skip_a && t < skip_e) dx[u] = d_inf();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) any = any || (s + u < e && dx[u] < thr && dy[u] < thr);
    if (any) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (s + u < e && dx[u] < thr && dy[u] < thr) {
          topk_insert<K1T>(best, fmax(dx[u], dy[u]));
          thr = best[K1T - 1];
        }
      }
    }
  }
}

__device__ __forceinline__ void cmp_swap(double& x, double& y) {
  const bool sw = y < x;
  const double lo = sw ? y : x, hi = sw ? x : y;
  x = lo;
  y = hi;
}

// One thread per query, every lane of a warp in the same step of the search at the same time:
//   1. eight slots around the query's own (its neighbours in y inside the bucket, itself included at distance 0) are
//      sorted by a fixed 19-comparator network: the list starts full, no branches
//   2. the rest of the home bucket's window
//   3. up to `near` non-empty buckets to the right, nearest first, while the gap in x is below the current k-th distance
//      (rounded subtraction is monotone => exact), then the same to the left
// Queries that need more buckets (sparse regions of y: their k-th distance spans many buckets of x) are handed to the
// leftover kernel, which deals the buckets of ONE query to the lanes of a warp.
template <int K1T>
__global__ void __launch_bounds__(kThreadsB) knn_kernel2(const Col* cols, const Prob* probs, long long n, int NB, int k, int near,
                                                          unsigned left_cap, const Shard sh) {
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  const int lane = threadIdx.x & 31;
  // CTAs are dispatched in index order: take the slots from both ends of x inwards (even CTAs from the front, odd ones
  // from the back) - the outermost buckets hold the sparse tails, whose queries search the widest windows, and must not
  // be the last ones to start
  const unsigned nblk = gridDim.x;
  const unsigned blk = (blockIdx.x & 1u) ? nblk - 1u - (blockIdx.x >> 1) : (blockIdx.x >> 1);
  const long long s64 = (long long)blk * kThreadsB + threadIdx.x;
  const int slot = (int)min(s64, n - 1);
  const int b = pr.pbkt[slot];
  const int off = cx.boff[b];
  const bool valid = s64 < n && off >= sh.row_lo && off < sh.row_hi;      // the shard owns whole buckets
  const double kInf = d_inf();
  const double* __restrict__ px = pr.px;
  const double* __restrict__ py = pr.py;
  const double qx = px[slot], qy = py[slot];
  const int len = cx.count[b];
  unsigned long long np = 0;
  // 1. seeds
  const int sa = max(off, min(slot - 4, off + len - 8));
  const int se = min(sa + 8, off + len);
  double d[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int t = min(sa + u, se - 1);
    const double v = fmax(fabs(qx - px[t]), fabs(qy - py[t]));
    d[u] = (valid && sa + u < se) ? v : kInf;
  }
  cmp_swap(d[0], d[1]); cmp_swap(d[2], d[3]); cmp_swap(d[4], d[5]); cmp_swap(d[6], d[7]);
  cmp_swap(d[0], d[2]); cmp_swap(d[1], d[3]); cmp_swap(d[4], d[6]); cmp_swap(d[5], d[7]);
  cmp_swap(d[1], d[2]); cmp_swap(d[5], d[6]); cmp_swap(d[0], d[4]); cmp_swap(d[3], d[7]);
  cmp_swap(d[1], d[5]); cmp_swap(d[2], d[6]);
  cmp_swap(d[1], d[4]); cmp_swap(d[3], d[6]);
  cmp_swap(d[2], d[4]); cmp_swap(d[3], d[5]);
  cmp_swap(d[3], d[4]);
  double best[K1T];
#pragma unroll
  for (int t = 0; t < K1T; ++t) best[t] = d[t];
  double thr = best[K1T - 1];
  if (valid) np += (unsigned long long)(se - sa);
  // A window of more than kHeavy slots (a query in a sparse region: its k-th distance covers a good part of a bucket) is
  // not walked by one thread: the query goes to the leftover kernel from that bucket on, where a warp scans it.
  int rstart = NB, lend = 0, skip_a = -16;
  bool heavy = false;
  // 2. the rest of the home bucket
  {
    int a = 0, e = 0;
    if (valid) bucket_window(pr, off, len, pr.bymin[b], pr.bysc[b], qy, thr, a, e);
    if (e - a > kHeavy) {
      heavy = true;
      rstart = b; lend = b; skip_a = sa;
      a = e = 0;
    }
    scan_range<K1T, true>(px, py, a, e, qx, qy, best, thr, np, sa, se);
  }
  // 3. outwards
  {
    bool more = valid && !heavy;
    int j = b + 1;
    for (int t = 0; t < near && __any_sync(kFull, more); ++t) {
      int a = 0, e = 0;
      if (more) {
        int lj = 0;
        while (j < NB && (lj = cx.count[j]) == 0) ++j;
        if (j >= NB || (cx.vlo[j] - qx) >= thr) {
          more = false;
        } else {
          bucket_window(pr, cx.boff[j], lj, pr.bymin[j], pr.bysc[j], qy, thr, a, e);
          if (e - a > kHeavy) { rstart = j; more = false; a = e = 0; }
          else ++j;
        }
      }
      scan_range<K1T, false>(px, py, a, e, qx, qy, best, thr, np);
    }
    if (more) {
      while (j < NB && cx.count[j] == 0) ++j;
      if (j < NB && !((cx.vlo[j] - qx) >= thr)) rstart = j;
    }
  }
  {
    bool more = valid && !heavy;
    int j = b - 1;
    for (int t = 0; t < near && __any_sync(kFull, more); ++t) {
      int a = 0, e = 0;
      if (more) {
        int lj = 0;
        while (j >= 0 && (lj = cx.count[j]) == 0) --j;
        if (j < 0 || (qx - cx.vhi[j]) >= thr) {
          more = false;
        } else {
          bucket_window(pr, cx.boff[j], lj, pr.bymin[j], pr.bysc[j], qy, thr, a, e);
          if (e - a > kHeavy) { lend = j + 1; more = false; a = e = 0; }
          else --j;
        }
      }
      scan_range<K1T, false>(px, py, a, e, qx, qy, best, thr, np);
    }
    if (more) {
      while (j >= 0 && cx.count[j] == 0) --j;
      if (j >= 0 && !((qx - cx.vhi[j]) >= thr)) lend = j + 1;
    }
  }
  // ---- hand the rest to the leftover kernel: one reservation per warp
  const bool want = valid && (rstart < NB || lend > 0);
  const unsigned m = __ballot_sync(kFull, want);
  int my_e = -1;
  if (m != 0) {
    const int leader = __ffs(m) - 1;
    unsigned base_e = 0;
    if (lane == leader) base_e = atomicAdd(pr.left_count, (unsigned)__popc(m));
    base_e = __shfl_sync(kFull, base_e, leader);
    if (base_e + __popc(m) <= left_cap) {
      if (want) my_e = (int)(base_e + __popc(m & ((1u << lane) - 1u)));
    } else {
      // rare (the list is full): finish both directions here, one bucket and one slot at a time.  The reservation is NOT
      // taken back (a subtraction racing with other warps' reservations would leave holes and lose entries): the counter
      // only grows, the leftover kernel clamps it at the capacity, and the part of this reservation that still lies
      // inside the list is marked void.
      if (want) {
        const unsigned idx = base_e + __popc(m & ((1u << lane) - 1u));
        if (idx < left_cap) pr.left[idx].slot = -1;
        for (int jj = rstart; jj < NB; ++jj) {
          const int lj = cx.count[jj];
          if (lj == 0) continue;
          if ((cx.vlo[jj] - qx) >= thr) break;
          int a, e;
          bucket_window(pr, cx.boff[jj], lj, pr.bymin[jj], pr.bysc[jj], qy, thr, a, e);
          for (int s = a; s < e; ++s) {
            if (s >= skip_a && s < skip_a + 8) continue;
            const double xh = px[s], yh = py[s];
            if (fabs(qx - xh) < thr && fabs(qy - yh) < thr) { topk_insert<K1T>(best, fmax(fabs(qx - xh), fabs(qy - yh))); thr = best[K1T - 1]; }
          }
        }
        for (int jj = lend - 1; jj >= 0; --jj) {
          const int lj = cx.count[jj];
          if (lj == 0) continue;
          if ((qx - cx.vhi[jj]) >= thr) break;
          int a, e;
          bucket_window(pr, cx.boff[jj], lj, pr.bymin[jj], pr.bysc[jj], qy, thr, a, e);
          for (int s = a; s < e; ++s) {
            const double xh = px[s], yh = py[s];
            if (fabs(qx - xh) < thr && fabs(qy - yh) < thr) { topk_insert<K1T>(best, fmax(fabs(qx - xh), fabs(qy - yh))); thr = best[K1T - 1]; }
          }
        }
      }
    }
  }
  if (valid) {
    double r = best[0];
#pragma unroll
    for (int t = 1; t < K1T; ++t) r = (t <= k) ? best[t] : r;       // = best[k] (a chain of selects, no indexed access)
    pr.eps[slot] = r;
    if (my_e >= 0) {
      LeftEnt le;
      le.slot = slot; le.rstart = rstart; le.lend = lend; le.skip_a = skip_a;
      pr.left[my_e] = le;
#pragma unroll
      for (int t = 0; t < K1T; ++t) pr.left_best[(long long)my_e * K1T + t] = best[t];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np) atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 4), np);
}

// ---- leftover: one warp finishes one deferred query --------------------------------------------------------------------
// Its remaining buckets are dealt to the lanes, 32 at a time: every lane works out the window of ITS bucket (cell
// arithmetic, two table loads) and walks it if it is short; long windows (sparse regions: a good part of a bucket) are
// then scanned by the whole warp, lanes striding over the slots.  Private lists per lane, gated by the k-th distance the
// search kernel had reached; one merge across the lanes at the end.
template <int K1T>
__global__ void __launch_bounds__(kThreadsB) leftover_kernel2(const Col* cols, const Prob* probs, int NB, int k, unsigned left_cap) {
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nent = min(*pr.left_count, left_cap);       // (the counter runs past the capacity when the list was full)
  const double kInf = d_inf();
  const int k1 = k + 1;
  const double* __restrict__ px = pr.px;
  const double* __restrict__ py = pr.py;
  unsigned long long np = 0;
  for (unsigned e = blockIdx.x * (kThreadsB / 32) + warp; e < nent; e += gridDim.x * (kThreadsB / 32)) {
    const LeftEnt le = pr.left[e];
    if (le.slot < 0) continue;                               // void: its query was finished by the search kernel
    const double qx = px[le.slot], qy = py[le.slot];
    const double gate = pr.left_best[(long long)e * K1T + (K1T - 1)];      // current k-th distance: upper bound of eps
    const int skip_a = le.skip_a, skip_e = le.skip_a + 8;
    double best[K1T];
#pragma unroll
    for (int t = 0; t < K1T; ++t) best[t] = kInf;
    double thr = gate;
    auto test = [&](int s) {
      if (s >= skip_a && s < skip_e) return;                               // a seed: already in the list left behind
      const double dx = fabs(qx - px[s]), dy = fabs(qy - py[s]);
      if (dx < thr && dy < thr) {
        topk_insert<K1T>(best, fmax(dx, dy));
        thr = fmin(best[K1T - 1], gate);
      }
    };
    // one round: lane's bucket j (needed or not) -> short windows per lane, long ones by the warp
    auto round = [&](int j, bool needed) {
      int a = 0, en = 0;
      if (needed) bucket_window(pr, cx.boff[j], cx.count[j], pr.bymin[j], pr.bysc[j], qy, thr, a, en);
      const bool big = en - a > kHeavy;
      if (!big) {
        for (int s = a; s < en; ++s) test(s);
        if (en > a) np += (unsigned long long)(en - a);
      }
      unsigned todo = __ballot_sync(kFull, big);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int aa = __shfl_sync(kFull, a, src), ee = __shfl_sync(kFull, en, src);
        for (int s = aa + lane; s < ee; s += 32) test(s);
        if (lane == 0) np += (unsigned long long)(ee - aa);
      }
    };
    for (int j0 = le.rstart; j0 < NB; j0 += 32) {
      const int j = j0 + lane;
      const int len = j < NB ? cx.count[j] : 0;
      const bool fail = len > 0 && (cx.vlo[j] - qx) >= gate;               // monotone: every bucket further right fails too
      round(j, len > 0 && !fail);
      if (__any_sync(kFull, fail || j >= NB)) break;
    }
    for (int j0 = le.lend - 1; j0 >= 0; j0 -= 32) {
      const int j = j0 - lane;
      const int len = j >= 0 ? cx.count[j] : 0;
      const bool fail = len > 0 && (qx - cx.vhi[j]) >= gate;
      round(j, len > 0 && !fail);
      if (__any_sync(kFull, fail || j < 0)) break;
    }
    // the list the search kernel left behind joins lane 0's (disjoint candidates), then one merge across the lanes
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < K1T; ++t) {
        const double v = pr.left_best[(long long)e * K1T + t];
        if (v < best[K1T - 1]) topk_insert<K1T>(best, v);
      }
    }
    const double fin = warp_merge_lists<K1T>(best, k1);
    if (lane == k) pr.eps[le.slot] = fin;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np) atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 4), np);
}

// ---- marginal counts + digamma terms -------------------------------------------------------------------------------------
struct GridView {           // a column's fine grid; the bucket function's tables live in shared memory
  BucketFn fn;
  const FineGrid* fg;
  const double* fval;
  const int* fstart;
};

// quantile stretch of t, found by walking from a stretch nearby (the query's own): the unique c with
// split[c - 1] < t <= split[c], i.e. the number of splitters below t, as the binary search of bucket_fn gives
__device__ __forceinline__ int stretch_near(const BucketFn& f, int c, double t) {
  while (c > 0 && !(f.split[c - 1] < t)) --c;
  while (c < f.Bc - 1 && f.split[c] < t) ++c;
  return c;
}

// index of the fine cell a threshold t falls into (monotone non-decreasing in t); an empty bucket has no cells of its
// own and yields the first cell of the next bucket
__device__ __forceinline__ int cell_of(const FineGrid& fg, double t) {
  return fg.ncell == 0 ? fg.boff : fg.boff + lin_cell(t, fg.vlo, fg.fsc, fg.ncell);
}
// ... for a widened pair of thresholds lo_t <= hi_t a few ulps apart (almost always the same bucket: one table load).
// c: a stretch near lo_t, updated to lo_t's own.
__device__ __forceinline__ void cell_index_pair(const GridView& g, int& c, double lo_t, double hi_t, int& i_lo, int& i_hi) {
  c = stretch_near(g.fn, c, lo_t);
  const FineGrid f1 = g.fg[c];                 // (kSub == 1: bucket = stretch)
  i_lo = cell_of(f1, lo_t);
  if (c < g.fn.Bc - 1 && g.fn.split[c] < hi_t) {
    const FineGrid f2 = g.fg[stretch_near(g.fn, c + 1, hi_t)];
    i_hi = cell_of(f2, hi_t);
  } else {
    i_hi = cell_of(f1, hi_t);
  }
}

// #{j : |fl(v - s_j)| <= r} over the whole column (the point itself included): cells strictly between the boundary
// cells of the two thresholds are inside whatever their values (the thresholds are widened by more than the rounding
// of the subtraction), values in boundary cells are tested exactly.  In two steps so that the table loads of the two
// marginals of a query are in flight together: count_plan (cell indices -> slot ranges), count_run (the exact tests).
struct CountPlan {
  int a1, e1, a2, e2, whole;
};
// c0: the quantile stretch of v
__device__ __forceinline__ CountPlan count_plan(const GridView& g, int c0, double v, double r) {
  CountPlan p{0, 0, 0, 0, 0};
  if (!(r >= 0.0)) return p;
  const double w = (fabs(v) + r) * kSlack;
  int c = c0, i1, i2, i3, i4;
  cell_index_pair(g, c, (v - r) - w, (v - r) + w, i1, i2);
  c = c0;
  cell_index_pair(g, c, (v + r) - w, (v + r) + w, i3, i4);
  p.a1 = g.fstart[i1];
  p.e1 = g.fstart[i2 + 1];
  const int i3b = max(i3, i2 + 1);
  if (i4 >= i3b) {
    p.a2 = g.fstart[i3b];
    p.e2 = g.fstart[i4 + 1];
    if (i3b > i2 + 1) p.whole = p.a2 - p.e1;             // whole cells between the boundaries
  }
  return p;
}
__device__ __forceinline__ int count_run(const GridView& g, const CountPlan& p, double v, double r) {
  int cnt = p.whole;
  for (int s = p.a1; s < p.e1; ++s) cnt += (int)(fabs(v - g.fval[s]) <= r);
  for (int s = p.a2; s < p.e2; ++s) cnt += (int)(fabs(v - g.fval[s]) <= r);
  return cnt;
}

__device__ __forceinline__ double psi_lookup(const double* tab, int tab_n, int c) { return c < tab_n ? tab[c] : psi_ref((double)c); }

__global__ void __launch_bounds__(kThreadsB) count_psi_kernel(const Col* cols, const Prob* probs, long long n, int Bc,
                                                               const Shard sh, const double* psi_tab, int tab_n) {
  __shared__ double s_tab[2][kMaxCoarse];
  __shared__ long long s_red[kThreadsB / 32];
  __shared__ int s_redi[3][kThreadsB / 32];
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  for (int q = threadIdx.x; q < Bc; q += blockDim.x) {
    s_tab[0][q] = q < Bc - 1 ? cx.split[q] : d_inf();
    s_tab[1][q] = q < Bc - 1 ? cy.split[q] : d_inf();
  }
  __syncthreads();
  const GridView gx{{s_tab[0], nullptr, nullptr, Bc}, cx.fg, cx.fval, cx.fstart};
  const GridView gy{{s_tab[1], nullptr, nullptr, Bc}, cy.fg, cy.fval, cy.fstart};
  const long long s64 = (long long)blockIdx.x * kThreadsB + threadIdx.x;
  const int slot = (int)min(s64, n - 1);
  const int off = cx.boff[pr.pbkt[slot]];
  const bool act = s64 < n && off >= sh.row_lo && off < sh.row_hi;
  long long q = 0;
  int zx = 0, zy = 0;
  if (act) {
    const double e = pr.eps[slot];
    const double r = e - 1e-12;                                // _entropy_estimators.py:109
    const int row = pr.prow[slot];
    const double vx = pr.px[slot], vy = pr.py[slot];
    const CountPlan plx = count_plan(gx, pr.pbkt[slot], vx, r);
    const CountPlan ply = count_plan(gy, cy.bkt[row], vy, r);
    const int nx = count_run(gx, plx, vx, r);
    const int ny = count_run(gy, ply, vy, r);
    if (pr.eps_row) pr.eps_row[row] = e;
    if (pr.nx_row) pr.nx_row[row] = nx;
    if (pr.ny_row) pr.ny_row[row] = ny;
    // a zero count makes the reference's _psi return a scalar +inf (:338-339): reported through the zero counters
    double term = 0.0;
    if (nx == 0) zx = 1; else term = psi_lookup(psi_tab, tab_n, nx);
    if (ny == 0) zy = 1; else term = term + psi_lookup(psi_tab, tab_n, ny);
    q = __double2ll_rn(term * 281474976710656.0);             // units of 2^-48: |term| < 32, so |q| < 2^53
  }
  int na = act ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q += __shfl_down_sync(kFull, q, o);
    zx += __shfl_down_sync(kFull, zx, o);
    zy += __shfl_down_sync(kFull, zy, o);
    na += __shfl_down_sync(kFull, na, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_red[w] = q; s_redi[0][w] = zx; s_redi[1][w] = zy; s_redi[2][w] = na; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tq = 0;
    int tzx = 0, tzy = 0, tna = 0;
    for (int i = 0; i < kThreadsB / 32; ++i) { tq += s_red[i]; tzx += s_redi[0][i]; tzy += s_redi[1][i]; tna += s_redi[2][i]; }
    if (tna) {
      // 128-bit signed accumulation in two words: integer addition is associative, the total is order-free
      const unsigned long long add_lo = (unsigned long long)tq;
      const unsigned long long old = atomicAdd(&pr.acc[0], add_lo);
      const unsigned long long carry = (old + add_lo < old) ? 1ull : 0ull;
      const unsigned long long add_hi = (tq < 0 ? ~0ull : 0ull) + carry;
      if (add_hi) atomicAdd(&pr.acc[1], add_hi);
      if (tzx) atomicAdd(&pr.acc[2], (unsigned long long)tzx);
      if (tzy) atomicAdd(&pr.acc[3], (unsigned long long)tzy);
      atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 6), (unsigned long long)tna);     // rows reduced
    }
  }
}

__global__ void final_kernel(const Col* cols, const Prob* probs) {
  const Prob pr = probs[blockIdx.x];
  const int fl = *cols[pr.cx].flag | *cols[pr.cy].flag;
  if (fl != 0) {
    atomicOr(reinterpret_cast<int*>(pr.out + 5), fl);
    return;
  }
  const unsigned long long lo = pr.acc[0];
  const long long hi = (long long)pr.acc[1];
  pr.out[0] = ((double)hi * 18446744073709551616.0 + (double)lo) * (1.0 / 281474976710656.0);   // informative; the host
  pr.out[1] = (double)pr.acc[2];                                                                // uses the exact words
  pr.out[2] = (double)pr.acc[3];
  pr.out[3] = 0.0;
  reinterpret_cast<unsigned long long*>(pr.out)[8] = lo;
  reinterpret_cast<long long*>(pr.out)[9] = hi;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// =====================================================================================================================
// Three-level grid for spaces of three and more dimensions
// =====================================================================================================================
constexpr int kG3Threads = 256;
constexpr int kG3Heavy = 48;          // counts: cells in a bucket's window above which a query goes to the warp kernel
                                      // (the search takes its limit as a launch argument: knn3)

// ---- layout: the rows of every bucket of coordinate 0 grouped into C1 x C2 cells over its own ranges of coordinates 1, 2
__global__ void __launch_bounds__(kG3Threads) layout3_kernel(const Col* col0, const Grid3* gp, long long n, int NB) {
  __shared__ int s_hist[kMaxCells];
  __shared__ unsigned short s_cell[kBucketCap];
  __shared__ double redd[64];
  __shared__ int red[32];
  const Grid3 g = *gp;
  const Col c = *col0;
  const int b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0) { *g.left_count = 0u; *g.heavy_count = 0u; }
  if (*c.flag != 0) return;
  const int len = c.count[b], off = c.boff[b];
  if (b == NB - 1 && threadIdx.x == 0) { g.cstart[n] = (int)n; g.cstart[n + 1] = (int)n; }
  if (len == 0) {
    if (threadIdx.x == 0) { g.g1min[b] = 0.0; g.g1sc[b] = 0.0; g.g2min[b] = 0.0; g.g2sc[b] = 0.0; g.c1n[b] = 1; g.c2n[b] = 1; }
    return;
  }
  if (len > kBucketCap) {                       // (colgrid's fine_cells is not run for these columns: the check lives here)
    if (threadIdx.x == 0) atomicOr(c.flag, kFlagOverflow);
    return;
  }
  double mn1 = d_inf(), mx1 = -d_inf(), mn2 = d_inf(), mx2 = -d_inf();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int r = c.srow[off + i];
    const double v1 = g.raw[1][r];
    mn1 = fmin(mn1, v1); mx1 = fmax(mx1, v1);
    if (g.G >= 3) { const double v2 = g.raw[2][r]; mn2 = fmin(mn2, v2); mx2 = fmax(mx2, v2); }
  }
  block_minmax(mn1, mx1, redd);
  if (g.G >= 3) block_minmax(mn2, mx2, redd);
  const int cells = min(len, kMaxCells);
  int C1 = cells, C2 = 1;
  if (g.G >= 3) {
    C1 = max(1, (int)sqrt((double)cells));
    while (C1 * C1 > cells) --C1;
    C2 = C1;
  }
  // The range the cells of a coordinate span: the bucket's [min, max] cut to mean +- 3.5 sd of its rows within 3 sd (three
  // rounds of trimming) - heavy-tailed data would otherwise spend the cells on the range of a few outliers and leave its
  // core in a handful of them.  Values outside fall into the end cells: lin_cell clamps, the map stays monotone, which is
  // all that exactness needs (the row order inside a bucket, hence the last bits of these sums, differs from run to
  // run: the cell geometry may, the results cannot).
  auto robust_range = [&](int coord, double mn, double mx, double& lo, double& hi) {
    lo = mn; hi = mx;
    if (!(mx > mn) || !(mx - mn < d_inf())) return;
    double tl = -d_inf(), th = d_inf(), mean = 0.0, sd = 0.0, cnt = 0.0;
    for (int round = 0; round < 3; ++round) {
      double c0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const double v = g.raw[coord][c.srow[off + i]];
        if (v >= tl && v <= th) { const double u = v - mn; c0 += 1.0; s1 += u; s2 += u * u; }
      }
      block_sum3(c0, s1, s2, redd);
      if (!(c0 >= 2.0)) return;
      cnt = c0;
      const double m1 = s1 / c0;
      mean = mn + m1;
      sd = sqrt(fmax(0.0, s2 / c0 - m1 * m1));
      if (!(sd > 0.0) || !(sd < d_inf())) return;
      tl = mean - 3.0 * sd; th = mean + 3.0 * sd;
    }
    if (cnt < 0.5 * (double)len) return;
    const double l2 = fmax(mn, mean - 3.5 * sd), h2 = fmin(mx, mean + 3.5 * sd);
    if (h2 > l2) { lo = l2; hi = h2; }
  };
  double lo1, hi1, lo2 = 0.0, hi2 = 0.0;
  robust_range(1, mn1, mx1, lo1, hi1);
  if (g.G >= 3) robust_range(2, mn2, mx2, lo2, hi2);
  mn1 = lo1; mx1 = hi1; mn2 = lo2; mx2 = hi2;
  const double sc1 = (mx1 > mn1 && mx1 - mn1 < d_inf()) ? (double)C1 / (mx1 - mn1) : 0.0;
  const double sc2 = (g.G >= 3 && mx2 > mn2 && mx2 - mn2 < d_inf()) ? (double)C2 / (mx2 - mn2) : 0.0;
  if (threadIdx.x == 0) {
    g.g1min[b] = mn1; g.g1sc[b] = sc1; g.g2min[b] = g.G >= 3 ? mn2 : 0.0; g.g2sc[b] = sc2; g.c1n[b] = C1; g.c2n[b] = C2;
  }
  const int ncell = C1 * C2;
  for (int f = threadIdx.x; f < ncell; f += blockDim.x) s_hist[f] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int r = c.srow[off + i];
    int cell = lin_cell(g.raw[1][r], mn1, sc1, C1) * C2;
    if (g.G >= 3) cell += lin_cell(g.raw[2][r], mn2, sc2, C2);
    s_cell[i] = (unsigned short)cell;
    atomicAdd(&s_hist[cell], 1);
  }
  __syncthreads();
  block_excl_scan(s_hist, ncell, red);
  for (int f = threadIdx.x; f < len; f += blockDim.x) g.cstart[off + f] = f < ncell ? off + s_hist[f] : off + len;
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int r = c.srow[off + i];
    const int pos = off + atomicAdd(&s_hist[s_cell[i]], 1);
    g.pc[0][pos] = c.sval[off + i];
    for (int d = 1; d < g.D; ++d) g.pc[d][pos] = g.raw[d][r];
    g.prow[pos] = r;
    g.pbkt[pos] = (unsigned short)b;
  }
}

template <int D>
struct Pt {
  double v[D];
};
template <int D>
__device__ __forceinline__ Pt<D> load_pt(const Grid3& g, int s) {
  Pt<D> p;
#pragma unroll
  for (int d = 0; d < D; ++d) p.v[d] = g.pc[d][s];
  return p;
}
// Chebyshev distance if every coordinate difference is below thr, else +inf (the exact test of the search).  The last
// coordinate goes first and alone: the cells of a window already bound coordinates 0, 1, 2, so in four and more
// dimensions most rows are turned away by a coordinate the grid knows nothing about, with one load instead of D.
template <int D>
__device__ __forceinline__ double cheb_below(const Pt<D>& q, const Grid3& g, int s, double thr) {
  double m = fabs(q.v[D - 1] - g.pc[D - 1][s]);
  if (!(m < thr)) return d_inf();
  bool in = true;
#pragma unroll
  for (int d = 0; d < D - 1; ++d) {
    const double v = fabs(q.v[d] - g.pc[d][s]);
    in = in && v < thr;
    m = v > m ? v : m;
  }
  return in ? m : d_inf();
}

// cell windows of bucket j for a box of half-width t around q (coordinates 1 and 2), widened by 2^-50 relative
struct Win3 {
  int f1lo, f1hi, f2lo, f2hi, C2, off;
};
__device__ __forceinline__ Win3 window3(const Grid3& g, const Col& c, int j, double q1, double q2, double t) {
  Win3 w;
  w.off = c.boff[j];
  const int C1 = g.c1n[j];
  w.C2 = g.c2n[j];
  w.f1lo = 0; w.f1hi = C1 - 1; w.f2lo = 0; w.f2hi = w.C2 - 1;
  if (t < d_inf()) {
    const double m1 = g.g1min[j], s1 = g.g1sc[j];
    w.f1lo = lin_cell((q1 - t) - (fabs(q1) + t) * kSlack, m1, s1, C1);
    w.f1hi = lin_cell((q1 + t) + (fabs(q1) + t) * kSlack, m1, s1, C1);
    if (w.C2 > 1) {
      const double m2 = g.g2min[j], s2 = g.g2sc[j];
      w.f2lo = lin_cell((q2 - t) - (fabs(q2) + t) * kSlack, m2, s2, w.C2);
      w.f2hi = lin_cell((q2 + t) + (fabs(q2) + t) * kSlack, m2, s2, w.C2);
    }
  }
  return w;
}

template <int K1T>
__device__ __forceinline__ double kth_of(const double (&best)[K1T], int k) {
  double r = best[0];
#pragma unroll
  for (int t = 1; t < K1T; ++t) r = (t <= k) ? best[t] : r;
  return r;
}
// the next double above v >= 0 (a gate that lets distances EQUAL to a known bound through the strict test)
__device__ __forceinline__ double next_up(double v) {
  return v < d_inf() ? __longlong_as_double(__double_as_longlong(v) + 1) : v;
}

constexpr int kG3Pre = 2;             // bounding pass: cells either side of the query's own, per grid coordinate

// The cells of coordinates 1, 2 say nothing about the remaining coordinates of a row: a list seeded from the rows next
// to the query in slot order starts with a k-th distance of the order of the whole range of those coordinates, and every
// window cut at it is most of its bucket.  So the search runs in two passes.  Bounding pass: the rows of the
// (2 kG3Pre + 1)^2 cells around the query's own cell in its own bucket (a few dozen rows, all close in coordinates
// 0, 1, 2) are real points of the set, the query among them: their (k + 1)-th smallest distance bounds the true one
// from above.  Exact pass: a fresh list, gated by the next double above that bound, over the windows of every bucket
// the gate reaches, nearest first, the gate tightening to the running k-th distance.  Every row within the true k-th
// distance passes the gate and lies in the windows (monotone cell maps, thresholds widened by 2^-50), so the list ends
// as the exhaustive search's: bit-exact.  Windows of more than `heavy_cells` cells, or more than `near` buckets on a side,
// are left to leftover3_kernel from that bucket on.
template <int D, int K1T>
__global__ void __launch_bounds__(kG3Threads) knn3_kernel(const Col* col0, const Grid3* gp, long long n, int NB, int k, int near,
                                                           int heavy_cells) {
  const Grid3 g = *gp;
  const Col c = *col0;
  if (*c.flag != 0) return;
  const int lane = threadIdx.x & 31;
  const unsigned nblk = gridDim.x;
  const unsigned blk = (blockIdx.x & 1u) ? nblk - 1u - (blockIdx.x >> 1) : (blockIdx.x >> 1);     // both ends of coordinate 0 first
  const long long s64 = (long long)blk * kG3Threads + threadIdx.x;
  const int slot = (int)min(s64, n - 1);
  const bool valid = s64 < n;
  const int b = g.pbkt[slot];
  const int off = c.boff[b];
  const Pt<D> q = load_pt<D>(g, slot);
  const double kInf = d_inf();
  unsigned long long np = 0;
  double best[K1T];
#pragma unroll
  for (int t = 0; t < K1T; ++t) best[t] = kInf;
  double thr = kInf;
  if (valid) {                                        // ---- bounding pass
    const int C1 = g.c1n[b], C2 = g.c2n[b];
    const int c1q = lin_cell(q.v[1], g.g1min[b], g.g1sc[b], C1);
    const int c2q = C2 > 1 ? lin_cell(q.v[2], g.g2min[b], g.g2sc[b], C2) : 0;
    // the neighbourhood grows until it holds a few times k + 1 rows (sparse corners of a bucket's cell grid) or the bucket
    const int need = 4 * (k + 1);
    int r = kG3Pre, f1lo, f1hi, f2lo, f2hi;
    for (;;) {
      const int r1 = C2 > 1 ? r : 2 * r * (r + 1);                          // (one cell coordinate: as many cells in a row)
      f1lo = max(c1q - r1, 0); f1hi = min(c1q + r1, C1 - 1);
      f2lo = max(c2q - r, 0); f2hi = min(c2q + r, C2 - 1);
      if (f1lo == 0 && f1hi == C1 - 1 && f2lo == 0 && f2hi == C2 - 1) break;
      int cnt = 0;
      for (int c1 = f1lo; c1 <= f1hi; ++c1) cnt += g.cstart[off + c1 * C2 + f2hi + 1] - g.cstart[off + c1 * C2 + f2lo];
      if (cnt >= need) break;
      r *= 2;
    }
    for (int c1 = f1lo; c1 <= f1hi; ++c1) {
      const int a = g.cstart[off + c1 * C2 + f2lo], e = g.cstart[off + c1 * C2 + f2hi + 1];
      for (int s = a; s < e; ++s) {
        const double m = cheb_below<D>(q, g, s, thr);
        if (m < thr) {
          topk_insert<K1T>(best, m);
          thr = kth_of<K1T>(best, k);
        }
      }
      if (e > a) np += (unsigned long long)(e - a);
    }
  }
  const double gate = next_up(thr);
#pragma unroll
  for (int t = 0; t < K1T; ++t) best[t] = kInf;
  thr = gate;
  // ---- exact pass.  Every run of bucket j's window; false: the window is too large for one thread (nothing was examined)
  auto visit = [&](int j) -> bool {
    const Win3 w = window3(g, c, j, q.v[1], q.v[2], thr);
    if ((w.f1hi - w.f1lo + 1) * (w.f2hi - w.f2lo + 1) > heavy_cells) return false;
    for (int c1 = w.f1lo; c1 <= w.f1hi; ++c1) {
      const int a = g.cstart[w.off + c1 * w.C2 + w.f2lo], e = g.cstart[w.off + c1 * w.C2 + w.f2hi + 1];
      for (int s = a; s < e; ++s) {
        const double m = cheb_below<D>(q, g, s, thr);
        if (m < thr) {
          topk_insert<K1T>(best, m);
          thr = fmin(gate, kth_of<K1T>(best, k));
        }
      }
      if (e > a) np += (unsigned long long)(e - a);
    }
    return true;
  };
  int rstart = NB, lend = 0;
  bool heavy = false;
  if (valid && !visit(b)) { heavy = true; rstart = b; lend = b; }
  {
    bool more = valid && !heavy;
    int j = b + 1;
    for (int t = 0; t < near && __any_sync(kFull, more); ++t) {
      if (more) {
        while (j < NB && c.count[j] == 0) ++j;
        if (j >= NB || (c.vlo[j] - q.v[0]) >= thr) more = false;
        else if (!visit(j)) { rstart = j; more = false; }
        else ++j;
      }
    }
    if (more) {
      while (j < NB && c.count[j] == 0) ++j;
      if (j < NB && !((c.vlo[j] - q.v[0]) >= thr)) rstart = j;
    }
  }
  {
    bool more = valid && !heavy;
    int j = b - 1;
    for (int t = 0; t < near && __any_sync(kFull, more); ++t) {
      if (more) {
        while (j >= 0 && c.count[j] == 0) --j;
        if (j < 0 || (q.v[0] - c.vhi[j]) >= thr) more = false;
        else if (!visit(j)) { lend = j + 1; more = false; }
        else --j;
      }
    }
    if (more) {
      while (j >= 0 && c.count[j] == 0) --j;
      if (j >= 0 && !((q.v[0] - c.vhi[j]) >= thr)) lend = j + 1;
    }
  }
  const bool want = valid && (rstart < NB || lend > 0);
  const unsigned m = __ballot_sync(kFull, want);
  int my_e = -1;
  if (m != 0) {
    const int leader = __ffs(m) - 1;
    unsigned base_e = 0;
    if (lane == leader) base_e = atomicAdd(g.left_count, (unsigned)__popc(m));
    base_e = __shfl_sync(kFull, base_e, leader);
    if (want) my_e = (int)(base_e + __popc(m & ((1u << lane) - 1u)));       // (the list holds one entry per row: never full)
  }
  if (valid) {
    if (my_e >= 0) {
      LeftEnt le;
      le.slot = slot; le.rstart = rstart; le.lend = lend; le.skip_a = 0;       // (no seeds to skip on this path)
      g.left[my_e] = le;
      g.eps[slot] = thr;                                 // the gate the leftover kernel starts from
#pragma unroll
      for (int t = 0; t < K1T; ++t) g.left_best[(long long)my_e * K1T + t] = best[t];
    } else {
      const double r = kth_of<K1T>(best, k);
      g.eps[slot] = r;
      g.eps_row[g.prow[slot]] = r;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np && g.pairs) atomicAdd(g.pairs, np);
}

// One warp per deferred query: the lanes take the remaining buckets round-robin, each walking the runs of its bucket's
// window on its own; private lists per lane, gated by min(gate, the lane's own k-th distance), merged at the end with the
// list the query's thread left.  A query without a bound (fewer than k + 1 rows in its own bucket) first gets one from
// whole buckets, its own outwards.
template <int D, int K1T>
__global__ void __launch_bounds__(kG3Threads) leftover3_kernel(const Col* col0, const Grid3* gp, int NB, int k) {
  const Grid3 g = *gp;
  const Col c = *col0;
  if (*c.flag != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nent = *g.left_count;
  const double kInf = d_inf();
  const int k1 = k + 1;
  unsigned long long np = 0;
  for (unsigned e = blockIdx.x * (kG3Threads / 32) + warp; e < nent; e += gridDim.x * (kG3Threads / 32)) {
    const LeftEnt le = g.left[e];
    const Pt<D> q = load_pt<D>(g, le.slot);
    double gate = g.eps[le.slot];
    double best[K1T];
    if (!(gate < kInf)) {
      // no bound yet (the query's own bucket holds fewer than k + 1 rows): whole buckets, its own first, then outwards,
      // until k + 1 rows have been seen; their (k + 1)-th smallest distance is the bound
      const int b = g.pbkt[le.slot];
#pragma unroll
      for (int u = 0; u < K1T; ++u) best[u] = kInf;
      double bound = kInf;
      for (int t = 0; !(bound < kInf) && t < NB; ++t) {
        for (int side = 0; side < (t == 0 ? 1 : 2); ++side) {
          const int j = side == 0 ? b + t : b - t;
          if (j < 0 || j >= NB) continue;
          const int jo = c.boff[j], jl = c.count[j];
          for (int s = jo + lane; s < jo + jl; s += 32) {
            const double m = cheb_below<D>(q, g, s, kInf);
            if (m < best[K1T - 1]) topk_insert<K1T>(best, m);
          }
        }
        bound = __shfl_sync(kFull, warp_merge_lists<K1T>(best, k1), k);
      }
      gate = next_up(bound);
    }
#pragma unroll
    for (int t = 0; t < K1T; ++t) best[t] = kInf;
    double thr = gate;
    // the lanes take the remaining buckets round-robin, each walking the runs of its bucket's window on its own
    auto bucket = [&](int j) {
      const Win3 w = window3(g, c, j, q.v[1], q.v[2], thr);
      for (int c1 = w.f1lo; c1 <= w.f1hi; ++c1) {
        const int a = g.cstart[w.off + c1 * w.C2 + w.f2lo], en = g.cstart[w.off + c1 * w.C2 + w.f2hi + 1];
        for (int s = a; s < en; ++s) {
          const double m = cheb_below<D>(q, g, s, thr);
          if (m < thr) {
            topk_insert<K1T>(best, m);
            thr = fmin(best[K1T - 1], gate);
          }
        }
        if (en > a) np += (unsigned long long)(en - a);
      }
    };
    for (int j = le.rstart + lane; j < NB; j += 32) {
      if (c.count[j] == 0) continue;
      if ((c.vlo[j] - q.v[0]) >= thr) break;                 // (vlo ascends with j: every later bucket of this lane fails too)
      bucket(j);
    }
    for (int j = le.lend - 1 - lane; j >= 0; j -= 32) {
      if (c.count[j] == 0) continue;
      if ((q.v[0] - c.vhi[j]) >= thr) break;
      bucket(j);
    }
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < K1T; ++t) {
        const double v = g.left_best[(long long)e * K1T + t];
        if (v < best[K1T - 1]) topk_insert<K1T>(best, v);
      }
    }
    const double fin = warp_merge_lists<K1T>(best, k1);
    if (lane == k) { g.eps[le.slot] = fin; g.eps_row[g.prow[le.slot]] = fin; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np && g.pairs) atomicAdd(g.pairs, np);
}


// Frenzel-Pompe counts (:152-154): coordinates [0, C) are the condition z, C is x, C + 1 is y; radius r = eps - 1e-12,
// inclusive.  One thread per query walks the runs of the windows of the buckets its radius reaches; queries with a wide
// radius (sparse regions: windows of many cells, or many buckets) are left to count3_heavy_kernel, a warp each.
constexpr int kG3CountBuckets = 64;
template <int C>
__global__ void __launch_bounds__(kG3Threads) count3_kernel(const Col* col0, const Grid3* gp, long long n, int NB) {
  constexpr int D = C + 2;
  const Grid3 g = *gp;
  const Col c = *col0;
  if (*c.flag != 0) return;
  const int lane = threadIdx.x & 31;
  const unsigned nblk = gridDim.x;
  const unsigned blk = (blockIdx.x & 1u) ? nblk - 1u - (blockIdx.x >> 1) : (blockIdx.x >> 1);
  const long long s64 = (long long)blk * kG3Threads + threadIdx.x;
  const bool valid = s64 < n;
  const int slot = (int)min(s64, n - 1);
  const Pt<D> q = load_pt<D>(g, slot);
  const double r = g.eps[slot] - 1e-12;                      // _entropy_estimators.py:109
  int nz = 0, nxz = 0, nyz = 0;
  unsigned long long np = 0;
  bool defer = false;
  if (valid && r >= 0.0) {
    const int b = g.pbkt[slot];
    auto bucket = [&](int j) -> bool {
      const Win3 w = window3(g, c, j, q.v[1], q.v[2], r);      // (the window of a box of half-width r, widened: a superset)
      if ((w.f1hi - w.f1lo + 1) * (w.f2hi - w.f2lo + 1) > kG3Heavy) return false;
      for (int c1 = w.f1lo; c1 <= w.f1hi; ++c1) {
        const int a = g.cstart[w.off + c1 * w.C2 + w.f2lo], e = g.cstart[w.off + c1 * w.C2 + w.f2hi + 1];
        for (int s = a; s < e; ++s) {
          bool in = true;
#pragma unroll
          for (int d = 0; d < C; ++d) in = in && fabs(q.v[d] - g.pc[d][s]) <= r;
          if (in) {
            ++nz;
            nxz += (int)(fabs(q.v[C] - g.pc[C][s]) <= r);
            nyz += (int)(fabs(q.v[C + 1] - g.pc[C + 1][s]) <= r);
          }
        }
        if (e > a) np += (unsigned long long)(e - a);
      }
      return true;
    };
    defer = !bucket(b);
    // a bucket can hold a row within r only while the gap in coordinate 0 does not exceed r (rounded subtraction is
    // monotone; the values of bucket j lie in (vlo, vhi])
    int visited = 0;
    for (int j = b + 1; j < NB && !defer; ++j) {
      if (c.count[j] == 0) continue;
      if ((c.vlo[j] - q.v[0]) > r) break;
      if (++visited > kG3CountBuckets || !bucket(j)) defer = true;
    }
    for (int j = b - 1; j >= 0 && !defer; --j) {
      if (c.count[j] == 0) continue;
      if ((q.v[0] - c.vhi[j]) > r) break;
      if (++visited > kG3CountBuckets || !bucket(j)) defer = true;
    }
  }
  const unsigned m = __ballot_sync(kFull, defer);
  if (m != 0) {
    const int leader = __ffs(m) - 1;
    unsigned base_e = 0;
    if (lane == leader) base_e = atomicAdd(g.heavy_count, (unsigned)__popc(m));
    base_e = __shfl_sync(kFull, base_e, leader);
    if (defer) g.heavy[base_e + __popc(m & ((1u << lane) - 1u))] = slot;
  }
  if (valid && !defer) {
    const int row = g.prow[slot];
    g.cnt_row[0][row] = nz;
    g.cnt_row[1][row] = nxz;
    g.cnt_row[2][row] = nyz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np && g.pairs) atomicAdd(g.pairs, np);
}

template <int C>
__global__ void __launch_bounds__(kG3Threads) count3_heavy_kernel(const Col* col0, const Grid3* gp, int NB) {
  constexpr int D = C + 2;
  const Grid3 g = *gp;
  const Col c = *col0;
  if (*c.flag != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nent = *g.heavy_count;
  unsigned long long np = 0;
  for (unsigned e = blockIdx.x * (kG3Threads / 32) + warp; e < nent; e += gridDim.x * (kG3Threads / 32)) {
    const int slot = g.heavy[e];
    const Pt<D> q = load_pt<D>(g, slot);
    const double r = g.eps[slot] - 1e-12;
    int nz = 0, nxz = 0, nyz = 0;
    auto bucket = [&](int j) {
      const Win3 w = window3(g, c, j, q.v[1], q.v[2], r);
      for (int c1 = w.f1lo; c1 <= w.f1hi; ++c1) {
        const int a = g.cstart[w.off + c1 * w.C2 + w.f2lo], en = g.cstart[w.off + c1 * w.C2 + w.f2hi + 1];
        for (int s = a + lane; s < en; s += 32) {
          bool in = true;
#pragma unroll
          for (int d = 0; d < C; ++d) in = in && fabs(q.v[d] - g.pc[d][s]) <= r;
          if (in) {
            ++nz;
            nxz += (int)(fabs(q.v[C] - g.pc[C][s]) <= r);
            nyz += (int)(fabs(q.v[C + 1] - g.pc[C + 1][s]) <= r);
          }
        }
        if (lane == 0 && en > a) np += (unsigned long long)(en - a);
      }
    };
    const int b = g.pbkt[slot];
    bucket(b);
    for (int j = b + 1; j < NB; ++j) {
      if (c.count[j] == 0) continue;
      if ((c.vlo[j] - q.v[0]) > r) break;
      bucket(j);
    }
    for (int j = b - 1; j >= 0; --j) {
      if (c.count[j] == 0) continue;
      if ((q.v[0] - c.vhi[j]) > r) break;
      bucket(j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nz += __shfl_xor_sync(kFull, nz, o);
      nxz += __shfl_xor_sync(kFull, nxz, o);
      nyz += __shfl_xor_sync(kFull, nyz, o);
    }
    if (lane == 0) {
      const int row = g.prow[slot];
      g.cnt_row[0][row] = nz;
      g.cnt_row[1][row] = nxz;
      g.cnt_row[2][row] = nyz;
    }
  }
  if (lane == 0 && np && g.pairs) atomicAdd(g.pairs, np);
}

}  // namespace

Plan make_plan(int64_t n, bool* ok) {
  Plan p;
  p.n = n;
  // rows per bucket ~ 1.4 sqrt(n) (measured optimum): a bucket should be about as wide in x as a typical k-th neighbour distance, which
  // shrinks like 1 / sqrt(n) in two dimensions
  int64_t rows = 256, factor4 = 8;                            // factor4 / 4 = (rows / sqrt(n))^2
  if (const char* e = getenv("EB2_K2_ROWS4")) factor4 = atoll(e);     // tuning knob
  while (rows < kCoarseRows && rows * rows * 4 < factor4 * n) rows += 64;
  int64_t Bc = (n + rows - 1) / rows;
  if (Bc < 1) Bc = 1;
  bool fits = n >= 2 && n < (int64_t(1) << 30);
  if (Bc > kMaxCoarse) {
    Bc = kMaxCoarse;
    if (n > int64_t(kMaxCoarse) * (kCoarseRows + kCoarseRows / 2)) fits = false;     // buckets would run too full
  }
  p.Bc = static_cast<int>(Bc);
  p.NB = p.Bc * kSub;
  int64_t over = n / Bc;
  if (over > kOversample) over = kOversample;
  while (over > 1 && over * Bc > 4096) --over;                 // the sample (over * Bc values) is ranked in shared memory
  if (over < 1) over = 1;
  p.over = static_cast<int>(over);
  if (ok) *ok = fits;
  return p;
}

size_t col_bytes(int64_t n) {
  size_t b = 0;
  b += align256(sizeof(double) * kMaxCoarse) * 3;            // split, clo, csc
  b += align256(sizeof(double) * kMaxCoarse * kOversample) * 2;  // ssort, sraw
  b += align256(sizeof(int) * (kMaxBuckets + 1)) * 4;        // count, fill, boff, ncell
  b += align256(sizeof(double) * kMaxBuckets) * 3;           // vlo, vhi, fsc
  b += align256(sizeof(FineGrid) * kMaxBuckets);             // fg
  b += align256(sizeof(unsigned short) * n);                 // bkt
  b += align256(sizeof(int) * n);                            // srow
  b += align256(sizeof(double) * n) * 2;                     // sval, fval
  b += align256(sizeof(int) * (n + 2));                      // fstart
  b += 256;                                                  // flag
  return b;
}

Col carve_col(char* base, int64_t n, const double* vals) {
  Col c;
  size_t o = 0;
  auto take = [&](size_t bytes) { char* p = base + o; o += align256(bytes); return p; };
  c.vals = vals;
  c.split = reinterpret_cast<double*>(take(sizeof(double) * kMaxCoarse));
  c.clo = reinterpret_cast<double*>(take(sizeof(double) * kMaxCoarse));
  c.csc = reinterpret_cast<double*>(take(sizeof(double) * kMaxCoarse));
  c.ssort = reinterpret_cast<double*>(take(sizeof(double) * kMaxCoarse * kOversample));
  c.sraw = reinterpret_cast<double*>(take(sizeof(double) * kMaxCoarse * kOversample));
  c.count = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.fill = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.boff = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.ncell = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.vlo = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.vhi = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.fsc = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.fg = reinterpret_cast<FineGrid*>(take(sizeof(FineGrid) * kMaxBuckets));
  c.bkt = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * n));
  c.srow = reinterpret_cast<int*>(take(sizeof(int) * n));
  c.sval = reinterpret_cast<double*>(take(sizeof(double) * n));
  c.fval = reinterpret_cast<double*>(take(sizeof(double) * n));
  c.fstart = reinterpret_cast<int*>(take(sizeof(int) * (n + 2)));
  c.flag = reinterpret_cast<int*>(take(sizeof(int)));
  return c;
}

// entries the deferral list of a problem can hold (a full list only means that the search kernel keeps going itself)
static size_t left_cap(const Plan& p) {
  size_t cap = static_cast<size_t>(p.n / 4 + 1024);
  if (const char* e = getenv("EB2_K2_LEFTCAP")) cap = std::min(cap, static_cast<size_t>(std::max(1LL, atoll(e))));   // test knob: a full list
  return cap;
}

size_t prob_bytes(const Plan& p, int k1t) {
  size_t b = 0;
  b += align256(sizeof(double) * p.n) * 3;                    // px, py, eps
  b += align256(sizeof(int) * p.n);                           // prow
  b += align256(sizeof(unsigned short) * p.n);                // pbkt
  b += align256(sizeof(int) * (p.n + 2));                     // cstart
  b += align256(sizeof(double) * kMaxBuckets) * 2;            // bymin, bysc
  b += align256(sizeof(LeftEnt) * left_cap(p));               // left
  b += align256(sizeof(double) * left_cap(p) * k1t);          // left_best
  b += 256 * 3;                                               // left_count, acc, out
  return b;
}

Prob carve_prob(char* base, const Plan& p, int k1t, int cx, int cy) {
  Prob q;
  size_t o = 0;
  auto take = [&](size_t bytes) { char* ptr = base + o; o += align256(bytes); return ptr; };
  q.cx = cx; q.cy = cy;
  q.px = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  q.py = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  q.eps = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  q.prow = reinterpret_cast<int*>(take(sizeof(int) * p.n));
  q.pbkt = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * p.n));
  q.cstart = reinterpret_cast<int*>(take(sizeof(int) * (p.n + 2)));
  q.bymin = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  q.bysc = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  q.left = reinterpret_cast<LeftEnt*>(take(sizeof(LeftEnt) * left_cap(p)));
  q.left_best = reinterpret_cast<double*>(take(sizeof(double) * left_cap(p) * k1t));
  q.left_count = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int)));
  q.acc = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * 4));
  q.out = reinterpret_cast<double*>(take(sizeof(double) * 16));
  q.eps_row = nullptr; q.nx_row = nullptr; q.ny_row = nullptr;
  return q;
}

cudaError_t init() { return cudaSuccess; }      // (no kernel needs an opt-in shared-memory size at present)

cudaError_t colgrid_buckets(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches) {
  const int nb = static_cast<int>((p.n + kHistRows - 1) / kHistRows);
  const int S = p.Bc * p.over;
  sample_gather_kernel<<<dim3((S + kRankThreads - 1) / kRankThreads, ncol), kRankThreads, 0, s>>>(cols, p.n, S);
  sample_rank_kernel<<<dim3((S + 31) / 32, ncol), kRankThreads, 0, s>>>(cols, S);
  split_tables_kernel<<<ncol, kSplitThreads, 0, s>>>(cols, p.Bc, p.over, p.NB);
  bucket_hist_kernel<<<dim3(nb, ncol), kThreadsB, 0, s>>>(cols, p.n, p.Bc, p.NB);
  bucket_scatter_kernel<<<dim3(nb, ncol), kThreadsB, 0, s>>>(cols, p.n, p.NB);
  if (launches) *launches += 5;
  return cudaGetLastError();
}

cudaError_t colgrid_cells(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches) {
  fine_cells_kernel<<<dim3(p.NB, ncol), kThreadsB, 0, s>>>(cols, p.n, p.NB);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t colgrid(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches) {
  const cudaError_t e = colgrid_buckets(cols, ncol, p, s, launches);
  return e != cudaSuccess ? e : colgrid_cells(cols, ncol, p, s, launches);
}

cudaError_t layout(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches) {
  layout_kernel<<<dim3(p.NB, nprob), kLayoutThreads, 0, s>>>(cols, probs, p.n, p.NB);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t knn(const Col* cols, const Prob* probs, int nprob, const Plan& p, int k, const Shard& sh, int sm_count,
                cudaStream_t s, int* launches) {
  int near = 3;
  if (const char* e = getenv("EB2_K2_NEAR")) near = atoi(e);       // tuning knob: buckets per side before a query is deferred
  const int grid = static_cast<int>((p.n + kThreadsB - 1) / kThreadsB);
  const int lgrid = nprob > 1 ? std::max(1, sm_count * 8 / nprob) : sm_count * 8;
  const unsigned cap = static_cast<unsigned>(left_cap(p));
  if (k + 1 <= 4) {
    knn_kernel2<4><<<dim3(grid, nprob), kThreadsB, 0, s>>>(cols, probs, p.n, p.NB, k, near, cap, sh);
    leftover_kernel2<4><<<dim3(lgrid, nprob), kThreadsB, 0, s>>>(cols, probs, p.NB, k, cap);
  } else if (k + 1 <= 8) {
    knn_kernel2<8><<<dim3(grid, nprob), kThreadsB, 0, s>>>(cols, probs, p.n, p.NB, k, near, cap, sh);
    leftover_kernel2<8><<<dim3(lgrid, nprob), kThreadsB, 0, s>>>(cols, probs, p.NB, k, cap);
  } else {
    return cudaErrorInvalidValue;
  }
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t count_psi(const Col* cols, const Prob* probs, int nprob, const Plan& p, const Shard& sh, const double* psi_tab,
                      int tab_n, cudaStream_t s, int* launches) {
  const int grid = static_cast<int>((p.n + kThreadsB - 1) / kThreadsB);
  count_psi_kernel<<<dim3(grid, nprob), kThreadsB, 0, s>>>(cols, probs, p.n, p.Bc, sh, psi_tab, tab_n);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t finalize(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches) {
  (void)p;
  final_kernel<<<nprob, 1, 0, s>>>(cols, probs);
  if (launches) *launches += 1;
  return cudaGetLastError();
}


// ---- three-level grid: host side ---------------------------------------------------------------------------------------
Plan make_plan3(int64_t n, int D, bool* ok) {
  (void)D;
  Plan p;
  p.n = n;
  // buckets of ~3,072 rows (C1 x C2 = 55 x 55 cells; three eighths of what a bucket may hold: with up to 32 samples per
  // splitter the bucket sizes scatter by ~20 %): in three and more dimensions a k-th neighbour distance spans several
  // buckets anyway.  Measured with the window limit of knn3 (4-D, N = 5*10^5): 2,048 rows / 256 cells 8.45 ms, 3,072 / 512 7.84 ms
  int64_t rows = 3072;
  if (const char* e = getenv("EB2_G3_ROWS")) rows = atoll(e);           // tuning knob
  if (rows < 256) rows = 256;
  if (rows > kBucketCap / 2) rows = kBucketCap / 2;
  int64_t Bc = (n + rows - 1) / rows;
  if (Bc < 1) Bc = 1;
  bool fits = n >= 2 && n < (int64_t(1) << 30);
  if (Bc > kMaxCoarse) { Bc = kMaxCoarse; if (n > int64_t(kMaxCoarse) * (kBucketCap / 2)) fits = false; }
  p.Bc = static_cast<int>(Bc);
  p.NB = p.Bc * kSub;
  int64_t over = n / Bc;                                          // (as many samples per bucket as the ranking holds: the fuller
  if (over > 32) over = 32;                                       //  buckets of this grid need tighter quantiles)
  if (over < 1) over = 1;
  while (over > 1 && over * Bc > 4096) --over;
  p.over = static_cast<int>(over);
  if (ok) *ok = fits;
  return p;
}

size_t grid3_bytes(const Plan& p, int D, int k1t) {
  size_t b = 0;
  b += align256(sizeof(double) * p.n) * (D + 2);              // pc, eps, eps_row
  b += align256(sizeof(int) * p.n) * 5;                       // prow, cnt_row x3, heavy
  b += align256(sizeof(unsigned short) * p.n);                // pbkt
  b += align256(sizeof(int) * (p.n + 2));                     // cstart
  b += align256(sizeof(double) * kMaxBuckets) * 4;            // cell maps
  b += align256(sizeof(int) * kMaxBuckets) * 2;               // c1n, c2n
  b += align256(sizeof(LeftEnt) * p.n);                       // left
  b += align256(sizeof(double) * p.n * k1t);                  // left_best
  b += 256 * 3;                                               // left_count, pairs (caller's), spare
  return b;
}

Grid3 carve_grid3(char* base, const Plan& p, int D, int G, int k1t) {
  Grid3 g;
  size_t o = 0;
  auto take = [&](size_t bytes) { char* ptr = base + o; o += align256(bytes); return ptr; };
  g.D = D; g.G = G;
  for (int d = 0; d < kG3MaxD; ++d) { g.raw[d] = nullptr; g.pc[d] = nullptr; }
  for (int d = 0; d < D; ++d) g.pc[d] = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  g.eps = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  g.eps_row = reinterpret_cast<double*>(take(sizeof(double) * p.n));
  g.prow = reinterpret_cast<int*>(take(sizeof(int) * p.n));
  for (int i = 0; i < 3; ++i) g.cnt_row[i] = reinterpret_cast<int*>(take(sizeof(int) * p.n));
  g.pbkt = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * p.n));
  g.cstart = reinterpret_cast<int*>(take(sizeof(int) * (p.n + 2)));
  g.g1min = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  g.g1sc = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  g.g2min = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  g.g2sc = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  g.c1n = reinterpret_cast<int*>(take(sizeof(int) * kMaxBuckets));
  g.c2n = reinterpret_cast<int*>(take(sizeof(int) * kMaxBuckets));
  g.left = reinterpret_cast<LeftEnt*>(take(sizeof(LeftEnt) * p.n));
  g.left_best = reinterpret_cast<double*>(take(sizeof(double) * p.n * k1t));
  g.left_count = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int)));
  g.heavy_count = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int)));
  g.heavy = reinterpret_cast<int*>(take(sizeof(int) * p.n));
  g.pairs = nullptr;
  g.flag = nullptr;
  return g;
}

cudaError_t layout3(const Col* col0, const Grid3* g, const Plan& p, cudaStream_t s, int* launches) {
  layout3_kernel<<<p.NB, kG3Threads, 0, s>>>(col0, g, p.n, p.NB);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

template <int D>
static cudaError_t knn3_d(const Col* col0, const Grid3* g, const Plan& p, int k, int near, int sm_count, cudaStream_t s) {
  const int grid = static_cast<int>((p.n + kG3Threads - 1) / kG3Threads);
  int heavy = 512;       // cells of one bucket's window a single thread still walks (the lanes of a warp are neighbours:
                         // their windows are alike)
  if (const char* e = getenv("EB2_G3_HEAVY")) heavy = atoi(e);        // tuning knob
  if (k + 1 <= 4) {
    knn3_kernel<D, 4><<<grid, kG3Threads, 0, s>>>(col0, g, p.n, p.NB, k, near, heavy);
    leftover3_kernel<D, 4><<<sm_count * 8, kG3Threads, 0, s>>>(col0, g, p.NB, k);
  } else {
    knn3_kernel<D, 8><<<grid, kG3Threads, 0, s>>>(col0, g, p.n, p.NB, k, near, heavy);
    leftover3_kernel<D, 8><<<sm_count * 8, kG3Threads, 0, s>>>(col0, g, p.NB, k);
  }
  return cudaGetLastError();
}

cudaError_t knn3(const Col* col0, const Grid3* g, const Grid3& h, const Plan& p, int k, int sm_count, cudaStream_t s, int* launches) {
  if (k + 1 > 8) return cudaErrorInvalidValue;
  int near = 48;         // (buckets are narrower than a k-th distance in 3+ dimensions: a query visits a few dozen)
  if (const char* e = getenv("EB2_G3_NEAR")) near = atoi(e);          // tuning knob
  if (launches) *launches += 2;
  switch (h.D) {
    case 3: return knn3_d<3>(col0, g, p, k, near, sm_count, s);
    case 4: return knn3_d<4>(col0, g, p, k, near, sm_count, s);
    case 5: return knn3_d<5>(col0, g, p, k, near, sm_count, s);
    case 6: return knn3_d<6>(col0, g, p, k, near, sm_count, s);
    case 7: return knn3_d<7>(col0, g, p, k, near, sm_count, s);
    case 8: return knn3_d<8>(col0, g, p, k, near, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t count3(const Col* col0, const Grid3* g, const Grid3& h, const Plan& p, int C, cudaStream_t s, int* launches) {
  (void)h;
  const int grid = static_cast<int>((p.n + kG3Threads - 1) / kG3Threads);
  const int hgrid = 148 * 8;
  switch (C) {
#define EB2_G3_COUNT(CC)                                                                  \
    case CC:                                                                              \
      count3_kernel<CC><<<grid, kG3Threads, 0, s>>>(col0, g, p.n, p.NB);                  \
      count3_heavy_kernel<CC><<<hgrid, kG3Threads, 0, s>>>(col0, g, p.NB);                \
      break;
    EB2_G3_COUNT(2) EB2_G3_COUNT(3) EB2_G3_COUNT(4) EB2_G3_COUNT(5) EB2_G3_COUNT(6)
#undef EB2_G3_COUNT
    default: return cudaErrorInvalidValue;
  }
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace k2
}  // namespace eb2
