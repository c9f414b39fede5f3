// ennemi_b200 — kernels of the bivariate KSG pipeline (sm_100a).  See eb2_ksg2.h for the data layout.
//
// What each stage replaces in the reference (ennemi/_entropy_estimators.py):
//   colsort + layout   the three cKDTree builds (:100-102)
//   knn                grid.query(xy, k=[k+1], p=inf)                         (:108)
//   count_psi          x_grid / y_grid.query_ball_point(.., eps - 1e-12, p=inf, return_length=True) and the
//                      digamma terms of the mean                              (:109-110, :113, :327-350)
// Bit-exactness rules are those of eb2_kernels.cuh: one rounded fp64 subtraction per coordinate, exact
// comparisons, conservative (widened) brackets decided by the exact per-candidate test.
#include <cub/block/block_merge_sort.cuh>

#include <algorithm>
#include <cstdlib>

#include "eb2_kernels.cuh"
#include "eb2_ksg2.h"

namespace eb2 {
namespace k2 {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSortThreads = 512;
constexpr int kSplitThreads = 1024;
constexpr int kSplitItems = 8;                 // kSplitThreads * kSplitItems = 8,192 samples at most
constexpr int kCountRows = 2048;               // rows per CTA of the bucket count / scatter kernels
constexpr int kPiece = 128;                    // slots of a chunk window staged per warp at a time
constexpr int kSeed = 4;                       // slots on either side of a query's own slot that seed its list
constexpr int kWarps = 8;                      // warps per CTA of the search kernels
constexpr double kSlack = 8.881784197001252e-16;   // 2^-50

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// (value, row): a total order, so that every sort below has ONE possible result (deterministic layouts)
struct KV {
  double v;
  int r;
  int pad;
};
struct KVLess {
  __device__ __forceinline__ bool operator()(const KV& a, const KV& b) const { return a.v < b.v || (a.v == b.v && a.r < b.r); }
};
struct DLess {
  __device__ __forceinline__ bool operator()(double a, double b) const { return a < b; }
};

// bucket of v: number of splitters strictly below it
__device__ __forceinline__ int bucket_of(const double* split, int nsplit, double v) {
  int lo = 0, hi = nsplit;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (split[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// sum of count[0 .. b) and of the same counts rounded up to 32, by the whole CTA (b <= kMaxBuckets)
__device__ __forceinline__ void bucket_offsets(const int* count, int b, int* red /* 2 * 32 ints */, int& rank_off, int& slot_off) {
  int a = 0, s = 0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) {
    const int c = count[i];
    a += c;
    s += (c + 31) & ~31;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    s += __shfl_xor_sync(kFull, s, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[w] = a; red[32 + w] = s; }
  __syncthreads();
  a = 0; s = 0;
  for (int i = 0; i < nw; ++i) { a += red[i]; s += red[32 + i]; }
  rank_off = a;
  slot_off = s;
}

// ---- colsort 1: splitters from a sorted regular sample ---------------------------------------------------------
__global__ void __launch_bounds__(kSplitThreads) split_kernel(const Col* cols, long long n, int B, int over) {
  using Sort = cub::BlockMergeSort<double, kSplitThreads, kSplitItems>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename Sort::TempStorage& tmp = *reinterpret_cast<typename Sort::TempStorage*>(smem_raw);
  const Col c = cols[blockIdx.x];
  const int S = B * over;
  double key[kSplitItems];
#pragma unroll
  for (int i = 0; i < kSplitItems; ++i) {
    const int g = threadIdx.x * kSplitItems + i;
    // sample g sits in the middle of the g-th of S equal stretches of the column
    key[i] = g < S ? c.vals[(long long)(((2 * (long long)g + 1) * n) / (2 * (long long)S))] : d_inf();
    if (key[i] != key[i]) key[i] = d_inf();      // NaN input: reported by the count kernel, must not upset the sort
  }
  if (B > 1) {
    Sort(tmp).Sort(key, DLess(), S, d_inf());
#pragma unroll
    for (int i = 0; i < kSplitItems; ++i) {
      const int g = threadIdx.x * kSplitItems + i;
      if (g < S && (g + 1) % over == 0) {
        const int j = (g + 1) / over - 1;
        if (j < B - 1) c.split[j] = key[i];
      }
    }
  }
  for (int b = threadIdx.x; b < kMaxBuckets; b += blockDim.x) { c.count[b] = 0; c.fill[b] = 0; }
}

// ---- colsort 2: bucket of every row, rows per bucket -------------------------------------------------------------
__global__ void __launch_bounds__(256) bucket_count_kernel(const Col* cols, long long n, int B) {
  __shared__ double s_split[kMaxBuckets];
  __shared__ int s_hist[kMaxBuckets];
  const Col c = cols[blockIdx.y];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    s_split[b] = b < B - 1 ? c.split[b] : d_inf();
    s_hist[b] = 0;
  }
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * kCountRows;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kCountRows / 256; ++i) {
    const long long r = r0 + i * 256 + threadIdx.x;
    if (r < n) {
      const double v = c.vals[r];
      if (!(fabs(v) < d_inf())) bad = true;
      const int b = bucket_of(s_split, B - 1, v);
      c.bid[r] = (unsigned short)b;
      atomicAdd(&s_hist[b], 1);
    }
  }
  if (bad) atomicOr(c.flag, kFlagNonFinite);
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    if (s_hist[b]) atomicAdd(&c.count[b], s_hist[b]);
}

// ---- colsort 3: scatter (value, row) into the buckets --------------------------------------------------------------
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const Col* cols, long long n, int B) {
  __shared__ int s_off[kMaxBuckets];
  __shared__ int s_warp[8];
  const Col c = cols[blockIdx.y];
  // exclusive scan of the bucket counts: 4 consecutive buckets per thread
  {
    int v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = threadIdx.x * 4 + i;
      v[i] = b < B ? c.count[b] : 0;
      sum += v[i];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) base += s_warp[w];
    int run = base + inc - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = threadIdx.x * 4 + i;
      if (b < kMaxBuckets) s_off[b] = run;
      run += v[i];
    }
    __syncthreads();
  }
  const long long r0 = (long long)blockIdx.x * kCountRows;
#pragma unroll
  for (int i = 0; i < kCountRows / 256; ++i) {
    const long long r = r0 + i * 256 + threadIdx.x;
    if (r < n) {
      const int b = c.bid[r];
      const int pos = s_off[b] + atomicAdd(&c.fill[b], 1);
      c.st_val[pos] = c.vals[r];
      c.st_row[pos] = (int)r;
    }
  }
}

// one bucket of (value, row) pairs, sorted by the CTA in shared memory
template <int IPT, typename ValueT>
struct BucketSort {
  using Sort = cub::BlockMergeSort<KV, kSortThreads, IPT, ValueT>;
};

// ---- colsort 4: sort every bucket; per-bucket tables ---------------------------------------------------------------
template <int IPT>
__device__ __forceinline__ void sort_bucket_store(const Col& c, int off, int len, unsigned char* smem_raw) {
  using Sort = cub::BlockMergeSort<KV, kSortThreads, IPT>;
  typename Sort::TempStorage& tmp = *reinterpret_cast<typename Sort::TempStorage*>(smem_raw);
  KV key[IPT];
  const KV oob{d_inf(), 0x7fffffff, 0};
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int g = threadIdx.x * IPT + i;
    key[i] = oob;
    if (g < len) { key[i].v = c.st_val[off + g]; key[i].r = c.st_row[off + g]; }
  }
  Sort(tmp).Sort(key, KVLess(), len, oob);
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int g = threadIdx.x * IPT + i;
    if (g < len) { c.sorted[off + g] = key[i].v; c.perm[off + g] = key[i].r; }
  }
}

__global__ void __launch_bounds__(kSortThreads) bucket_sort_kernel(const Col* cols, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int red[64];
  const Col c = cols[blockIdx.y];
  const int b = blockIdx.x;
  const int len = c.count[b];
  int off, soff;
  bucket_offsets(c.count, b, red, off, soff);
  if (threadIdx.x == 0) {
    c.boff[b] = off;
    c.soff[b] = soff;
    if (b == B - 1) { c.boff[B] = off + len; c.soff[B] = soff + ((len + 31) & ~31); }
    if (len > kBucketCap) atomicOr(c.flag, kFlagOverflow);
  }
  if (len > kBucketCap || (*c.flag & (kFlagNonFinite | kFlagNaN))) return;
  if (len > kSortThreads * 4) sort_bucket_store<8>(c, off, len, smem_raw);
  else if (len > 0) sort_bucket_store<4>(c, off, len, smem_raw);
  __syncthreads();
  if (threadIdx.x == 0) {
    // value range of the bucket; an empty bucket takes its upper splitter so that lo / hi stay non-decreasing
    double lo, hi;
    if (len > 0) { lo = c.sorted[off]; hi = c.sorted[off + len - 1]; }
    else { lo = hi = (b < B - 1) ? c.split[b] : (B > 1 ? c.split[B - 2] : 0.0); }
    c.lo[b] = lo;
    c.hi[b] = hi;
  }
}

// ---- layout: the rows of every x-bucket in ascending y -------------------------------------------------------------
template <int IPT>
__device__ __forceinline__ void layout_bucket(const Col& cx, const Col& cy, const Prob& pr, int off, int soff, int len,
                                              unsigned char* smem_raw) {
  using Sort = cub::BlockMergeSort<KV, kSortThreads, IPT, double>;
  typename Sort::TempStorage& tmp = *reinterpret_cast<typename Sort::TempStorage*>(smem_raw);
  KV key[IPT];
  double xval[IPT];
  const KV oob{d_inf(), 0x7fffffff, 0};
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int g = threadIdx.x * IPT + i;
    key[i] = oob;
    xval[i] = d_nan();
    if (g < len) {
      const int r = cx.perm[off + g];
      key[i].v = cy.vals[r];
      key[i].r = r;
      xval[i] = cx.sorted[off + g];
    }
  }
  Sort(tmp).Sort(key, xval, KVLess(), len, oob);
  const int padded = (len + 31) & ~31;
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const int g = threadIdx.x * IPT + i;
    if (g < len) {
      pr.px[soff + g] = xval[i];
      pr.py[soff + g] = key[i].v;
      pr.slot_row[soff + g] = key[i].r;
    } else if (g < padded) {
      pr.px[soff + g] = d_nan();
      pr.py[soff + g] = d_nan();
      pr.slot_row[soff + g] = -1;
    }
  }
}

__global__ void __launch_bounds__(kSortThreads) layout_kernel(const Col* cols, const Prob* probs, int B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  const int b = blockIdx.x;
  const int len = cx.count[b], off = cx.boff[b], soff = cx.soff[b];
  if (b == 0 && threadIdx.x == 0) *pr.left_count = 0u;
  if (len > kSortThreads * 4) layout_bucket<8>(cx, cy, pr, off, soff, len, smem_raw);
  else if (len > 0) layout_bucket<4>(cx, cy, pr, off, soff, len, smem_raw);
}

// blocks [blo, bhi) of `nblocks` that belong to the shard: block b maps to row floor(b * n / nblocks)
__device__ __forceinline__ void shard_blocks(const Shard& sh, int nblocks, int& blo, int& bhi) {
  // smallest b with floor(b * n / nblocks) >= row  <=>  b >= ceil(row * nblocks / n)
  blo = (int)((sh.row_lo * (long long)nblocks + sh.n - 1) / sh.n);
  bhi = (int)((sh.row_hi * (long long)nblocks + sh.n - 1) / sh.n);
  if (sh.row_hi >= sh.n) bhi = nblocks;
  if (blo > nblocks) blo = nblocks;
  if (bhi > nblocks) bhi = nblocks;
}

// Every lane that wants the chunk walks ITS OWN window of the chunk's ascending y (exact predicates on rounded
// differences: monotone in the slot index); the warp-wide bracket of those windows is staged through shared memory in
// pieces.  Slots within kSeed of own_rel were examined when the list was seeded and must not enter it twice.
template <int K1T>
__device__ __forceinline__ void scan_chunk(int len, const double* __restrict__ gx, const double* __restrict__ gy, bool want,
                                           int own_rel, double qx, double qy, double (&best)[K1T], double& thr,
                                           unsigned long long& np, double* sx, double* sy, int lane) {
  if (len == 0) return;
  const double kInf = d_inf();
  const double t = warp_max(want ? thr : 0.0);
  int a = 0, b = len;
  if (t < kInf) {
    const double y0 = warp_min(want ? qy : kInf), y1 = warp_max(want ? qy : -kInf);
    const double lo_v = (y0 - t) - (fabs(y0) + t) * kSlack;
    const double hi_v = (y1 + t) + (fabs(y1) + t) * kSlack;
    a = warp_first_true(0, len, [&](int s) { return !(gy[s] < lo_v); });
    b = warp_first_true(a, len, [&](int s) { return gy[s] > hi_v; });
  }
  for (int p0 = a; p0 < b; p0 += kPiece) {
    const int pl = min(kPiece, b - p0);
    for (int u = lane; u < pl; u += 32) { sx[u] = gx[p0 + u]; sy[u] = gy[p0 + u]; }
    __syncwarp();
    if (want) {
      int lo = 0, hi = pl;
      const double t0 = thr;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((qy - sy[mid]) < t0) hi = mid; else lo = mid + 1;
      }
#pragma unroll 1
      for (int s = lo; s < pl; ++s) {
        const double yv = sy[s];
        if (!((yv - qy) < thr)) break;
        const double xv = sx[s];
        ++np;
        if (fabs(qx - xv) < thr && fabs(qy - yv) < thr) {
          const int rel = p0 + s - own_rel;
          if (rel < -kSeed || rel > kSeed) {
            topk_insert<K1T>(best, fmax(fabs(qx - xv), fabs(qy - yv)));
            thr = best[K1T - 1];
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---- knn: one warp = 32 y-neighbours of one chunk --------------------------------------------------------------------
template <int K1T>
__global__ void __launch_bounds__(kWarps * 32) knn_kernel2(const Col* cols, const Prob* probs, int B, int k, int defer_below,
                                                           int near, unsigned left_cap, const Shard sh) {
  __shared__ int s_soff[kMaxBuckets + 1];
  __shared__ __align__(16) double s_stage[kWarps][2][kPiece];
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  for (int b = threadIdx.x; b <= B; b += blockDim.x) s_soff[b] = cx.soff[b];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = s_soff[B];
  const int nblocks = (total + kBlockSlots - 1) / kBlockSlots;
  int blo, bhi;
  shard_blocks(sh, nblocks, blo, bhi);
  // work items: the first and last two warps of every chunk first (they sit in the sparse ends of y and search the
  // widest windows), then everything else in slot order
  const int item = blockIdx.x * kWarps + warp;
  int c, wi;
  if (item < 4 * B) {
    c = item >> 2;
    const int e = item & 3;
    const int nw = (s_soff[c + 1] - s_soff[c]) >> 5;
    wi = e < 2 ? e : nw - 1 - (e - 2);
    if (e < 2 ? (e >= nw) : (wi < 2)) return;
  } else {
    const int g = item - 4 * B;
    if (g >= (total >> 5)) return;
    const int s0 = g << 5;
    int lo = 0, hi = B;              // last chunk whose first slot is <= s0 (empty chunks share a first slot: take the last)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_soff[mid] <= s0) lo = mid; else hi = mid - 1;
    }
    c = lo;
    const int nw = (s_soff[c + 1] - s_soff[c]) >> 5;
    wi = (s0 - s_soff[c]) >> 5;
    if (wi < 2 || wi >= nw - 2) return;
  }
  const int base = s_soff[c];
  const int own = wi * 32 + lane;              // chunk-relative slot of this lane's query
  const int slot = base + own;
  if (slot < blo * kBlockSlots || slot >= bhi * kBlockSlots) return;     // (uniform over the warp: 32 | kBlockSlots)
  const int len_home = cx.count[c];
  const bool valid = own < len_home;
  const double qx = pr.px[slot], qy = pr.py[slot];      // NaN in padding slots: every test fails
  const double kInf = d_inf();
  double best[K1T];
#pragma unroll
  for (int t = 0; t < K1T; ++t) best[t] = kInf;
  double thr = kInf;
  unsigned long long np = 0;
  double* sx = s_stage[warp][0];
  double* sy = s_stage[warp][1];

  // 1. seed the list from the query's own neighbourhood in y (coalesced: lane l reads slot own + o)
#pragma unroll 1
  for (int o = -kSeed; o <= kSeed; ++o) {
    const int s = own + o;
    if (s >= 0 && s < len_home) {
      const double xv = pr.px[base + s], yv = pr.py[base + s];
      if (fabs(qx - xv) < thr && fabs(qy - yv) < thr) {
        topk_insert<K1T>(best, fmax(fabs(qx - xv), fabs(qy - yv)));
        thr = best[K1T - 1];
      }
    }
  }
  np += 2 * kSeed + 1;

  // 2. the rest of the home chunk
  scan_chunk<K1T>(cx.count[c], pr.px + base, pr.py + base, valid, own, qx, qy, best, thr, np, sx, sy, lane);
  // 3. outwards over the chunks: a query needs chunk j only while the gap to the chunk's x range is below its
  //    current k-th distance (rounded subtraction is monotone => exact)
  int rstart = B, lend = 0;
  int my_e = -1;                 // this lane's entry in the deferral list, once reserved
  bool may_defer = defer_below > 0;
  // the lanes of `m` hand the rest of a direction to the leftover kernel: entries are reserved here (the list has
  // room for left_cap of them; when it is full the warp simply keeps searching)
  auto reserve = [&](unsigned m) -> bool {
    const unsigned fresh = __ballot_sync(kFull, ((m >> lane) & 1u) && my_e < 0);
    if (fresh == 0) return true;
    unsigned base_e = 0;
    if (lane == 0) base_e = atomicAdd(pr.left_count, (unsigned)__popc(fresh));
    base_e = __shfl_sync(kFull, base_e, 0);
    if (base_e + __popc(fresh) > left_cap) {
      if (lane == 0) atomicSub(pr.left_count, (unsigned)__popc(fresh));
      return false;
    }
    if ((fresh >> lane) & 1u) my_e = (int)(base_e + __popc(fresh & ((1u << lane) - 1u)));
    return true;
  };
  for (int j = c + 1; j < B; ++j) {
    const double cmin = cx.lo[j];
    const bool need = valid && !((cmin - qx) >= thr);
    const unsigned m = __ballot_sync(kFull, need);
    if (m == 0) break;
    if (may_defer && j - c > near && __popc(m) < defer_below) {
      if (reserve(m)) {
        if (need) rstart = j;
        break;
      }
      may_defer = false;
    }
    scan_chunk<K1T>(cx.count[j], pr.px + s_soff[j], pr.py + s_soff[j], need, -(1 << 28), qx, qy, best, thr, np, sx, sy, lane);
  }
  for (int j = c - 1; j >= 0; --j) {
    const double cmax = cx.hi[j];
    const bool need = valid && !((qx - cmax) >= thr);
    const unsigned m = __ballot_sync(kFull, need);
    if (m == 0) break;
    if (may_defer && c - j > near && __popc(m) < defer_below) {
      if (reserve(m)) {
        if (need) lend = j + 1;
        break;
      }
      may_defer = false;
    }
    scan_chunk<K1T>(cx.count[j], pr.px + s_soff[j], pr.py + s_soff[j], need, -(1 << 28), qx, qy, best, thr, np, sx, sy, lane);
  }
  if (valid) {
    double r = best[0];
#pragma unroll
    for (int t = 1; t < K1T; ++t) r = (t <= k) ? best[t] : r;       // = best[k] (a chain of selects, no indexed access)
    pr.eps[slot] = r;
    if (my_e >= 0) {
      const unsigned e = (unsigned)my_e;
      LeftEnt le;
      le.slot = slot; le.rstart = rstart; le.lend = lend;
      pr.left[e] = le;
#pragma unroll
      for (int t = 0; t < K1T; ++t) pr.left_best[(long long)e * K1T + t] = best[t];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np += __shfl_down_sync(kFull, np, o);
  if (lane == 0 && np) atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 4), np);
}

// ---- leftover: one warp finishes one deferred query ------------------------------------------------------------------
template <int K1T>
__global__ void __launch_bounds__(kWarps * 32) leftover_kernel2(const Col* cols, const Prob* probs, int B, int k, long long n,
                                                                int ypath_min_chunks, int ypath_cost) {
  const Prob pr = probs[blockIdx.y];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nent = *pr.left_count;
  const double kInf = d_inf();
  const int k1 = k + 1;
  unsigned long long np = 0;
  for (unsigned e = blockIdx.x * kWarps + warp; e < nent; e += gridDim.x * kWarps) {
    const LeftEnt le = pr.left[e];
    const double qx = pr.px[le.slot], qy = pr.py[le.slot];
    const double gate = pr.left_best[(long long)e * K1T + (K1T - 1)];      // current k-th distance: upper bound of eps
    double best[K1T];
#pragma unroll
    for (int t = 0; t < K1T; ++t) best[t] = kInf;
    double thr = gate;
    // chunks that can still hold a neighbour: gap in x below the gate (monotone => exact)
    int r_lo = B, r_hi = B, l_lo = 0, l_hi = 0;
    if (le.rstart < B) {
      r_lo = le.rstart;
      r_hi = warp_first_true(r_lo, B, [&](int j) { return (cx.lo[j] - qx) >= gate; });
    }
    if (le.lend > 0) {
      l_hi = le.lend;
      l_lo = warp_first_true(0, l_hi, [&](int j) { return !((qx - cx.hi[j]) >= gate); });
    }
    const int nr = r_hi - r_lo, nl = l_hi - l_lo;
    const double lo_v = (qy - gate) - (fabs(qy) + gate) * kSlack;
    const double hi_v = (qy + gate) + (fabs(qy) + gate) * kSlack;
    bool by_y = false;
    int ya = 0, yb = 0;
    if (nr + nl >= ypath_min_chunks && gate < kInf) {
      // a query in a sparse region of y: the rows inside its y window are few - take them from the y-sorted column
      // instead of searching a window in each of many chunks
      ya = warp_first_true(0, (int)n, [&](int t) { return !(cy.sorted[t] < lo_v); });
      yb = warp_first_true(ya, (int)n, [&](int t) { return cy.sorted[t] > hi_v; });
      by_y = (long long)(yb - ya) < (long long)(nr + nl) * ypath_cost;
    }
    if (by_y) {
      for (int t = ya + lane; t < yb; t += 32) {
        const int row = cy.perm[t];
        const int cb = cx.bid[row];
        if (cb >= le.lend && cb < le.rstart) continue;        // chunks the search kernel has already examined
        const double yv = cy.sorted[t], xv = cx.vals[row];
        const double m = fmax(fabs(qx - xv), fabs(qy - yv));
        if (m < thr) { topk_insert<K1T>(best, m); thr = fmin(best[K1T - 1], gate); }
      }
      if (lane == 0) np += (unsigned long long)(yb - ya);
    } else {
      // batches of 32 chunks: every lane binary-searches the y window of ITS chunk (the searches overlap their L2
      // latency), then the warp walks the non-empty windows together
      for (int b0 = 0; b0 < nr + nl; b0 += 32) {
        const int idx = b0 + lane;
        int wlo = 0, whi = 0, cbase = 0;
        if (idx < nr + nl) {
          const int ch = idx < nr ? r_lo + idx : l_lo + (idx - nr);
          cbase = cx.soff[ch];
          const int lenv = cx.count[ch];
          const double* yrow = pr.py + cbase;
          wlo = lower_bound_ge(yrow, lenv, lo_v);
          whi = upper_bound_gt(yrow, lenv, hi_v);
        }
        unsigned todo = __ballot_sync(kFull, wlo < whi);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int slo = __shfl_sync(kFull, wlo, src), shi = __shfl_sync(kFull, whi, src);
          const int sbase = __shfl_sync(kFull, cbase, src);
          for (int j = slo + lane; j < shi; j += 32) {
            const double xv = pr.px[sbase + j], yv = pr.py[sbase + j];
            const double m = fmax(fabs(qx - xv), fabs(qy - yv));
            if (m < thr) { topk_insert<K1T>(best, m); thr = fmin(best[K1T - 1], gate); }
          }
          if (lane == 0) np += (unsigned long long)(shi - slo);
        }
      }
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
  if (lane == 0 && np) atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 4), np);
}

// ---- marginal counts + digamma terms, per 256-slot block -------------------------------------------------------------
__device__ __forceinline__ double psi_lookup(const double* tab, int tab_n, int c) { return c < tab_n ? tab[c] : psi_ref((double)c); }

__global__ void __launch_bounds__(kBlockSlots) count_psi_kernel(const Col* cols, const Prob* probs, int B, long long n, int nblk,
                                                                 const Shard sh, const double* psi_tab, int tab_n) {
  __shared__ double red[kBlockSlots / 32];
  const Prob pr = probs[blockIdx.z];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  if ((*cx.flag | *cy.flag) != 0) return;
  const int which = blockIdx.y;                      // 0: n_x in the sorted x, 1: n_y in the sorted y
  const int total = cx.soff[B];
  const int nblocks = (total + kBlockSlots - 1) / kBlockSlots;
  int blo, bhi;
  shard_blocks(sh, nblocks, blo, bhi);
  const int blk = blockIdx.x;
  if (blk < blo || blk >= bhi) return;
  const int slot = blk * kBlockSlots + threadIdx.x;
  const int row = slot < total ? pr.slot_row[slot] : -1;
  const bool act = row >= 0;
  const double* s = which ? cy.sorted : cx.sorted;
  const double x = act ? (which ? pr.py[slot] : pr.px[slot]) : 0.0;
  const double e = act ? pr.eps[slot] : 0.0;
  const double r = act ? e - 1e-12 : 0.0;            // _entropy_estimators.py:109
  const int len = (int)n;
  const double kInf = d_inf();
  // the queries of a warp are neighbours in the layout: two warp-uniform 33-way searches bracket the stretch of the
  // sorted column their answers lie in (widened by 2^-50 relative), the exact per-lane searches run inside it
  double xmin = warp_min(act ? x : kInf), xmax = warp_max(act ? x : -kInf), rmax = warp_max(act ? fmax(r, 0.0) : 0.0);
  int wl = 0, wh = len;
  if (xmin <= xmax) {
    const double lo_v = (xmin - rmax) - (fabs(xmin) + rmax) * kSlack;
    const double hi_v = (xmax + rmax) + (fabs(xmax) + rmax) * kSlack;
    wl = warp_first_true(0, len, [&](int j) { return !(s[j] < lo_v); });
    wh = warp_first_true(wl, len, [&](int j) { return s[j] > hi_v; });
  }
  int cnt = 0;
  if (act) {
    // lower bound: first j with fl(x - s_j) <= r (non-increasing in j); upper bound: first j with fl(s_j - x) > r
    int first = wl, fhi = wh, lo = wl, hi = wh;
    while (first < fhi || lo < hi) {
      const int m1 = (first + fhi) >> 1, m2 = (lo + hi) >> 1;
      const double v1 = s[min(m1, len - 1)], v2 = s[min(m2, len - 1)];
      if (first < fhi) { if ((x - v1) <= r) fhi = m1; else first = m1 + 1; }
      if (lo < hi) { if ((v2 - x) > r) hi = m2; else lo = m2 + 1; }
    }
    cnt = max(0, lo - first);
    if (which == 0) {
      if (pr.nx_row) pr.nx_row[row] = cnt;
      if (pr.eps_row) pr.eps_row[row] = e;
    } else if (pr.ny_row) {
      pr.ny_row[row] = cnt;
    }
  }
  double term = 0.0, zero = 0.0;
  if (act) {
    if (cnt == 0) zero = 1.0; else term = psi_lookup(psi_tab, tab_n, cnt);
  }
  const int nact = __syncthreads_count(act);
  term = block_sum<kBlockSlots>(term, red);
  zero = block_sum<kBlockSlots>(zero, red);
  if (threadIdx.x == 0) {
    double* p = pr.partial + ((long long)which * nblk + blk) * 2;
    p[0] = term;
    p[1] = zero;
    if (which == 0 && nact) atomicAdd(reinterpret_cast<unsigned long long*>(pr.out + 6), (unsigned long long)nact);   // rows reduced
  }
}

// ---- fixed-order fold of the block partials ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_kernel(const Col* cols, const Prob* probs, int B, int nblk) {
  __shared__ double red[8];
  const Prob pr = probs[blockIdx.x];
  const Col cx = cols[pr.cx], cy = cols[pr.cy];
  const int fl = *cx.flag | *cy.flag;
  int* oflag = reinterpret_cast<int*>(pr.out + 5);
  if (fl != 0) {
    if (threadIdx.x == 0) atomicOr(oflag, fl);
    return;
  }
  const int total = cx.soff[B];
  const int nblocks = (total + kBlockSlots - 1) / kBlockSlots;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};          // psi(n_x) sum, zeros_x, psi(n_y) sum, zeros_y
  for (int b = threadIdx.x; b < nblocks; b += 256) {
    const double* px = pr.partial + (long long)b * 2;
    const double* py = pr.partial + ((long long)nblk + b) * 2;
    acc[0] += px[0]; acc[1] += px[1]; acc[2] += py[0]; acc[3] += py[1];
  }
  double out[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = block_sum<256>(acc[c], red);
  if (threadIdx.x == 0) {
    pr.out[0] = out[0] + out[2];
    pr.out[1] = out[1];
    pr.out[2] = out[3];
    pr.out[3] = 0.0;
  }
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

Plan make_plan(int64_t n, bool* ok) {
  Plan p;
  p.n = n;
  int64_t B = (n + kBucketMean - 1) / kBucketMean;
  if (B < 1) B = 1;
  bool fits = n >= 2 && n < (int64_t(1) << 30);
  if (B > kMaxBuckets) {
    B = kMaxBuckets;
    if (n > int64_t(kMaxBuckets) * (kBucketMean + kBucketMean / 4)) fits = false;     // buckets would run too full
  }
  p.B = static_cast<int>(B);
  int64_t over = n / B;
  if (over > kOversample) over = kOversample;
  if (over < 1) over = 1;
  p.over = static_cast<int>(over);
  p.smax = (n + 32 * B + kBlockSlots - 1) / kBlockSlots * kBlockSlots;
  p.nblk = static_cast<int>(p.smax / kBlockSlots);
  if (ok) *ok = fits;
  return p;
}

size_t col_bytes(int64_t n) {
  size_t b = 0;
  b += align256(sizeof(double) * n);                 // sorted
  b += align256(sizeof(int) * n);                    // perm
  b += align256(sizeof(unsigned short) * n);         // bid
  b += align256(sizeof(double) * n);                 // st_val
  b += align256(sizeof(int) * n);                    // st_row
  b += align256(sizeof(double) * kMaxBuckets) * 3;   // split, lo, hi
  b += align256(sizeof(int) * (kMaxBuckets + 1)) * 4;  // count, fill, boff, soff
  b += 256;                                          // flag
  return b;
}

Col carve_col(char* base, int64_t n, const double* vals) {
  Col c;
  size_t o = 0;
  auto take = [&](size_t bytes) { char* p = base + o; o += align256(bytes); return p; };
  c.vals = vals;
  c.sorted = reinterpret_cast<double*>(take(sizeof(double) * n));
  c.perm = reinterpret_cast<int*>(take(sizeof(int) * n));
  c.bid = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * n));
  c.st_val = reinterpret_cast<double*>(take(sizeof(double) * n));
  c.st_row = reinterpret_cast<int*>(take(sizeof(int) * n));
  c.split = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.lo = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.hi = reinterpret_cast<double*>(take(sizeof(double) * kMaxBuckets));
  c.count = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.fill = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.boff = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.soff = reinterpret_cast<int*>(take(sizeof(int) * (kMaxBuckets + 1)));
  c.flag = reinterpret_cast<int*>(take(sizeof(int)));
  return c;
}

// entries the deferral list of a problem can hold (a full list only means that the search kernel keeps going itself)
size_t left_cap(const Plan& p) { return static_cast<size_t>(p.n / 8 + 1024); }

size_t prob_bytes(const Plan& p, int k1t) {
  size_t b = 0;
  b += align256(sizeof(double) * p.smax) * 3;                 // px, py, eps
  b += align256(sizeof(int) * p.smax);                        // slot_row
  b += align256(sizeof(LeftEnt) * left_cap(p));               // left
  b += align256(sizeof(double) * left_cap(p) * k1t);          // left_best
  b += align256(sizeof(double) * 4 * p.nblk);                 // partial
  b += 256;                                                   // left_count
  b += 256;                                                   // out
  return b;
}

Prob carve_prob(char* base, const Plan& p, int k1t, int cx, int cy) {
  Prob q;
  size_t o = 0;
  auto take = [&](size_t bytes) { char* ptr = base + o; o += align256(bytes); return ptr; };
  q.cx = cx; q.cy = cy;
  q.px = reinterpret_cast<double*>(take(sizeof(double) * p.smax));
  q.py = reinterpret_cast<double*>(take(sizeof(double) * p.smax));
  q.eps = reinterpret_cast<double*>(take(sizeof(double) * p.smax));
  q.slot_row = reinterpret_cast<int*>(take(sizeof(int) * p.smax));
  q.left = reinterpret_cast<LeftEnt*>(take(sizeof(LeftEnt) * left_cap(p)));
  q.left_best = reinterpret_cast<double*>(take(sizeof(double) * left_cap(p) * k1t));
  q.partial = reinterpret_cast<double*>(take(sizeof(double) * 4 * p.nblk));
  q.left_count = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int)));
  q.out = reinterpret_cast<double*>(take(sizeof(double) * 8));
  q.eps_row = nullptr; q.nx_row = nullptr; q.ny_row = nullptr;
  return q;
}

namespace {
constexpr size_t kSplitSmem = sizeof(double) * (kSplitThreads * kSplitItems + 1) + 64;
constexpr size_t kSortSmem = sizeof(KV) * (kSortThreads * 8 + 1) + 64;
}  // namespace

cudaError_t init() {
  cudaError_t e = cudaFuncSetAttribute(split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplitSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
}

cudaError_t colsort(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches) {
  const int nb = static_cast<int>((p.n + kCountRows - 1) / kCountRows);
  split_kernel<<<ncol, kSplitThreads, kSplitSmem, s>>>(cols, p.n, p.B, p.over);
  bucket_count_kernel<<<dim3(nb, ncol), 256, 0, s>>>(cols, p.n, p.B);
  bucket_scatter_kernel<<<dim3(nb, ncol), 256, 0, s>>>(cols, p.n, p.B);
  bucket_sort_kernel<<<dim3(p.B, ncol), kSortThreads, kSortSmem, s>>>(cols, p.B);
  if (launches) *launches += 4;
  return cudaGetLastError();
}

cudaError_t layout(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches) {
  layout_kernel<<<dim3(p.B, nprob), kSortThreads, kSortSmem, s>>>(cols, probs, p.B);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t knn(const Col* cols, const Prob* probs, int nprob, const Plan& p, int k, const Shard& sh, int sm_count,
                cudaStream_t s, int* launches) {
  int defer_below = 8, near = 1, ymin = 8, ycost = 24;
  if (const char* e = getenv("EB2_K2_DEFER")) defer_below = atoi(e);       // tuning knobs
  if (const char* e = getenv("EB2_K2_NEAR")) near = atoi(e);
  if (const char* e = getenv("EB2_K2_YMIN")) ymin = atoi(e);
  if (const char* e = getenv("EB2_K2_YCOST")) ycost = atoi(e);
  const int items = 4 * p.B + static_cast<int>(p.smax / 32);
  const int grid = (items + kWarps - 1) / kWarps;
  const int lgrid = nprob > 1 ? std::max(1, sm_count * 4 / nprob) : sm_count * 8;
  if (k + 1 <= 4) {
    knn_kernel2<4><<<dim3(grid, nprob), kWarps * 32, 0, s>>>(cols, probs, p.B, k, defer_below, near, (unsigned)left_cap(p), sh);
    leftover_kernel2<4><<<dim3(lgrid, nprob), kWarps * 32, 0, s>>>(cols, probs, p.B, k, p.n, ymin, ycost);
  } else if (k + 1 <= 8) {
    knn_kernel2<8><<<dim3(grid, nprob), kWarps * 32, 0, s>>>(cols, probs, p.B, k, defer_below, near, (unsigned)left_cap(p), sh);
    leftover_kernel2<8><<<dim3(lgrid, nprob), kWarps * 32, 0, s>>>(cols, probs, p.B, k, p.n, ymin, ycost);
  } else {
    return cudaErrorInvalidValue;
  }
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t count_psi(const Col* cols, const Prob* probs, int nprob, const Plan& p, const Shard& sh, const double* psi_tab,
                      int tab_n, cudaStream_t s, int* launches) {
  count_psi_kernel<<<dim3(p.nblk, 2, nprob), kBlockSlots, 0, s>>>(cols, probs, p.B, p.n, p.nblk, sh, psi_tab, tab_n);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t finalize(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches) {
  final_kernel<<<nprob, 256, 0, s>>>(cols, probs, p.B, p.nblk);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace k2
}  // namespace eb2
