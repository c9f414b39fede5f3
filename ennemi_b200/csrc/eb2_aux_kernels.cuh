// ennemi_b200 — non-templated kernels: 1-D marginal search, digamma reduction, layout helpers.
// Included by eb2_lib.cu only (one definition per library).
#pragma once
#include "eb2_kernels.cuh"

namespace eb2 {

// ----------------------------------------------------------------------------------------------
// (2a) 1-D marginal counts: binary search in an ascending coordinate array with the EXACT
//      predicate |fl(x_i - s_j)| <= r_i (rounded subtraction is monotone in s_j, so the predicate
//      is monotone on each side of x_i and the count equals what the all-pairs test would give).
// ----------------------------------------------------------------------------------------------
struct SearchArgs {
  const double* qcoord;   // query coordinate per query slot
  const double* radius;   // per query slot
  int from_eps;           // `radius` holds the k-th neighbour distances: the radius is fl(eps - 1e-12), taken here
  const double* sorted;   // ascending candidate coordinates
  // optional second marginal searched by the same launch (blockIdx.y == 1)
  const double* qcoord2;
  const double* sorted2;
  int* cnt2;
  const Tile* tiles;      // q_lo/q_n: query slots; c_lo/c_len: slice of `sorted` to search
  int ntiles;
  int* cnt;               // out per query slot
};

__global__ void __launch_bounds__(kThreads) search_kernel(const SearchArgs a) {
  const bool second = blockIdx.y == 1;
  const double* qcoord = second ? a.qcoord2 : a.qcoord;
  const double* sorted = second ? a.sorted2 : a.sorted;
  int* cnt = second ? a.cnt2 : a.cnt;
  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    const double* s = sorted + tile.c_lo;
    for (int base = 0; base < tile.q_n; base += kThreads) {
      const int qi = base + threadIdx.x;
      const bool act = qi < tile.q_n;
      const int slot = tile.q_lo + (act ? qi : 0);
      const double x = act ? qcoord[slot] : 0.0;
      const double r = act ? (a.from_eps ? a.radius[slot] - 1e-12 : a.radius[slot]) : 0.0;   // _entropy_estimators.py:109
      // The queries of a warp are neighbours in the layout (same chunk of the across-chunk coordinate, consecutive in
      // the in-chunk one), so their answers lie in a short stretch of `s`: two warp-uniform searches (broadcast loads)
      // bracket it conservatively, the per-lane searches with the exact predicates run inside the bracket only.
      const double kInf = __longlong_as_double(0x7ff0000000000000LL);
      double xmin = act ? x : kInf, xmax = act ? x : -kInf, rmax = act ? fmax(r, 0.0) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      }
      int wl = 0, wh = tile.c_len;
      if (xmin <= xmax) {
        const double slack = 8.881784197001252e-16;  // 2^-50: the rounded per-lane tests can never disagree with the bracket
        const double lo_v = (xmin - rmax) - (fabs(xmin) + rmax) * slack;
        const double hi_v = (xmax + rmax) + (fabs(xmax) + rmax) * slack;
        // 33-way searches: 32 probes per round, ~4 dependent rounds for 10^6 values instead of 20 bisection steps
        wl = warp_first_true(0, tile.c_len, [&](int j) { return !(s[j] < lo_v); });
        wh = warp_first_true(wl, tile.c_len, [&](int j) { return s[j] > hi_v; });
      }
      if (!act) continue;
      // lower bound: first j with fl(x - s_j) <= r   (x - s_j is non-increasing in j)
      // upper bound: first j with fl(s_j - x) > r    (s_j - x is non-decreasing in j)
      // both bisections advance together so that their loads are in flight at the same time
      int first = wl, fhi = wh, lo = wl, hi = wh;
      while (first < fhi || lo < hi) {
        const int m1 = (first + fhi) >> 1, m2 = (lo + hi) >> 1;
        const double v1 = s[min(m1, tile.c_len - 1)], v2 = s[min(m2, tile.c_len - 1)];
        if (first < fhi) { if ((x - v1) <= r) fhi = m1; else first = m1 + 1; }
        if (lo < hi) { if ((v2 - x) > r) hi = m2; else lo = m2 + 1; }
      }
      cnt[slot] = max(0, lo - first);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// (1') k-th neighbour distance in ONE dimension (Ross within a class, :194; 1-D entropy, :39): inside a segment the
//      coordinate ascends, so the k nearest neighbours of a row are among its k predecessors and k successors and
//      their distances ascend with the offset on either side (rounded subtraction is monotone): a two-pointer merge
//      of the two sides yields the k smallest distances in order.  Exactly the values the all-pairs kernel would
//      select (|fl(x_i - x_j)|, self included at 0; inf when the segment holds fewer than k + 1 rows).
// ----------------------------------------------------------------------------------------------
struct Knn1dArgs {
  const double* x;        // the sorted coordinate row of the padded point set
  const Tile* tiles;      // q_lo/q_n: query slots; c_lo/c_len: their segment
  int ntiles;
  int k;
  double* eps;            // out per query slot
  unsigned long long* pairs;
};

__global__ void __launch_bounds__(kThreads) knn1d_kernel(const Knn1dArgs a) {
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    const double* seg = a.x + tile.c_lo;
    for (int qi = threadIdx.x; qi < tile.q_n; qi += kThreads) {
      const int slot = tile.q_lo + qi;
      const int rel = slot - tile.c_lo;
      const double xq = seg[rel];
      int l = rel - 1, r = rel + 1;
      double dl = l >= 0 ? xq - seg[l] : kInf, dr = r < tile.c_len ? seg[r] - xq : kInf;
      double cur = 0.0;                                   // the row itself
      for (int step = 0; step < a.k; ++step) {
        if (dl <= dr) {
          cur = dl;
          --l;
          dl = l >= 0 ? xq - seg[l] : kInf;
        } else {
          cur = dr;
          ++r;
          dr = r < tile.c_len ? seg[r] - xq : kInf;
        }
      }
      a.eps[slot] = cur;
    }
    if (threadIdx.x == 0 && a.pairs) atomicAdd(a.pairs, (unsigned long long)tile.q_n * (unsigned long long)(2 * a.k + 1));
  }
}

// ----------------------------------------------------------------------------------------------
// (3) digamma terms + deterministic reduction
// ----------------------------------------------------------------------------------------------
enum PsiMode { PSI_A = 1, PSI_AB = 2, PSI_AB_MINUS_C = 3, LOG_DIST = 4 };

struct PsiArgs {
  const int* cnt_a;
  const int* cnt_b;
  const int* cnt_c;
  const double* dist;     // LOG_DIST mode
  const Tile* tiles;
  int ntiles;
  int mode;
  double* partial;        // [ntiles][4]: sum, zeros_a, zeros_b, zeros_c
  const double* tab;      // tab[n] = psi_ref(n) for n < tab_n, filled by psi_table_kernel with the same device function
  int tab_n;              // (so a lookup returns the very bits an evaluation would); 0: no table
};

__global__ void psi_table_kernel(double* tab, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tab[i] = i == 0 ? 0.0 : psi_ref((double)i);
}

// digamma of a positive count: most counts are small (the neighbourhood of a point holds a few thousand rows at
// N = 10^6), so the two fp64 logs per row become two cached loads
__device__ __forceinline__ double psi_count(const PsiArgs& a, int n) {
  return n < a.tab_n ? a.tab[n] : psi_ref((double)n);
}

__global__ void __launch_bounds__(kThreads) psi_kernel(const PsiArgs a) {
  __shared__ double red[kThreads / 32];
  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    double sum = 0.0, za = 0.0, zb = 0.0, zc = 0.0;
#pragma unroll
    for (int i = 0; i < kMaxQpt; ++i) {
      const int qi = threadIdx.x + i * kThreads;
      if (qi < tile.q_n) {
        const int slot = tile.q_lo + qi;
        if (a.mode == LOG_DIST) {
          sum += log(a.dist[slot]);
        } else {
          // a zero count makes the reference's _psi return a scalar +inf (:338-339); it is reported
          // through the zero counters and contributes nothing to the finite sum.
          const int na = a.cnt_a[slot];
          double term = 0.0;
          if (na == 0) za += 1.0; else term = psi_count(a, na);
          if (a.mode >= PSI_AB) {
            const int nb = a.cnt_b[slot];
            if (nb == 0) zb += 1.0; else term = term + psi_count(a, nb);
          }
          if (a.mode == PSI_AB_MINUS_C) {
            const int nc = a.cnt_c[slot];
            if (nc == 0) zc += 1.0; else term = term - psi_count(a, nc);
          }
          sum += term;
        }
      }
    }
    sum = block_sum<kThreads>(sum, red);
    za = block_sum<kThreads>(za, red);
    zb = block_sum<kThreads>(zb, red);
    zc = block_sum<kThreads>(zc, red);
    if (threadIdx.x == 0) {
      double* p = a.partial + (int64_t)tile_id * 4;
      p[0] = sum; p[1] = za; p[2] = zb; p[3] = zc;
    }
  }
}

// one CTA folds the per-tile partials in a fixed order: thread t sums tiles t, t+256, ... then the tree
__global__ void __launch_bounds__(kThreads) psi_final_kernel(const double* partial, int ntiles, double* out /*4*/) {
  __shared__ double red[kThreads / 32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int t = threadIdx.x; t < ntiles; t += kThreads) {
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] += partial[(int64_t)t * 4 + c];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double r = block_sum<kThreads>(acc[c], red);
    if (threadIdx.x == 0) out[c] = r;
  }
}

// psi of a plain count array (eb2_psi)
__global__ void psi_array_kernel(const long long* counts, int64_t n, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const long long c = counts[i];
    out[i] = c == 0 ? __longlong_as_double(0x7ff0000000000000LL) : psi_ref((double)c);
  }
}

// ----------------------------------------------------------------------------------------------
// layout helpers
// ----------------------------------------------------------------------------------------------
__global__ void iota_kernel(int* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// rank -> padded slot.  Unsegmented: slot = rank.  Segmented: ranks are class-major, class c owns
// ranks [seg_rank[c], seg_rank[c+1]) and slots starting at seg_slot[c].
struct GatherArgs {
  const double* rows[kMaxDimAny + 1];   // source of each of the d rows (n values, caller's row order)
  int64_t n;
  int d;
  const int* perm;        // rank -> input row (NULL: identity)
  const int* cls_sorted;  // class of each rank (NULL: one segment)
  const int* seg_rank;    // per class
  const int* seg_slot;    // per class
  double* P;              // d x stride, pre-filled with NaN
  int64_t stride;
  int* slot_row;          // slot -> input row, pre-filled with -1
};

__global__ void gather_kernel(const GatherArgs a) {
  const int64_t rank = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= a.n) {
    if (!a.cls_sorted && rank < a.stride) {        // single segment: the padding slots after the last row
      for (int t = 0; t < a.d; ++t) a.P[t * a.stride + rank] = __longlong_as_double(0x7ff8000000000000LL);
      a.slot_row[rank] = -1;
    }
    return;
  }
  const int row = a.perm ? a.perm[rank] : (int)rank;
  int64_t slot = rank;
  if (a.cls_sorted) {
    const int c = a.cls_sorted[rank];
    slot = a.seg_slot[c] + (rank - a.seg_rank[c]);
  }
  for (int t = 0; t < a.d; ++t) a.P[t * a.stride + slot] = a.rows[t][row];
  a.slot_row[slot] = row;
}

// two-level layout by partition (build_point_set_rows): chunk of every row from its rank in coordinate 0 ...
__global__ void row_chunk_kernel(const int* __restrict__ perm0, int n, int tc, int* __restrict__ row_chunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) row_chunk[perm0[i]] = i / tc;
}
// ... read in the order of coordinate 1 (a stable sort of these keys then lists every chunk's rows in that order)
__global__ void chunk_key_kernel(const int* __restrict__ perm1, const int* __restrict__ row_chunk, int n, int* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = row_chunk[perm1[i]];
}
// range of coordinate 0 per chunk, from its ascending values
__global__ void cell_bounds_kernel(const double* __restrict__ keys0, long long n, int tc, int nchunks, double* lo, double* hi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const long long a = (long long)c * tc;
  const long long b = (a + tc < n ? a + tc : n) - 1;
  lo[c] = keys0[a];
  hi[c] = keys0[b];
}

__global__ void gather_int_kernel(const int* src, const int* perm, int* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__global__ void radius_kernel(const double* eps, double* radius, int64_t nslots) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nslots) radius[i] = eps[i] - 1e-12;   // _entropy_estimators.py:109
}

// slot-order results -> caller's row order
__global__ void scatter_f64_kernel(const double* src, const int* slot_row, int64_t nslots, double* dst) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nslots) { const int r = slot_row[s]; if (r >= 0) dst[r] = src[s]; }
}
__global__ void scatter_i64_kernel(const int* src, const int* slot_row, int64_t nslots, long long* dst) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nslots) { const int r = slot_row[s]; if (r >= 0) dst[r] = (long long)src[s]; }
}

// 1 if any of the first `count` values is NaN or +-inf
__global__ void nonfinite_kernel(const double* v, int64_t count, int* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const double x = v[i];
    if (!(fabs(x) < __longlong_as_double(0x7ff0000000000000LL))) atomicOr(flag, 2);
  }
}


// cached columns -> the d x n block the estimators work on.  Rescaling is the reference's
// (xs - xs.mean()) / std ; xs += noise  (ennemi/_driver.py:882-883): three correctly rounded fp64
// operations in that order, nothing fused, so the block is bit-identical to the host-prepared one.
struct PrepCol {
  const double* src;
  long long off, stride;
  double mean, std;        // std == 0: pass the values through unchanged (no rescaling, no noise)
  const double* noise;     // NULL: no noise
  long long noff, nstride;
  const double* dstats;    // non-NULL (EB2_FLAG_DEVICE_STATS): mean = dstats[1], std = dstats[3], computed in this call
  int* flag;               // non-NULL: this column's own flag word (instead of PrepArgs::flags)
  double* dst;             // non-NULL: this column's own destination (instead of raw + t * n)
};
struct PrepArgs {
  PrepCol col[kMaxDimAny];
  int d;
  long long n;
  double* raw;
  int* flags;              // bit0: NaN among the input values, bit3: a window with device statistics is constant
};

__global__ void prep_kernel(const PrepArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n * a.d) return;
  const int t = (int)(idx / a.n);
  const long long i = idx - (long long)t * a.n;
  const PrepCol& c = a.col[t];
  double v = c.src[c.off + i * c.stride];
  int* flags = c.flag ? c.flag : a.flags;
  if (v != v) atomicOr(flags, 1);
  double mean = c.mean, std = c.std;
  if (c.dstats && std != 0.0) {
    mean = c.dstats[1];
    std = c.dstats[3];
    if (fabs(std) < 1e-20) {       // ennemi/_driver.py:879: the caller has to take the warning path
      if (i == 0) atomicOr(flags, 8);
      std = 0.0;
    }
  }
  if (std != 0.0) {
    v = __ddiv_rn(__dsub_rn(v, mean), std);
    if (c.noise) v = __dadd_rn(v, c.noise[c.noff + i * c.nstride]);
  }
  if (c.dst) c.dst[i] = v; else a.raw[idx] = v;
}

// ---- NumPy-exact mean / standard deviation of a cached column window ------------------------------
// np.add.reduce on float64 is a fixed pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum): blocks of <= 128 elements are summed with 8 interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; larger ranges split at n/2 rounded down to a
// multiple of 8, recursively.  Reproducing that association exactly gives the same bits as ndarray.mean()
// / ndarray.std() (tests/test_gpu_api.py checks it), so the statistics the rescaling uses can be computed
// where the data already is.  The host supplies the leaf table and the tree of the recursion for this n.
struct NpLeaf {
  long long off;   // first element of the leaf (window-relative)
  int len;         // <= 128
};

// Statistics of up to kNpCols windows of one length per launch (blockIdx.y = window).
constexpr int kNpCols = 8;
struct NpCols {
  const double* src[kNpCols];
  long long off[kNpCols], stride[kNpCols];
  double* out[kNpCols];      // 4 doubles per window: sum, mean, sum of squared deviations, std
  double* val[kNpCols];      // nleaves + ninner node values of the summation tree
};

// mode 0: leaf sums of x, mode 1: of (x - mean)^2 with the mean read from out[1].
// Eight lanes per leaf: lane j owns NumPy's accumulator r_j, so a leaf's 64-byte groups are read coalesced
// and the shuffle tree below is exactly ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).
__global__ void np_leaf_sum_kernel(const NpCols a, const NpLeaf* leaves, int nleaves, int mode) {
  const int w = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = g >> 3, j = g & 7;
  const bool live = t < nleaves;
  NpLeaf lf{0, 0};
  if (live) lf = leaves[t];
  const double mean = mode ? a.out[w][1] : 0.0;
  const long long stride = a.stride[w];
  const double* p = a.src[w] + a.off[w] + lf.off * stride;
  auto at = [&](int i) {
    const double v = p[(long long)i * stride];
    if (mode == 0) return v;
    const double c = __dsub_rn(v, mean);
    return __dmul_rn(c, c);
  };
  const int n = lf.len;
  const int body = n - (n % 8);
  double r = 0.0;
  if (n >= 8) {
    r = at(j);
    for (int i = 8; i < body; i += 8) r = __dadd_rn(r, at(i + j));
  }
  // every lane of the warp takes part in the shuffles (leaves of different lengths share a warp)
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));     // r0+r1, r2+r3, r4+r5, r6+r7
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));     // lane 0 of the group: NumPy's association
  double res = 0.0;
  if (j == 0) {
    if (n >= 8) res = r;
    for (int i = (n >= 8 ? body : 0); i < n; ++i) res = __dadd_rn(res, at(i));
  }
  if (live && j == 0) a.val[w][t] = res;
}

// The recursion above the leaves, bottom-up: inner node i (value slot nleaves + i) adds its two children;
// nodes are ordered by height and level_start[h] .. level_start[h+1] are the nodes of height h+1, which only
// depend on lower ones.  One CTA per window, a barrier per height (~log2(n/128) + 1 of them).
constexpr int kNpMaxLevels = 40;
struct NpLevels {
  int start[kNpMaxLevels + 1];
  int nlevels;
};
// finish: 0 -> out[0] = total, out[1] = total / n (the mean); 1 -> out[2] = total, out[3] = sqrt(total / n) (the std)
__global__ void np_tree_kernel(const NpCols a, const int2* children, const NpLevels lv, int nleaves, long long n, int finish) {
  double* val = a.val[blockIdx.x];
  for (int h = 0; h < lv.nlevels; ++h) {
    for (int i = lv.start[h] + threadIdx.x; i < lv.start[h + 1]; i += blockDim.x) {
      const int2 c = children[i];
      val[nleaves + i] = __dadd_rn(val[c.x], val[c.y]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int ninner = lv.start[lv.nlevels];
    const double total = val[nleaves + ninner - 1];        // the root is the last node (or the only leaf)
    const double q = __ddiv_rn(total, (double)n);
    double* out = a.out[blockIdx.x] + 2 * finish;
    out[0] = total;
    out[1] = finish ? __dsqrt_rn(q) : q;
  }
}

// (n x ncols) row-major block -> ncols separate columns (eb2_cache_put_block): 32 x 32 tiles through shared
// memory so that both the reads (along a row of the block) and the writes (along a column) are coalesced
__global__ void deinterleave_kernel(const double* __restrict__ block, long long n, int ncols, long long ld, double* const* __restrict__ cols) {
  __shared__ double tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i;
    const int cc = c0 + threadIdx.x;
    if (r < n && cc < ncols) tile[i][threadIdx.x] = block[r * ld + cc];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int cc = c0 + j;
    const long long r = r0 + threadIdx.x;
    if (cc < ncols && r < n) cols[cc][r] = tile[threadIdx.x][j];
  }
}

__global__ void or_flags_kernel(int* dst, const int* a, const int* b) {
  const int v = (a ? *a : 0) | (b ? *b : 0);
  if (v) atomicOr(dst, v);
}

// register-resident DADD chains: the FP64 issue rate that bounds the all-pairs kernels
__global__ void fp64_peak_kernel(double* out, double c, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int it = 0; it < iters; ++it) {
    a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
    a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---- generic run-time-dimension brute-force kernels (spaces wider than kMaxDim) ----------------
__global__ void __launch_bounds__(kThreads) knn_generic_kernel(const GenKnnArgs a) {
  __shared__ double sbuf[kMaxDimAny * kGenChunk];
  const int tid = threadIdx.x;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    const bool valid = tid < tile.q_n;
    double q[kMaxDimAny];
    for (int t = 0; t < a.d; ++t) q[t] = valid ? a.P[a.rows.row[t] * a.stride + tile.q_lo + tid] : kNaN;
    HeapRef heap;
    heap.nq = (int64_t)gridDim.x * kThreads;
    heap.base = a.heap + (int64_t)blockIdx.x * kThreads + tid;
    heap.k1 = a.k + 1;
    heap.fill_inf();
    double thr = kInf;
    const int len_pad = (tile.c_len + kSegAlign - 1) / kSegAlign * kSegAlign;
    for (int c0 = 0; c0 < len_pad; c0 += kGenChunk) {
      const int len = min(kGenChunk, len_pad - c0);
      for (int idx = tid; idx < a.d * len; idx += kThreads) {
        const int t = idx / len, j = idx - t * len;
        sbuf[t * kGenChunk + j] = a.P[a.rows.row[t] * a.stride + tile.c_lo + c0 + j];
      }
      __syncthreads();
      for (int j = 0; j < len; ++j) {
        bool in = true;
        double m = 0.0;
        for (int t = 0; t < a.d && in; ++t) {
          const double v = fabs(q[t] - sbuf[t * kGenChunk + j]);
          in = v < thr;                  // NaN padding fails here
          m = v > m ? v : m;
        }
        if (in) thr = heap.replace_root(m);
      }
      __syncthreads();
    }
    if (valid) a.eps[tile.q_lo + tid] = thr;
    if (tid == 0 && a.pairs) atomicAdd(a.pairs, (unsigned long long)tile.c_len * tile.q_n);
  }
}

__global__ void __launch_bounds__(kThreads) count_generic_kernel(const GenCountArgs a) {
  __shared__ double sbuf[(kMaxDimAny + 2) * kGenChunk];
  const int tid = threadIdx.x;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  const int D = a.C + a.E;
  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    const bool valid = tid < tile.q_n;
    const int slot = tile.q_lo + tid;
    double qs[kMaxDimAny];
    double qe[2] = {kNaN, kNaN};
    for (int t = 0; t < a.C; ++t) qs[t] = valid ? a.Q[a.q_srow.row[t] * a.qstride + slot] : kNaN;
    for (int t = 0; t < a.E; ++t) qe[t] = valid ? a.Q[a.q_erow.row[t] * a.qstride + slot] : kNaN;
    const double r = valid ? a.radius[slot] : kNaN;
    int ns = 0, ne0 = 0, ne1 = 0;
    const int len_pad = (tile.c_len + kSegAlign - 1) / kSegAlign * kSegAlign;
    for (int c0 = 0; c0 < len_pad; c0 += kGenChunk) {
      const int len = min(kGenChunk, len_pad - c0);
      for (int idx = tid; idx < D * len; idx += kThreads) {
        const int t = idx / len, j = idx - t * len;
        const int row = t < a.C ? a.b_srow.row[t] : a.b_erow.row[t - a.C];
        sbuf[t * kGenChunk + j] = a.B[row * a.bstride + tile.c_lo + c0 + j];
      }
      __syncthreads();
      for (int j = 0; j < len; ++j) {
        bool in = true;
        for (int t = 0; t < a.C && in; ++t) in = fabs(qs[t] - sbuf[t * kGenChunk + j]) <= r;
        if (a.C > 0) {
          if (in) {
            ++ns;
            if (a.E > 0) ne0 += (int)(fabs(qe[0] - sbuf[a.C * kGenChunk + j]) <= r);
            if (a.E > 1) ne1 += (int)(fabs(qe[1] - sbuf[(a.C + 1) * kGenChunk + j]) <= r);
          }
        } else {
          if (a.E > 0) ne0 += (int)(fabs(qe[0] - sbuf[j]) <= r);
          if (a.E > 1) ne1 += (int)(fabs(qe[1] - sbuf[kGenChunk + j]) <= r);
        }
      }
      __syncthreads();
    }
    if (valid) {
      if (a.C > 0) a.cnt_s[slot] = ns;
      if (a.E > 0) a.cnt_e0[slot] = ne0;
      if (a.E > 1) a.cnt_e1[slot] = ne1;
    }
    if (tid == 0 && a.pairs) atomicAdd(a.pairs, (unsigned long long)tile.c_len * tile.q_n);
  }
}

}  // namespace eb2
