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
    for (int qi = threadIdx.x; qi < tile.q_n; qi += kThreads) {
      const int slot = tile.q_lo + qi;
      const double x = qcoord[slot];
      const double r = a.radius[slot];
      // first j with fl(x - s_j) <= r   (x - s_j is non-increasing in j)
      int lo = 0, hi = tile.c_len;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((x - s[mid]) <= r) hi = mid; else lo = mid + 1;
      }
      const int first = lo;
      // first j with fl(s_j - x) > r    (s_j - x is non-decreasing in j)
      lo = 0; hi = tile.c_len;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((s[mid] - x) > r) hi = mid; else lo = mid + 1;
      }
      cnt[slot] = max(0, lo - first);
    }
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
};

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
          if (na == 0) za += 1.0; else term = psi_ref((double)na);
          if (a.mode >= PSI_AB) {
            const int nb = a.cnt_b[slot];
            if (nb == 0) zb += 1.0; else term = term + psi_ref((double)nb);
          }
          if (a.mode == PSI_AB_MINUS_C) {
            const int nc = a.cnt_c[slot];
            if (nc == 0) zc += 1.0; else term = term - psi_ref((double)nc);
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
  if (rank >= a.n) return;
  const int row = a.perm ? a.perm[rank] : (int)rank;
  int64_t slot = rank;
  if (a.cls_sorted) {
    const int c = a.cls_sorted[rank];
    slot = a.seg_slot[c] + (rank - a.seg_rank[c]);
  }
  for (int t = 0; t < a.d; ++t) a.P[t * a.stride + slot] = a.rows[t][row];
  a.slot_row[slot] = row;
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
};
struct PrepArgs {
  PrepCol col[kMaxDimAny];
  int d;
  long long n;
  double* raw;
  int* flags;              // bit0: NaN among the input values
};

__global__ void prep_kernel(const PrepArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n * a.d) return;
  const int t = (int)(idx / a.n);
  const long long i = idx - (long long)t * a.n;
  const PrepCol& c = a.col[t];
  double v = c.src[c.off + i * c.stride];
  if (v != v) atomicOr(a.flags, 1);
  if (c.std != 0.0) {
    v = __ddiv_rn(__dsub_rn(v, c.mean), c.std);
    if (c.noise) v = __dadd_rn(v, c.noise[c.noff + i * c.nstride]);
  }
  a.raw[idx] = v;
}

// ---- NumPy-exact mean / standard deviation of a cached column window ------------------------------
// np.add.reduce on float64 is a fixed pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum): blocks of <= 128 elements are summed with 8 interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; larger ranges split at n/2 rounded down to a
// multiple of 8, recursively.  Reproducing that association exactly gives the same bits as ndarray.mean()
// / ndarray.std() (tests/test_gpu_api.py checks it), so the statistics the rescaling uses can be computed
// where the data already is.  The host supplies the leaf table of the recursion for this n.
struct NpLeaf {
  long long off;   // first element of the leaf (window-relative)
  int len;         // <= 128
};

// one thread per leaf; mode 0: sum of x, mode 1: sum of (x - mean)^2 with mean read from *mean_ptr
__global__ void np_leaf_sum_kernel(const double* src, long long off, long long stride, const NpLeaf* leaves, int nleaves,
                                   int mode, const double* mean_ptr, double* leaf_sum) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nleaves) return;
  const NpLeaf lf = leaves[t];
  const double mean = mode ? *mean_ptr : 0.0;
  const double* p = src + off + lf.off * stride;
  auto at = [&](int i) {
    const double v = p[(long long)i * stride];
    if (mode == 0) return v;
    const double c = __dsub_rn(v, mean);
    return __dmul_rn(c, c);
  };
  const int n = lf.len;
  double res;
  if (n < 8) {
    res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, at(i));
  } else {
    double r0 = at(0), r1 = at(1), r2 = at(2), r3 = at(3), r4 = at(4), r5 = at(5), r6 = at(6), r7 = at(7);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
      r0 = __dadd_rn(r0, at(i)); r1 = __dadd_rn(r1, at(i + 1)); r2 = __dadd_rn(r2, at(i + 2)); r3 = __dadd_rn(r3, at(i + 3));
      r4 = __dadd_rn(r4, at(i + 4)); r5 = __dadd_rn(r5, at(i + 5)); r6 = __dadd_rn(r6, at(i + 6)); r7 = __dadd_rn(r7, at(i + 7));
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, at(i));
  }
  leaf_sum[t] = res;
}

// the recursion above the leaves; `next` walks the leaf sums in order
__device__ double np_combine(const double* leaf_sum, int& next, long long n) {
  if (n <= 128) return leaf_sum[next++];
  long long n2 = n / 2;
  n2 -= n2 % 8;
  const double a = np_combine(leaf_sum, next, n2);
  const double b = np_combine(leaf_sum, next, n - n2);
  return __dadd_rn(a, b);
}

// sub-trees (first leaf, element count) are summed by one thread each, then thread 0 walks the top of the tree
struct NpSub {
  int first_leaf;
  long long n;
};
__device__ double np_combine_top(const double* sub_sum, int& next, long long n, int depth, int top_depth) {
  if (depth == top_depth || n <= 128) return sub_sum[next++];
  long long n2 = n / 2;
  n2 -= n2 % 8;
  const double a = np_combine_top(sub_sum, next, n2, depth + 1, top_depth);
  const double b = np_combine_top(sub_sum, next, n - n2, depth + 1, top_depth);
  return __dadd_rn(a, b);
}
// out[0] = total; finish: 0 -> out[1] = total / n (the mean); 1 -> out[1] = sqrt(total / n) (the std)
__global__ void np_combine_kernel(const double* leaf_sum, const NpSub* subs, int nsubs, long long n, int top_depth,
                                  int finish, double* out) {
  __shared__ double sub_sum[256];
  const int t = threadIdx.x;
  if (t < nsubs) {
    int next = subs[t].first_leaf;
    sub_sum[t] = np_combine(leaf_sum, next, subs[t].n);
  }
  __syncthreads();
  if (t == 0) {
    int next = 0;
    const double total = np_combine_top(sub_sum, next, n, 0, top_depth);
    out[0] = total;
    const double q = __ddiv_rn(total, (double)n);
    out[1] = finish ? __dsqrt_rn(q) : q;
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
