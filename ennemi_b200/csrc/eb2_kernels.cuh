// ennemi_b200 — the hot-path kernels (sm_100a).
//
//  (1) knn_kernel      tiled all-pairs Chebyshev (k+1)-th-neighbour distance, replaces
//                      cKDTree.query(pts, k=[k+1], p=inf)      (_entropy_estimators.py:39,108,142,194,240)
//  (2a) search_kernel  1-D marginal counts by binary search in a sorted coordinate array,
//  (2b) count_kernel   tiled all-pairs counts for >=2-D marginals (fused over marginals that share
//                      coordinates), both replace cKDTree.query_ball_point(..., return_length=True)
//                                                               (:109-110,152-154,196,243-245)
//  (3) psi_kernel      per-row digamma terms + deterministic tree reduction, replaces
//                      np.mean(_psi(..) ...)                    (:42,113,156,200,247,327-350)
//
// Bit-exactness rules: the distance is max_t |q_t - c_t| with ONE rounded
// fp64 subtraction per dimension; a neighbour is counted iff that distance <= radius (inclusive);
// the k-th distance is an order statistic of those values.  Nothing is scaled, squared or fused.
// Padding slots hold NaN: every comparison with NaN is false, so they are never neighbours.
#pragma once
#include "eb2_common.cuh"

namespace eb2 {

// ----------------------------------------------------------------------------------------------
// candidate chunk staging: D coordinate rows of one chunk -> shared memory by TMA bulk copies
// ----------------------------------------------------------------------------------------------
struct RowSel {
  int row[kMaxDim];
};

template <int D, int TC>
__device__ __forceinline__ void stage_chunk(double* sbuf, uint64_t* bar, const double* __restrict__ P, int64_t stride,
                                            const RowSel& rows, int slot_lo, int len16) {
  // one elected thread arms the barrier with the byte count and issues D bulk copies
  const uint32_t bytes = static_cast<uint32_t>(len16) * 8u;
  mbar_expect_tx(bar, bytes * D);
#pragma unroll
  for (int t = 0; t < D; ++t) bulk_g2s(sbuf + t * TC, P + rows.row[t] * stride + slot_lo, bytes, bar);
}

// ----------------------------------------------------------------------------------------------
// (1) k-th neighbour distance
// ----------------------------------------------------------------------------------------------
struct KnnArgs {
  const double* P;       // padded dimension-major point set
  int64_t stride;        // slots per row
  RowSel rows;           // the D rows that span the search space
  const Tile* tiles;
  int k;                 // result = (k+1)-th smallest, i.e. best[k]
  int sort_row;          // row of P sorted ascending inside every segment (enables pruning), or -1
  double* eps;           // out, per query slot
  double* heap;          // scratch for the large-k variant: [k+1][gridDim.x * kTileQ]
  unsigned long long* pairs;  // work counter (pairs evaluated)
  int ntiles;
};

// sorted ascending register list; precondition v < best[K1T-1]
template <int K1T>
__device__ __forceinline__ void topk_insert(double (&best)[K1T], double v) {
  best[K1T - 1] = v;
#pragma unroll
  for (int t = K1T - 1; t > 0; --t) {
    const double lo = fmin(best[t - 1], best[t]);
    const double hi = fmax(best[t - 1], best[t]);
    best[t - 1] = lo;
    best[t] = hi;
  }
}

template <int D>
__device__ __forceinline__ double cheb(const double (&q)[D], const double (&c)[D]) {
  double m = fabs(q[0] - c[0]);
#pragma unroll
  for (int t = 1; t < D; ++t) m = fmax(m, fabs(q[t] - c[t]));
  return m;
}

// strictly inside the current k-th distance in every dimension <=> Chebyshev distance < thr
template <int D>
__device__ __forceinline__ bool inside_lt(const double (&q)[D], const double (&c)[D], double thr) {
  bool h = fabs(q[0] - c[0]) < thr;
#pragma unroll
  for (int t = 1; t < D; ++t) h = h && (fabs(q[t] - c[t]) < thr);
  return h;
}

// max-heap of K1 doubles per query in global scratch (element e of query g at heap[e * nq + g])
struct HeapRef {
  double* base;
  int64_t nq;
  int k1;
  __device__ __forceinline__ double& at(int e) const { return base[(int64_t)e * nq]; }
  __device__ __forceinline__ void fill_inf() const {
    for (int e = 0; e < k1; ++e) at(e) = __longlong_as_double(0x7ff0000000000000LL);
  }
  // replace the root (current maximum) by v < root and restore the heap; returns the new root
  __device__ __forceinline__ double replace_root(double v) const {
    int i = 0;
    for (;;) {
      const int l = 2 * i + 1;
      if (l >= k1) break;
      int big = l;
      double vb = at(l);
      if (l + 1 < k1) {
        const double vr = at(l + 1);
        if (vr > vb) { vb = vr; big = l + 1; }
      }
      if (vb <= v) break;
      at(i) = vb;
      i = big;
    }
    at(i) = v;
    return at(0);
  }
};

// K1T > 0: register-resident sorted top-K1T (k+1 <= K1T).  K1T == 0: heap in global scratch, any k.
template <int D, int K1T>
__global__ void __launch_bounds__(kThreads, (D <= 4 ? 3 : 2)) knn_kernel(const KnnArgs a) {
  constexpr int TC = chunk_len(D);
  constexpr int NB = (K1T > 0 ? K1T : 1);
  __shared__ __align__(128) double sbuf[D * TC];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double red[kThreads / 32 + 1];

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);

  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];

    double q[kQpt][D];
    double best[kQpt][NB];
    double thr[kQpt];
    bool valid[kQpt];
    HeapRef heap[kQpt];
#pragma unroll
    for (int i = 0; i < kQpt; ++i) {
      const int qi = tid + i * kThreads;
      valid[i] = qi < tile.q_n;
#pragma unroll
      for (int t = 0; t < D; ++t)
        q[i][t] = valid[i] ? a.P[a.rows.row[t] * a.stride + tile.q_lo + qi] : kNaN;
#pragma unroll
      for (int t = 0; t < NB; ++t) best[i][t] = kInf;
      thr[i] = kInf;
      if constexpr (K1T == 0) {
        heap[i].nq = (int64_t)gridDim.x * kTileQ;
        heap[i].base = a.heap + (int64_t)blockIdx.x * kTileQ + qi;
        heap[i].k1 = a.k + 1;
        heap[i].fill_inf();
      }
    }

    const int len_pad = (tile.c_len + kSegAlign - 1) / kSegAlign * kSegAlign;
    const int nchunks = (len_pad + TC - 1) / TC;
    const bool prune = a.sort_row >= 0;
    int home_lo = 0, home_hi = nchunks - 1;
    double q0min = 0.0, q0max = 0.0;
    if (prune) {
      home_lo = (tile.q_lo - tile.c_lo) / TC;
      home_hi = (tile.q_lo - tile.c_lo + tile.q_n - 1) / TC;
      q0min = a.P[a.sort_row * a.stride + tile.q_lo];
      q0max = a.P[a.sort_row * a.stride + tile.q_lo + tile.q_n - 1];
    }
    unsigned long long npairs = 0;

    // visit order: home chunks, then rightwards, then leftwards; a direction stops as soon as the
    // gap in the sorted coordinate is >= every query's current k-th distance (exact: rounding is monotone)
    int j = home_lo;
    int dir = +1;
    while (true) {
      const int c_off = j * TC;
      const int len = min(TC, len_pad - c_off);
      if (tid == 0) stage_chunk<D, TC>(sbuf, &bar, a.P, a.stride, a.rows, tile.c_lo + c_off, len);
      mbar_wait(&bar, phase);
      phase ^= 1;
      npairs += (unsigned long long)min(TC, tile.c_len - c_off);

#pragma unroll 2
      for (int jj = 0; jj < len; jj += 2) {
        double c0[D], c1[D];
#pragma unroll
        for (int t = 0; t < D; ++t) {
          const double2 v = *reinterpret_cast<const double2*>(&sbuf[t * TC + jj]);
          c0[t] = v.x;
          c1[t] = v.y;
        }
#pragma unroll
        for (int i = 0; i < kQpt; ++i) {
          const bool h0 = inside_lt<D>(q[i], c0, thr[i]);
          const bool h1 = inside_lt<D>(q[i], c1, thr[i]);
          if (h0 | h1) {
            if (h0) {
              const double m = cheb<D>(q[i], c0);
              if constexpr (K1T > 0) { topk_insert<K1T>(best[i], m); thr[i] = best[i][K1T - 1]; }
              else thr[i] = heap[i].replace_root(m);
            }
            if (h1) {
              const double m = cheb<D>(q[i], c1);
              if (m < thr[i]) {
                if constexpr (K1T > 0) { topk_insert<K1T>(best[i], m); thr[i] = best[i][K1T - 1]; }
                else thr[i] = heap[i].replace_root(m);
              }
            }
          }
        }
      }
      __syncthreads();  // everyone is done with sbuf before the next bulk copy lands in it

      // pick the next chunk (uniform across the CTA)
      if (dir > 0) {
        bool go = j + 1 < nchunks;
        if (go && prune && j + 1 > home_hi) {
          double tmax = 0.0;
#pragma unroll
          for (int i = 0; i < kQpt; ++i) tmax = fmax(tmax, valid[i] ? thr[i] : 0.0);
          tmax = block_max_bcast<kThreads>(tmax, red);
          const double cmin = a.P[a.sort_row * a.stride + tile.c_lo + (j + 1) * TC];
          go = !((cmin - q0max) >= tmax);
        }
        if (go) { ++j; continue; }
        dir = -1;
        j = home_lo;
      }
      {
        bool go = j - 1 >= 0;
        if (go && prune) {
          double tmax = 0.0;
#pragma unroll
          for (int i = 0; i < kQpt; ++i) tmax = fmax(tmax, valid[i] ? thr[i] : 0.0);
          tmax = block_max_bcast<kThreads>(tmax, red);
          const double cmax = a.P[a.sort_row * a.stride + tile.c_lo + min(j * TC, tile.c_len) - 1];
          go = !((q0min - cmax) >= tmax);
        }
        if (!go) break;
        --j;
      }
    }

#pragma unroll
    for (int i = 0; i < kQpt; ++i) {
      if (valid[i]) {
        double r;
        if constexpr (K1T > 0) {
          r = best[i][0];
#pragma unroll
          for (int t = 1; t < K1T; ++t) r = (t == a.k) ? best[i][t] : r;
        } else {
          r = thr[i];
        }
        a.eps[tile.q_lo + tid + i * kThreads] = r;
      }
    }
    if (tid == 0 && a.pairs) atomicAdd(a.pairs, npairs * (unsigned long long)tile.q_n);
  }
}

// ----------------------------------------------------------------------------------------------
// (2b) all-pairs neighbour counts for marginal spaces that share C coordinates (the condition)
//      and add one private coordinate each (E extras): outputs n_shared, n_shared+e0, n_shared+e1.
//      C = 0: the extras are independent 1-D marginals (KSG n_x, n_y by brute force).
// ----------------------------------------------------------------------------------------------
struct CountArgs {
  const double* Q;       // query point set (rows: see q_* below)
  int64_t qstride;
  const double* B;       // candidate point set
  int64_t bstride;
  RowSel q_srow, b_srow; // C shared rows
  RowSel q_erow, b_erow; // E extra rows
  const double* radius;  // per query slot: count distance <= radius
  const Tile* tiles;     // query tile + candidate segment (slots of B)
  int ntiles;
  int prune_q_row;       // row of Q / row of B holding the coordinate B is sorted by inside the segment, or -1
  int prune_b_row;
  int* cnt_s;            // out per query slot (C > 0)
  int* cnt_e0;           // out (E > 0)
  int* cnt_e1;           // out (E > 1)
  unsigned long long* pairs;
};

// first slot in [0, len) of the ascending array `a` whose value is >= v (NaN padding excluded by len)
__device__ __forceinline__ int lower_bound_ge(const double* a, int len, double v) {
  int lo = 0, hi = len;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first slot whose value is > v
__device__ __forceinline__ int upper_bound_gt(const double* a, int len, double v) {
  int lo = 0, hi = len;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int C, int E>
__global__ void __launch_bounds__(kThreads, (C + E <= 4 ? 3 : 2)) count_kernel(const CountArgs a) {
  constexpr int D = C + E;
  constexpr int TC = chunk_len(D);
  __shared__ __align__(128) double sbuf[D * TC];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double red[kThreads / 32 + 1];
  __shared__ int range[2];

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  RowSel brows;
#pragma unroll
  for (int t = 0; t < C; ++t) brows.row[t] = a.b_srow.row[t];
#pragma unroll
  for (int t = 0; t < E; ++t) brows.row[C + t] = a.b_erow.row[t];

  for (int tile_id = blockIdx.x; tile_id < a.ntiles; tile_id += gridDim.x) {
    const Tile tile = a.tiles[tile_id];
    double qs[kQpt][C > 0 ? C : 1];
    double qe[kQpt][E > 0 ? E : 1];
    double r[kQpt];
    int ns[kQpt], ne0[kQpt], ne1[kQpt];
    bool valid[kQpt];
#pragma unroll
    for (int i = 0; i < kQpt; ++i) {
      const int qi = tid + i * kThreads;
      valid[i] = qi < tile.q_n;
      const int slot = tile.q_lo + qi;
#pragma unroll
      for (int t = 0; t < C; ++t) qs[i][t] = valid[i] ? a.Q[a.q_srow.row[t] * a.qstride + slot] : kNaN;
#pragma unroll
      for (int t = 0; t < E; ++t) qe[i][t] = valid[i] ? a.Q[a.q_erow.row[t] * a.qstride + slot] : kNaN;
      r[i] = valid[i] ? a.radius[slot] : kNaN;
      ns[i] = ne0[i] = ne1[i] = 0;
    }

    const int len_pad = (tile.c_len + kSegAlign - 1) / kSegAlign * kSegAlign;
    int ch_lo = 0, ch_hi = (len_pad + TC - 1) / TC;   // chunk range [ch_lo, ch_hi)
    if (a.prune_b_row >= 0) {
      // candidates outside [min(q) - max(r), max(q) + max(r)] in the sorted coordinate cannot be
      // neighbours of any query of this tile; the bounds are widened by 2^-50 relative so that
      // the rounded subtraction in the exact test can never disagree with them.
      double vmin = __longlong_as_double(0x7ff0000000000000LL), vmax = -vmin, rmax = -vmin;
#pragma unroll
      for (int i = 0; i < kQpt; ++i) {
        if (valid[i]) {
          const double v = a.Q[a.prune_q_row * a.qstride + tile.q_lo + tid + i * kThreads];
          vmin = fmin(vmin, v);
          vmax = fmax(vmax, v);
          rmax = fmax(rmax, r[i]);
        }
      }
      vmin = block_min_bcast<kThreads>(vmin, red);
      vmax = block_max_bcast<kThreads>(vmax, red);
      rmax = block_max_bcast<kThreads>(rmax, red);
      if (tid == 0) {
        const double slack = 8.881784197001252e-16;  // 2^-50
        const double lo_v = (vmin - rmax) - (fabs(vmin) + fabs(rmax)) * slack;
        const double hi_v = (vmax + rmax) + (fabs(vmax) + fabs(rmax)) * slack;
        const double* col = a.B + a.prune_b_row * a.bstride + tile.c_lo;
        int s_lo = 0, s_hi = 0;
        if (rmax >= 0.0) {   // rmax < 0: every radius negative, nothing can be counted
          s_lo = lower_bound_ge(col, tile.c_len, lo_v);
          s_hi = upper_bound_gt(col, tile.c_len, hi_v);
        }
        range[0] = s_lo / TC;
        range[1] = s_hi > s_lo ? (s_hi + TC - 1) / TC : s_lo / TC;
      }
      __syncthreads();
      ch_lo = range[0];
      ch_hi = range[1];
      __syncthreads();
    }
    unsigned long long npairs = 0;

    for (int j = ch_lo; j < ch_hi; ++j) {
      const int c_off = j * TC;
      const int len = min(TC, len_pad - c_off);
      if (tid == 0) stage_chunk<D, TC>(sbuf, &bar, a.B, a.bstride, brows, tile.c_lo + c_off, len);
      mbar_wait(&bar, phase);
      phase ^= 1;
      npairs += (unsigned long long)min(TC, tile.c_len - c_off);

#pragma unroll 2
      for (int jj = 0; jj < len; jj += 2) {
        double c0[D], c1[D];
#pragma unroll
        for (int t = 0; t < D; ++t) {
          const double2 v = *reinterpret_cast<const double2*>(&sbuf[t * TC + jj]);
          c0[t] = v.x;
          c1[t] = v.y;
        }
#pragma unroll
        for (int i = 0; i < kQpt; ++i) {
          if constexpr (C > 0) {
            bool h0 = fabs(qs[i][0] - c0[0]) <= r[i];
            bool h1 = fabs(qs[i][0] - c1[0]) <= r[i];
#pragma unroll
            for (int t = 1; t < C; ++t) {
              h0 = h0 && (fabs(qs[i][t] - c0[t]) <= r[i]);
              h1 = h1 && (fabs(qs[i][t] - c1[t]) <= r[i]);
            }
            if (h0 | h1) {
              ns[i] += (int)h0 + (int)h1;
              if constexpr (E > 0) ne0[i] += (int)(h0 && fabs(qe[i][0] - c0[C]) <= r[i]) + (int)(h1 && fabs(qe[i][0] - c1[C]) <= r[i]);
              if constexpr (E > 1) ne1[i] += (int)(h0 && fabs(qe[i][1] - c0[C + 1]) <= r[i]) + (int)(h1 && fabs(qe[i][1] - c1[C + 1]) <= r[i]);
            }
          } else {
            if constexpr (E > 0) ne0[i] += (int)(fabs(qe[i][0] - c0[0]) <= r[i]) + (int)(fabs(qe[i][0] - c1[0]) <= r[i]);
            if constexpr (E > 1) ne1[i] += (int)(fabs(qe[i][1] - c0[1]) <= r[i]) + (int)(fabs(qe[i][1] - c1[1]) <= r[i]);
          }
        }
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < kQpt; ++i) {
      if (valid[i]) {
        const int slot = tile.q_lo + tid + i * kThreads;
        if constexpr (C > 0) a.cnt_s[slot] = ns[i];
        if constexpr (E > 0) a.cnt_e0[slot] = ne0[i];
        if constexpr (E > 1) a.cnt_e1[slot] = ne1[i];
      }
    }
    if (tid == 0 && a.pairs) atomicAdd(a.pairs, npairs * (unsigned long long)tile.q_n);
  }
}

}  // namespace eb2
