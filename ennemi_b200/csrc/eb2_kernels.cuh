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
  int row[kMaxDimAny];
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
// a query whose search was cut short: candidates still to be examined are the segment slots
// [c_lo + rstart, c_lo + c_len) on the right and [c_lo, c_lo + lend) on the left
struct LeftEntry {
  int slot;     // query slot in P
  int c_lo;
  int c_len;
  int rstart;   // == c_len: nothing left on the right
  int lend;     // == 0: nothing left on the left
};

struct KnnArgs {
  const double* P;       // padded dimension-major point set
  int64_t stride;        // slots per row
  RowSel rows;           // the D rows that span the search space
  const Tile* tiles;
  int k;                 // result = (k+1)-th smallest, i.e. best[k]
  int sort_row;          // row of P sorted ascending inside every segment (enables pruning), or -1
  // two-level layout ("cells", single-segment sets, D >= 2): the set is ordered by sort_row ACROSS chunks of
  // chunk_len(D) slots and by rows.row[1] INSIDE each chunk; cell_lo/cell_hi = range of sort_row per chunk
  const double* cell_lo;
  const double* cell_hi;
  double* eps;           // out, per query slot
  double* heap;          // scratch for the large-k variant: [k+1][gridDim.x * tile_rows(QPT)]
  unsigned long long* pairs;  // work counter (pairs evaluated)
  int ntiles;
  // straggler deferral (pruned mode, register top-k variants): when fewer than `defer_below` threads of a
  // CTA still need candidate chunks in a direction, their queries are appended to `left_list` (with
  // their current lists in `left_best`) and finished by knn_leftover_kernel, one warp per query.
  int defer_below;
  struct LeftEntry* left_list;
  unsigned int* left_count;
  double* left_best;     // [slot][K1T]
  // two-level layout, register top-k variants, <= 3-D: warp windows of at most this many slots are walked per lane
  // (each lane only ITS OWN window of the in-chunk coordinate, see knn_scan_lane); 0 = always the all-lanes scan
  int lane_scan;
};

// sorted ascending register list; precondition v < best[K1T-1].  No NaNs can reach here (a NaN
// distance never passes the `<` test), so plain compare+select is enough (no fmin/fmax NaN fix-ups).
template <int K1T>
__device__ __forceinline__ void topk_insert(double (&best)[K1T], double v) {
  best[K1T - 1] = v;
#pragma unroll
  for (int t = K1T - 1; t > 0; --t) {
    const bool sw = best[t] < best[t - 1];
    const double lo = sw ? best[t] : best[t - 1];
    const double hi = sw ? best[t - 1] : best[t];
    best[t - 1] = lo;
    best[t] = hi;
  }
}

template <int D>
__device__ __forceinline__ double cheb(const double (&q)[D], const double (&c)[D]) {
  double m = fabs(q[0] - c[0]);
#pragma unroll
  for (int t = 1; t < D; ++t) {
    const double v = fabs(q[t] - c[t]);
    m = (v > m) ? v : m;
  }
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

// ---- in-CTA reordering of the tile's queries -------------------------------------------------------
// Bitonic sort of the kTileQ per-query keys (ascending) in shared memory; afterwards rank r holds the
// home index sidx[r] of the query with the r-th smallest key.  Used to make warps homogeneous in
// search radius so that whole warps can skip candidate chunks (exact pruning at warp granularity).
template <int NQ>
__device__ __forceinline__ void cta_sort_keys(double* skey, int* sidx) {
  const int tid = threadIdx.x;
  // With one compare-exchange per thread (blockDim >= NQ/2), the pairs of a stage with stride j <= 32 that fall
  // into the 64-element block b are exactly those of warp b: such stages only need a warp barrier, unless the
  // following stage has a longer stride (51 of the 66 stages of a 2,048-element sort).
  const bool warp_local_ok = blockDim.x >= NQ / 2;
  for (int k = 2; k <= NQ; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = tid; p < NQ / 2; p += blockDim.x) {   // NQ/2 compare-exchanges per stage
        const int i = 2 * j * (p / j) + (p % j);
        const int l = i + j;
        const bool up = (i & k) == 0;
        const double a = skey[i], b = skey[l];
        if ((a > b) == up) {
          skey[i] = b; skey[l] = a;
          const int t = sidx[i]; sidx[i] = sidx[l]; sidx[l] = t;
        }
      }
      const int next_j = (j > 1) ? (j >> 1) : k;
      if (warp_local_ok && j <= 32 && next_j <= 32) __syncwarp(); else __syncthreads();
    }
  }
  __syncthreads();
}

// moves one per-query double from its home thread to the thread that owns the query after sorting
template <int QPT>
__device__ __forceinline__ void cta_permute(double* xch, const int (&src)[QPT], double (&v)[QPT]) {
#pragma unroll
  for (int i = 0; i < QPT; ++i) xch[threadIdx.x + i * kThreads] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < QPT; ++i) v[i] = xch[src[i]];
  __syncthreads();
}

// ---- two-level layout: order the slots of every chunk by a second coordinate -----------------------
// The point set arrives ordered by `row1` globally.  Each CTA takes one chunk of TC slots, records the
// range of row1 in it (cell_lo / cell_hi) and reorders the chunk's slots by `row2` (bitonic sort in
// shared memory); all d rows and slot_row are permuted alike.  NaN padding sorts last.
template <int TC>
__global__ void __launch_bounds__((TC / 2 < 1024 ? (TC / 2 < 256 ? 256 : TC / 2) : 1024)) cell_sort_kernel(double* P, int64_t stride, int d, int* slot_row, int64_t n,
                                                            int row1, int row2, double* cell_lo, double* cell_hi) {
  __shared__ __align__(16) double skey[TC];
  __shared__ int sidx[TC];
  __shared__ __align__(16) double sval[TC];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * TC;
  const int nvalid = (int)min((int64_t)TC, n - base);
  if (tid == 0) {
    cell_lo[blockIdx.x] = P[row1 * stride + base];
    cell_hi[blockIdx.x] = P[row1 * stride + base + nvalid - 1];
  }
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  for (int i = tid; i < TC; i += blockDim.x) {
    const double v = (i < nvalid) ? P[row2 * stride + base + i] : kInf;   // padding sorts last
    skey[i] = (v == v) ? v : kInf;
    sidx[i] = i;
  }
  __syncthreads();
  cta_sort_keys<TC>(skey, sidx);
  // slots past nvalid are padding (NaN / -1) and stay padding; the row may end before base + TC
  const int nslots = (int)min((int64_t)TC, stride - base);
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  for (int t = 0; t < d; ++t) {
    for (int i = tid; i < nvalid; i += blockDim.x) sval[i] = P[t * stride + base + i];
    __syncthreads();
    for (int i = tid; i < nslots; i += blockDim.x) P[t * stride + base + i] = (i < nvalid) ? sval[sidx[i]] : kNaN;
    __syncthreads();
  }
  int* ival = reinterpret_cast<int*>(sval);
  for (int i = tid; i < nvalid; i += blockDim.x) ival[i] = slot_row[base + i];
  __syncthreads();
  for (int i = tid; i < nslots; i += blockDim.x) slot_row[base + i] = (i < nvalid) ? ival[sidx[i]] : -1;
}

// candidates tested per branch in the all-pairs inner loops (register budget: 2*G*D for the group)
__host__ __device__ constexpr int group_len(int d) { return d <= 2 ? 4 : 2; }
// resident CTAs per SM the kernels are compiled for, from a register estimate:
// queries 2*QPT*D + lists 2*QPT*K1T + candidate group 2*G*D + ~40 bookkeeping
__host__ __device__ constexpr int knn_min_blocks(int d, int k1t, int qpt) {
  const int regs = 2 * qpt * d + 2 * qpt * k1t + 2 * group_len(d) * d + 40;
  return regs <= 80 ? 3 : (regs <= 128 ? 2 : 1);
}

// all queries of one thread against one staged candidate chunk (shared memory, broadcast LDS.128)
template <int D, int K1T, int TC, int QPT>
__device__ __forceinline__ void knn_scan_chunk(const double* sbuf, int lo, int hi, const double (&q)[QPT][D],
                                               double (&best)[QPT][(K1T > 0 ? K1T : 1)], double (&thr)[QPT],
                                               const HeapRef (&heap)[QPT]) {
  constexpr int kGroup = group_len(D);
#pragma unroll 1
  for (int jj = lo; jj < hi; jj += kGroup) {
    bool any = false;
    {
      double c[kGroup][D];
#pragma unroll
      for (int t = 0; t < D; ++t) {
#pragma unroll
        for (int u = 0; u < kGroup; u += 2) {
          const double2 v = *reinterpret_cast<const double2*>(&sbuf[t * TC + jj + u]);
          c[u][t] = v.x;
          c[u + 1][t] = v.y;
        }
      }
#pragma unroll
      for (int i = 0; i < QPT; ++i) {
#pragma unroll
        for (int u = 0; u < kGroup; ++u) any = any | inside_lt<D>(q[i], c[u], thr[i]);
      }
    }
    if (any) {   // some lane has a new neighbour: re-test pair by pair (the k-th distance moves);
                 // candidates are re-read from shared memory so the group need not stay in registers
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        double cu[D];
#pragma unroll
        for (int t = 0; t < D; ++t) cu[t] = sbuf[t * TC + jj + u];
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
          if (inside_lt<D>(q[i], cu, thr[i])) {
            const double m = cheb<D>(q[i], cu);
            if constexpr (K1T > 0) { topk_insert<K1T>(best[i], m); thr[i] = best[i][K1T - 1]; }
            else thr[i] = heap[i].replace_root(m);
          }
        }
      }
    }
  }
}

// Per-lane scan of the slots [rlo, rhi) of a staged chunk, which ascend in coordinate 1 (two-level layout).
// Each query walks only the slots that can still enter ITS list: the first slot s with fl(q1 - y_s) < thr is
// found by binary search (the rounded subtraction is monotone in y_s, so the predicate is monotone in s) and
// the walk stops at the first slot with fl(y_s - q1) >= thr (monotone again; thr only shrinks on the way).
// Slots outside that run fail the coordinate-1 test of inside_lt, so skipping them cannot change any list: the
// tests on the slots looked at are the same as in knn_scan_chunk, hence bit-exact.  Slots [skip_a, skip_b)
// were already examined (seeding) and must not enter a list twice.
template <int D, int K1T, int TC, bool SKIP>
__device__ __forceinline__ void knn_scan_lane(const double* sbuf, int rlo, int rhi, const double (&q)[D], double (&best)[K1T],
                                              double& thr, int skip_a, int skip_b, unsigned long long& npairs) {
  // rlo is a multiple of G; slots are walked in aligned groups of G (independent loads, one branch per group:
  // long windows in sparse regions are not one dependent chain per slot).  Slots of a group that lie before
  // the lane's window, or in the NaN padding past the chunk's last row, simply fail the exact test.
  constexpr int G = 4;
  const double* sy = sbuf + TC;
  const double q1 = q[1];
  int lo = rlo, hi = rhi;
  {
    const double t0 = thr;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((q1 - sy[mid]) < t0) hi = mid; else lo = mid + 1;
    }
  }
#pragma unroll 1
  for (int s = lo & ~(G - 1); s < rhi; s += G) {
    double c[G][D];
    {
      const double2 y01 = *reinterpret_cast<const double2*>(&sy[s]);
      if (!((y01.x - q1) < thr)) break;
      const double2 y23 = *reinterpret_cast<const double2*>(&sy[s + 2]);
      c[0][1] = y01.x; c[1][1] = y01.y; c[2][1] = y23.x; c[3][1] = y23.y;
    }
#pragma unroll
    for (int t = 0; t < D; ++t) {
      if (t == 1) continue;
      const double2 v01 = *reinterpret_cast<const double2*>(&sbuf[t * TC + s]);
      const double2 v23 = *reinterpret_cast<const double2*>(&sbuf[t * TC + s + 2]);
      c[0][t] = v01.x; c[1][t] = v01.y; c[2][t] = v23.x; c[3][t] = v23.y;
    }
    npairs += G;
    bool any = false;
#pragma unroll
    for (int u = 0; u < G; ++u) {
      bool h = inside_lt<D>(q, c[u], thr);
      if (SKIP) h = h && (s + u < skip_a || s + u >= skip_b);
      any = any || h;
    }
    if (any) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        if (SKIP && s + u >= skip_a && s + u < skip_b) continue;
        if (inside_lt<D>(q, c[u], thr)) {
          topk_insert<K1T>(best, cheb<D>(q, c[u]));
          thr = best[K1T - 1];
        }
      }
    }
  }
}

// every slot of [sa, sb) against one query, no window (seeding: thr starts at +inf)
template <int D, int K1T, int TC>
__device__ __forceinline__ void knn_seed_lane(const double* sbuf, int sa, int sb, const double (&q)[D], double (&best)[K1T],
                                              double& thr, unsigned long long& npairs) {
#pragma unroll 1
  for (int s = sa; s < sb; ++s) {
    double c[D];
#pragma unroll
    for (int t = 0; t < D; ++t) c[t] = sbuf[t * TC + s];
    if (inside_lt<D>(q, c, thr)) {
      topk_insert<K1T>(best, cheb<D>(q, c));
      thr = best[K1T - 1];
    }
  }
  npairs += (unsigned long long)max(0, sb - sa);
}

// K1T > 0: register-resident sorted top-K1T (k+1 <= K1T).  K1T == 0: heap in global scratch, any k.
template <int D, int K1T, int QPT>
__global__ void __launch_bounds__(kThreads, knn_min_blocks(D, K1T, QPT)) knn_kernel(const KnnArgs a) {
  constexpr int kTileQ = tile_rows(QPT);
  constexpr int TC = chunk_len(D);
  constexpr int kGroup = group_len(D);
  constexpr int NB = (K1T > 0 ? K1T : 1);
  __shared__ __align__(128) double sbuf[D * TC];
  __shared__ __align__(16) double xch[kTileQ];
  __shared__ int sidx[kTileQ];
  __shared__ __align__(8) uint64_t bar;

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

    double q[QPT][D];
    double best[QPT][NB];
    double thr[QPT];
    int qslot[QPT];          // position of the query inside the tile
    int rstart[QPT], lend[QPT];   // deferred remainder of the search (segment-relative slots)
    bool valid[QPT];
    HeapRef heap[QPT];
    const bool cells = a.cell_lo != nullptr;
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      // cells: a warp owns 32*QPT CONSECUTIVE slots (consecutive in the in-chunk coordinate)
      const int qi = cells ? QPT * tid + i : tid + i * kThreads;
      qslot[i] = qi;
      rstart[i] = tile.c_len;
      lend[i] = 0;
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
    const bool prune = a.sort_row >= 0;      // then rows.row[0] == sort_row (host guarantees it)
    unsigned long long npairs = 0;

    auto fetch_chunk = [&](int j) -> int {
      const int c_off = j * TC;
      const int len = min(TC, len_pad - c_off);
      if (tid == 0) stage_chunk<D, TC>(sbuf, &bar, a.P, a.stride, a.rows, tile.c_lo + c_off, len);
      mbar_wait(&bar, phase);
      phase ^= 1;
      return len;
    };

    if (!prune) {
      for (int j = 0; j < nchunks; ++j) {
        const int len = fetch_chunk(j);
        if (tid == 0) npairs += (unsigned long long)min(TC, tile.c_len - j * TC) * tile.q_n;
        knn_scan_chunk<D, K1T, TC, QPT>(sbuf, 0, len, q, best, thr, heap);
        __syncthreads();   // everyone is done with sbuf before the next bulk copy lands in it
      }
    } else if (cells) {
      if constexpr (D >= 2) {
        // ---- two-level search: chunks pruned by coordinate 0 (CTA-wide), candidates inside a staged chunk
        //      by coordinate 1 (per warp: its queries are consecutive in that coordinate, so a binary search
        //      in shared memory gives the slot range that can matter).  All bounds are conservative w.r.t.
        //      the rounded tests, the exact test decides: bit-exact.
        constexpr int NH = (kTileQ / TC) > 0 ? (kTileQ / TC) : 1;     // home chunks of a tile
        const int warp = tid >> 5;
        const int home_lo = (tile.q_lo - tile.c_lo) / TC;
        const int home_hi = (tile.q_lo - tile.c_lo + tile.q_n - 1) / TC;
        const int own_lo = (tile.q_lo - tile.c_lo) + 32 * QPT * warp;            // segment-relative slots of this warp
        const int own_hi = own_lo + 32 * QPT;
        double ymin = kInf, ymax = -kInf;
#pragma unroll
        for (int i = 0; i < QPT; ++i)
          if (valid[i]) { ymin = fmin(ymin, q[i][1]); ymax = fmax(ymax, q[i][1]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
          ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        const bool warp_has = ymin <= ymax;
        // slot range [lo, hi) of the staged chunk j that can hold neighbours of this warp's queries
        auto window = [&](int j, int len, int& lo, int& hi) {
          double tmax = 0.0;
#pragma unroll
          for (int i = 0; i < QPT; ++i) tmax = fmax(tmax, valid[i] ? thr[i] : 0.0);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
          const double slack = 8.881784197001252e-16;  // 2^-50
          const double lo_v = (ymin - tmax) - (fabs(ymin) + tmax) * slack;
          const double hi_v = (ymax + tmax) + (fabs(ymax) + tmax) * slack;
          const int lenv = min(TC, tile.c_len - j * TC);
          const double* sy = sbuf + TC;                                      // row 1 of the staged chunk, ascending
          lo = lower_bound_ge(sy, lenv, lo_v) & ~(kGroup - 1);
          hi = min(len, (upper_bound_gt(sy, lenv, hi_v) + kGroup - 1) & ~(kGroup - 1));
        };
        bool lane_done = false;
        if constexpr (K1T > 0 && D <= 5) {
          if (a.lane_scan > 0) {
            // ---- per-lane windows.  A warp-wide window of the staged chunk (as in the path below) that holds at
            //      most a.lane_scan slots is walked by every lane on its own (knn_scan_lane: only the slots inside
            //      the lane's own coordinate-1 window are looked at); wider windows (sparse regions, where the
            //      serial walk of one lane would hold up its warp) go through the all-lanes scan.
            static_assert(NH == 1, "per-lane scan expects one home chunk per tile");
            lane_done = true;
            constexpr int kSeed = 8;             // slots on either side of the query's own slot that seed its list
            const int lane_max = a.lane_scan;
            auto scan_range = [&](int lenv, int lo, int hi, const bool (&want)[QPT]) {
              // slots [lo, hi) of the staged chunk (multiples of kGroup; hi may reach into the NaN padding)
              if (lo >= hi) return;
              if (hi - lo > lane_max) {
                if ((tid & 31) == 0) npairs += (unsigned long long)(hi - lo) * 32 * QPT;
                knn_scan_chunk<D, K1T, TC, QPT>(sbuf, lo, hi, q, best, thr, heap);    // lanes that do not want the chunk lose nothing by looking
              } else {
#pragma unroll
                for (int i = 0; i < QPT; ++i)
                  if (want[i]) knn_scan_lane<D, K1T, TC, false>(sbuf, lo, min(hi, lenv), q[i], best[i], thr[i], -1, -1, npairs);
              }
            };
            // L1. home chunk: seed every list from the query's own neighbourhood (consecutive in coordinate 1),
            //     then the rest of the warp's neighbourhood per lane, then the rest of the warp window
            {
              const int j = home_lo;
              const int len = fetch_chunk(j);
              const int lenv = min(TC, tile.c_len - j * TC);
              int ska[QPT], skb[QPT];
#pragma unroll
              for (int i = 0; i < QPT; ++i) {
                const int so = (tile.q_lo - tile.c_lo) + qslot[i] - j * TC;     // own slot, chunk-relative
                ska[i] = skb[i] = -1;
                if (valid[i]) {
                  ska[i] = max(so - kSeed, 0);
                  skb[i] = min(so + kSeed + 1, lenv);
                  knn_seed_lane<D, K1T, TC>(sbuf, ska[i], skb[i], q[i], best[i], thr[i], npairs);
                }
              }
              if (warp_has) {
                int lo = 0, hi = len;
                if (lane_max < TC) window(j, len, lo, hi);
                // the slots some lane of the warp has seeded from: always per lane (each lane skips its own seeds)
                const int ua = max(lo, max(own_lo - j * TC - kSeed, 0) & ~(kGroup - 1));
                const int ub = min(hi, (min(own_hi - j * TC + kSeed, len) + kGroup - 1) & ~(kGroup - 1));
#pragma unroll
                for (int i = 0; i < QPT; ++i)
                  if (valid[i] && ua < ub)
                    knn_scan_lane<D, K1T, TC, true>(sbuf, ua, min(ub, lenv), q[i], best[i], thr[i], ska[i], skb[i], npairs);
                scan_range(lenv, lo, min(ua, hi), valid);
                scan_range(lenv, max(ub, lo), hi, valid);
              }
              __syncthreads();
            }
            // L2. outwards over the chunks: same chunk-level rules as the all-lanes path below
            for (int j = home_hi + 1; j < nchunks; ++j) {
              const double cmin = a.cell_lo[j];
              bool need[QPT], any_need = false;
#pragma unroll
              for (int i = 0; i < QPT; ++i) { need[i] = valid[i] && !((cmin - q[i][0]) >= thr[i]); any_need = any_need || need[i]; }
              const bool wneed = __any_sync(0xffffffffu, any_need);
              const int nneed = __syncthreads_count(any_need);
              if (nneed == 0) break;
              if (nneed < a.defer_below) {
#pragma unroll
                for (int i = 0; i < QPT; ++i)
                  if (need[i]) rstart[i] = j * TC;
                break;
              }
              const int len = fetch_chunk(j);
              if (wneed) {
                int lo = 0, hi = len;
                if (lane_max < TC) window(j, len, lo, hi);
                scan_range(min(TC, tile.c_len - j * TC), lo, hi, need);
              }
            }
            __syncthreads();
            for (int j = home_lo - 1; j >= 0; --j) {
              const double cmax = a.cell_hi[j];
              bool need[QPT], any_need = false;
#pragma unroll
              for (int i = 0; i < QPT; ++i) { need[i] = valid[i] && !((q[i][0] - cmax) >= thr[i]); any_need = any_need || need[i]; }
              const bool wneed = __any_sync(0xffffffffu, any_need);
              const int nneed = __syncthreads_count(any_need);
              if (nneed == 0) break;
              if (nneed < a.defer_below) {
#pragma unroll
                for (int i = 0; i < QPT; ++i)
                  if (need[i]) lend[i] = min((j + 1) * TC, tile.c_len);
                break;
              }
              const int len = fetch_chunk(j);
              if (wneed) {
                int lo = 0, hi = len;
                if (lane_max < TC) window(j, len, lo, hi);
                scan_range(min(TC, tile.c_len - j * TC), lo, hi, need);
              }
            }
            __syncthreads();
          }
        }
        if (!lane_done) {
        int sa[NH], sb[NH];
        // 1a. seed every list from the warp's own neighbourhood in its home chunk(s)
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int j = home_lo + h;
          sa[h] = sb[h] = 0;
          if (j <= home_hi) {
            const int len = fetch_chunk(j);
            int aa = max(own_lo - 64, j * TC) - j * TC, bb = min(own_hi + 64, j * TC + len) - j * TC;
            if (warp_has && aa < bb) {
              aa &= ~(kGroup - 1);
              bb = min(len, (bb + kGroup - 1) & ~(kGroup - 1));
              sa[h] = aa; sb[h] = bb;
              if ((tid & 31) == 0) npairs += (unsigned long long)(bb - aa) * 32 * QPT;
              knn_scan_chunk<D, K1T, TC, QPT>(sbuf, aa, bb, q, best, thr, heap);
            }
            __syncthreads();
          }
        }
        // 1b. the rest of the home chunk(s), windowed
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int j = home_lo + h;
          if (j <= home_hi) {
            const int len = fetch_chunk(j);
            if (warp_has) {
              int lo, hi;
              window(j, len, lo, hi);
              const int e1 = min(sa[h], hi), s2 = max(sb[h], lo);
              if ((tid & 31) == 0) npairs += (unsigned long long)(max(0, e1 - lo) + max(0, hi - s2)) * 32 * QPT;
              if (lo < e1) knn_scan_chunk<D, K1T, TC, QPT>(sbuf, lo, e1, q, best, thr, heap);
              if (s2 < hi) knn_scan_chunk<D, K1T, TC, QPT>(sbuf, s2, hi, q, best, thr, heap);
            }
            __syncthreads();
          }
        }
        // 2. outwards over the chunks
        for (int j = home_hi + 1; j < nchunks; ++j) {
          const double cmin = a.cell_lo[j];
          bool need = false;
#pragma unroll
          for (int i = 0; i < QPT; ++i) need = need || (valid[i] && !((cmin - q[i][0]) >= thr[i]));
          const bool wneed = __any_sync(0xffffffffu, need);
          const int nneed = __syncthreads_count(need);
          if (nneed == 0) break;
          if (nneed < a.defer_below) {
#pragma unroll
            for (int i = 0; i < QPT; ++i)
              if (valid[i] && !((cmin - q[i][0]) >= thr[i])) rstart[i] = j * TC;
            break;
          }
          const int len = fetch_chunk(j);
          if (wneed) {
            int lo, hi;
            window(j, len, lo, hi);
            if ((tid & 31) == 0) npairs += (unsigned long long)max(0, hi - lo) * 32 * QPT;
            if (lo < hi) knn_scan_chunk<D, K1T, TC, QPT>(sbuf, lo, hi, q, best, thr, heap);
          }
        }
        __syncthreads();
        for (int j = home_lo - 1; j >= 0; --j) {
          const double cmax = a.cell_hi[j];
          bool need = false;
#pragma unroll
          for (int i = 0; i < QPT; ++i) need = need || (valid[i] && !((q[i][0] - cmax) >= thr[i]));
          const bool wneed = __any_sync(0xffffffffu, need);
          const int nneed = __syncthreads_count(need);
          if (nneed == 0) break;
          if (nneed < a.defer_below) {
#pragma unroll
            for (int i = 0; i < QPT; ++i)
              if (valid[i] && !((q[i][0] - cmax) >= thr[i])) lend[i] = min((j + 1) * TC, tile.c_len);
            break;
          }
          const int len = fetch_chunk(j);
          if (wneed) {
            int lo, hi;
            window(j, len, lo, hi);
            if ((tid & 31) == 0) npairs += (unsigned long long)max(0, hi - lo) * 32 * QPT;
            if (lo < hi) knn_scan_chunk<D, K1T, TC, QPT>(sbuf, lo, hi, q, best, thr, heap);
          }
        }
        __syncthreads();
        }
      }
    } else {
      // 1. the chunks that overlap the tile itself: seeds every list with near neighbours
      const int home_lo = (tile.q_lo - tile.c_lo) / TC;
      const int home_hi = (tile.q_lo - tile.c_lo + tile.q_n - 1) / TC;
      for (int j = home_lo; j <= home_hi; ++j) {
        const int len = fetch_chunk(j);
        if (tid == 0) npairs += (unsigned long long)min(TC, tile.c_len - j * TC) * tile.q_n;
        knn_scan_chunk<D, K1T, TC, QPT>(sbuf, 0, len, q, best, thr, heap);
        __syncthreads();
      }
      if (home_lo > 0 || home_hi + 1 < nchunks) {
        // 2. regroup the tile's queries by current k-th distance: warp w gets ranks [64w, 64w+64)
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
          xch[tid + i * kThreads] = valid[i] ? thr[i] : 0.0;
          sidx[tid + i * kThreads] = tid + i * kThreads;
        }
        __syncthreads();
        cta_sort_keys<kTileQ>(xch, sidx);
        int src[QPT];
#pragma unroll
        for (int i = 0; i < QPT; ++i) src[i] = sidx[QPT * tid + i];
        __syncthreads();
#pragma unroll
        for (int t = 0; t < D; ++t) {
          double v[QPT];
#pragma unroll
          for (int i = 0; i < QPT; ++i) v[i] = q[i][t];
          cta_permute<QPT>(xch, src, v);
#pragma unroll
          for (int i = 0; i < QPT; ++i) q[i][t] = v[i];
        }
        if constexpr (K1T > 0) {
#pragma unroll
          for (int t = 0; t < K1T; ++t) {
            double v[QPT];
#pragma unroll
            for (int i = 0; i < QPT; ++i) v[i] = best[i][t];
            cta_permute<QPT>(xch, src, v);
#pragma unroll
            for (int i = 0; i < QPT; ++i) best[i][t] = v[i];
          }
#pragma unroll
          for (int i = 0; i < QPT; ++i) thr[i] = best[i][K1T - 1];
        } else {
          cta_permute<QPT>(xch, src, thr);
#pragma unroll
          for (int i = 0; i < QPT; ++i) heap[i].base = a.heap + (int64_t)blockIdx.x * kTileQ + src[i];
        }
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
          qslot[i] = src[i];
          valid[i] = src[i] < tile.q_n;
        }
        // 3. outwards; a query needs chunk j only while the gap to the chunk in the sorted coordinate is
        //    below its current k-th distance (rounded subtraction is monotone => exact); a warp scans
        //    the chunk iff one of its queries needs it; the CTA stops when no warp does.
        const double* srow = a.P + a.sort_row * a.stride + tile.c_lo;
        for (int j = home_hi + 1; j < nchunks; ++j) {
          const double cmin = srow[j * TC];
          bool need = false;
#pragma unroll
          for (int i = 0; i < QPT; ++i) need = need || (valid[i] && !((cmin - q[i][0]) >= thr[i]));
          const bool wneed = __any_sync(0xffffffffu, need);
          const int nneed = __syncthreads_count(need);
          if (nneed == 0) break;
          if (nneed < a.defer_below) {   // stragglers: hand the rest of this direction to the leftover kernel
#pragma unroll
            for (int i = 0; i < QPT; ++i)
              if (valid[i] && !((cmin - q[i][0]) >= thr[i])) rstart[i] = j * TC;
            break;
          }
          const int len = fetch_chunk(j);
          if (wneed) {
            if ((tid & 31) == 0) npairs += (unsigned long long)min(TC, tile.c_len - j * TC) * 32 * QPT;
            knn_scan_chunk<D, K1T, TC, QPT>(sbuf, 0, len, q, best, thr, heap);
          }
        }
        __syncthreads();
        for (int j = home_lo - 1; j >= 0; --j) {
          const double cmax = srow[min((j + 1) * TC, tile.c_len) - 1];
          bool need = false;
#pragma unroll
          for (int i = 0; i < QPT; ++i) need = need || (valid[i] && !((q[i][0] - cmax) >= thr[i]));
          const bool wneed = __any_sync(0xffffffffu, need);
          const int nneed = __syncthreads_count(need);
          if (nneed == 0) break;
          if (nneed < a.defer_below) {
#pragma unroll
            for (int i = 0; i < QPT; ++i)
              if (valid[i] && !((q[i][0] - cmax) >= thr[i])) lend[i] = min((j + 1) * TC, tile.c_len);
            break;
          }
          const int len = fetch_chunk(j);
          if (wneed) {
            if ((tid & 31) == 0) npairs += (unsigned long long)min(TC, tile.c_len - j * TC) * 32 * QPT;
            knn_scan_chunk<D, K1T, TC, QPT>(sbuf, 0, len, q, best, thr, heap);
          }
        }
        __syncthreads();
      }
    }

#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      if (valid[i]) {
        double r;
        if constexpr (K1T > 0) {
          r = best[i][0];
#pragma unroll
          for (int t = 1; t < K1T; ++t) r = (t == a.k) ? best[i][t] : r;
        } else {
          r = thr[i];
        }
        a.eps[tile.q_lo + qslot[i]] = r;
        if constexpr (K1T > 0) {
          if (rstart[i] < tile.c_len || lend[i] > 0) {
            const unsigned int e = atomicAdd(a.left_count, 1u);
            LeftEntry le;
            le.slot = tile.q_lo + qslot[i]; le.c_lo = tile.c_lo; le.c_len = tile.c_len;
            le.rstart = rstart[i]; le.lend = lend[i];
            a.left_list[e] = le;
#pragma unroll
            for (int t = 0; t < K1T; ++t) a.left_best[(int64_t)le.slot * K1T + t] = best[i][t];
          }
        }
      }
    }
    if (a.pairs) {
      // per-thread pair counts: thread 0 carries the CTA-wide phases, lane 0 of each warp its own chunks
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
      if ((tid & 31) == 0 && npairs) atomicAdd(a.pairs, npairs);
    }
  }
}

// first slot s in [lo, hi) of an ascending coordinate array with pred(s) true, pred monotone
// (false ... false true ... true); all 32 lanes cooperate: 32 probes per round, ~4 rounds for 10^6 slots
template <typename Pred>
__device__ __forceinline__ int warp_first_true(int lo, int hi, Pred pred) {
  const int lane = threadIdx.x & 31;
  while (hi - lo > 32) {
    const int step = (hi - lo + 32) / 33;             // probes at lo + step*(lane+1) - 1
    const int pos = min(lo + step * (lane + 1) - 1, hi - 1);
    const unsigned int m = __ballot_sync(0xffffffffu, pred(pos));
    if (m == 0) { lo = min(lo + step * 32, hi); if (lo >= hi) return hi; continue; }
    const int f = __ffs(m) - 1;                        // first probing lane whose slot satisfies pred
    hi = min(lo + step * (f + 1) - 1, hi - 1) + 1;
    lo = lo + step * f;
  }
  const int pos = lo + lane;
  const unsigned int m = __ballot_sync(0xffffffffu, pos < hi && pred(pos));
  return m ? lo + (__ffs(m) - 1) : hi;
}

// pops the k1 smallest values held across the lanes' sorted lists (K1T each); returns the r-th popped
// value in lane r (for r < k1, k1 <= 32) — i.e. the merged sorted list, one element per lane
template <int K1T>
__device__ __forceinline__ double warp_merge_lists(const double (&best)[K1T], int k1) {
  const int lane = threadIdx.x & 31;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  int head = 0;
  double out = kInf;
  for (int r = 0; r < k1; ++r) {
    double mine = kInf;
#pragma unroll
    for (int t = 0; t < K1T; ++t) mine = (t == head) ? best[t] : mine;
    double mn = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    const unsigned int who = __ballot_sync(0xffffffffu, mine == mn && head < K1T);
    if (who != 0 && lane == __ffs(who) - 1) ++head;
    if (lane == r) out = mn;
  }
  return out;
}

// One CTA per deferred query: the 256 threads share out the remaining candidates (coalesced loads
// straight from L2, four independent candidates per thread per step), each keeps a private sorted
// top-K1T gated by the query's current k-th distance; the lists are merged per warp, then across the
// warps together with the list the main kernel left behind.  Same exact tests, same order statistic.
template <int D, int K1T>
__global__ void __launch_bounds__(kThreads) knn_leftover_kernel(const KnnArgs a) {
  constexpr int U = 4;
  constexpr int NW = kThreads / 32;
  __shared__ double wlist[NW + 1][K1T];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int nent = *a.left_count;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  const int k1 = a.k + 1;
  unsigned long long npairs = 0;
  // Two-level layout: most deferred queries still need a handful of chunks only - phase 0 gives each of them to ONE
  // warp (no block barriers, eight queries in flight per CTA); the few that need more than kLight chunks, and all
  // queries of one-level layouts, get a whole CTA in phase 1.
  constexpr int kLight = 64;
  const bool cells = (D >= 2) && a.cell_lo != nullptr;
  for (int phase = cells ? 0 : 1; phase < 2; ++phase) {
  const bool per_warp = phase == 0;
  const unsigned int e_first = per_warp ? blockIdx.x * NW + warp : blockIdx.x;
  const unsigned int e_step = per_warp ? gridDim.x * NW : gridDim.x;
  for (unsigned int e = e_first; e < nent; e += e_step) {
    const LeftEntry le = a.left_list[e];
    double q[D];
#pragma unroll
    for (int t = 0; t < D; ++t) q[t] = a.P[a.rows.row[t] * a.stride + le.slot];
    double best[K1T];
#pragma unroll
    for (int t = 0; t < K1T; ++t) best[t] = kInf;
    const double gate = a.left_best[(int64_t)le.slot * K1T + (K1T - 1)];   // current k-th distance (upper bound of eps)
    double thr = gate;
    const double* srow = a.P + a.sort_row * a.stride + le.c_lo;
    const double q0 = q[0];
    // candidates that can still enter: gap in the sorted coordinate below the gate (monotone => exact)
    if (a.cell_lo != nullptr) {
      if constexpr (D >= 2) {
        // two-level layout: coordinate 0 is ordered across chunks, coordinate 1 inside each chunk.  Chunks whose
        // range of coordinate 0 can still matter are dealt to the warps; inside a chunk a warp-cooperative
        // search on the (ascending) coordinate-1 row gives the slots that can matter; the lanes scan those.
        constexpr int TCL = chunk_len(D);
        const int nch = (le.c_len + TCL - 1) / TCL;
        int r_lo = nch, r_hi = nch, l_lo = 0, l_hi = 0;
        if (le.rstart < le.c_len) {
          r_lo = le.rstart / TCL;
          r_hi = warp_first_true(r_lo, nch, [&](int j) { return (a.cell_lo[j] - q0) >= gate; });
        }
        if (le.lend > 0) {
          l_hi = (le.lend + TCL - 1) / TCL;
          l_lo = warp_first_true(0, l_hi, [&](int j) { return !((q0 - a.cell_hi[j]) >= gate); });
        }
        const double slack = 8.881784197001252e-16;  // 2^-50
        const double lo_v = (q[1] - gate) - (fabs(q[1]) + gate) * slack;
        const double hi_v = (q[1] + gate) + (fabs(q[1]) + gate) * slack;
        const int nr = r_hi - r_lo, nl = l_hi - l_lo;
        if (((nr + nl) <= kLight) != per_warp) continue;      // the other phase's entry (uniform over the CTA in phase 1)
        const int wi = per_warp ? 0 : warp, gw = per_warp ? 1 : NW;
        // batches of 32 chunks per warp: every lane binary-searches the coordinate-1 window of ITS chunk (the
        // 32 searches overlap their L2 latency), then the warp walks the non-empty windows together
        for (int b0 = wi * 32; b0 < nr + nl; b0 += gw * 32) {
          const int idx = b0 + lane;
          int wlo = 0, whi = 0, base = 0;
          if (idx < nr + nl) {
            const int c = idx < nr ? r_lo + idx : l_lo + (idx - nr);
            base = c * TCL;
            const int lenv = min(TCL, le.c_len - base);
            const double* yrow = a.P + a.rows.row[1] * a.stride + le.c_lo + base;
            wlo = lower_bound_ge(yrow, lenv, lo_v);
            whi = upper_bound_gt(yrow, lenv, hi_v);
          }
          unsigned int todo = __ballot_sync(0xffffffffu, wlo < whi);
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int slo = __shfl_sync(0xffffffffu, wlo, src), shi = __shfl_sync(0xffffffffu, whi, src);
            const int sbase = __shfl_sync(0xffffffffu, base, src);
            for (int j = slo + lane; j < shi; j += 32) {
              double c_[D];
#pragma unroll
              for (int t = 0; t < D; ++t) c_[t] = a.P[a.rows.row[t] * a.stride + le.c_lo + sbase + j];
              const double m = cheb<D>(q, c_);
              if (m < thr) { topk_insert<K1T>(best, m); thr = fmin(best[K1T - 1], gate); }
            }
            if (lane == 0) npairs += (unsigned long long)(shi - slo);
          }
        }
      }
    } else {
      int lo = le.rstart, hi = le.rstart;
      int lo2 = le.lend, hi2 = le.lend;
      if (le.rstart < le.c_len) hi = warp_first_true(le.rstart, le.c_len, [&](int s) { return (srow[s] - q0) >= gate; });
      if (le.lend > 0) lo2 = warp_first_true(0, le.lend, [&](int s) { return !((q0 - srow[s]) >= gate); });
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int b = pass == 0 ? lo : lo2, en = pass == 0 ? hi : hi2;
        for (int base = b; base < en; base += kThreads * U) {
          double c[U][D];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = min(base + u * kThreads + tid, en - 1);
#pragma unroll
            for (int t = 0; t < D; ++t) c[u][t] = a.P[a.rows.row[t] * a.stride + le.c_lo + j];
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (base + u * kThreads + tid < en) {
              const double m = cheb<D>(q, c[u]);
              if (m < thr) { topk_insert<K1T>(best, m); thr = fmin(best[K1T - 1], gate); }
            }
          }
        }
        if (tid == 0) npairs += (unsigned long long)(en - b);
      }
    }
    if (per_warp) {
      // the list the main kernel left behind joins lane 0's (disjoint candidates), then one merge across the lanes
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < K1T; ++t) {
          const double v = a.left_best[(int64_t)le.slot * K1T + t];
          if (v < best[K1T - 1]) topk_insert<K1T>(best, v);
        }
      }
      const double fin = warp_merge_lists<K1T>(best, k1);
      if (lane == a.k) a.eps[le.slot] = fin;
      continue;
    }
    // per-warp merge -> shared; then warp 0 merges the NW warp lists and the stored list
    const double mine = warp_merge_lists<K1T>(best, K1T);
    if (lane < K1T) wlist[warp][lane] = mine;
    if (warp == 0 && lane < K1T) wlist[NW][lane] = a.left_best[(int64_t)le.slot * K1T + lane];
    __syncthreads();
    if (warp == 0) {
      double l2[K1T];
#pragma unroll
      for (int t = 0; t < K1T; ++t) l2[t] = (lane <= NW) ? wlist[lane][t] : kInf;
      const double fin = warp_merge_lists<K1T>(l2, k1);
      if (lane == a.k) a.eps[le.slot] = fin;
    }
    __syncthreads();
  }
  }
  if (a.pairs && (tid & 31) == 0 && npairs) atomicAdd(a.pairs, npairs);
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
  // two-level candidate layout (single segment, C >= 2): shared coordinate 0 ordered across chunks of
  // chunk_len(C+E) slots, shared coordinate 1 inside each chunk; cell_lo/hi = range of coordinate 0 per chunk
  const double* cell_lo;
  const double* cell_hi;
  int split;             // >= 1: the chunk range of every tile is dealt to this many CTAs, which ADD their counts
                         // (outputs zeroed by the caller when split > 1): more, shorter work items for small sets
  int* cnt_s;            // out per query slot (C > 0)
  int* cnt_e0;           // out (E > 0)
  int* cnt_e1;           // out (E > 1)
  unsigned long long* pairs;
};

template <int C, int E, int QPT>
__global__ void __launch_bounds__(kThreads, knn_min_blocks(C + E, 2, QPT)) count_kernel(const CountArgs a) {
  constexpr int kTileQ = tile_rows(QPT);
  constexpr int D = C + E;
  constexpr int TC = chunk_len(D);
  constexpr int kGroup = group_len(D);
  constexpr int CS = (C > 0 ? C : 1), ES = (E > 0 ? E : 1);
  __shared__ __align__(128) double sbuf[D * TC];
  __shared__ __align__(16) double xch[kTileQ];
  __shared__ int sidx[kTileQ];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int wrange[2 * (kThreads / 32)];

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  RowSel brows;
#pragma unroll
  for (int t = 0; t < C; ++t) brows.row[t] = a.b_srow.row[t];
#pragma unroll
  for (int t = 0; t < E; ++t) brows.row[C + t] = a.b_erow.row[t];

  for (int work = blockIdx.x; work < a.ntiles * a.split; work += gridDim.x) {
    const int tile_id = work / a.split, part = work - tile_id * a.split;
    const Tile tile = a.tiles[tile_id];
    double qs[QPT][CS];
    double qe[QPT][ES];
    double r[QPT];
    int ns[QPT], ne0[QPT], ne1[QPT];
    int qslot[QPT];
    bool valid[QPT];
    const bool cells = C >= 2 && a.cell_lo != nullptr;
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      const int qi = cells ? QPT * tid + i : tid + i * kThreads;   // cells: a warp owns consecutive slots
      qslot[i] = qi;
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
    int ch_lo = 0, ch_hi = (len_pad + TC - 1) / TC;   // chunk range [ch_lo, ch_hi) of the CTA
    int w_lo = ch_lo, w_hi = ch_hi;                   // ... and of this warp
    double y2lo = -kInf, y2hi = kInf;                 // cells: this warp's window in shared coordinate 1
    if (cells) {
      if constexpr (C >= 2) {
        double vmin = kInf, vmax = -kInf, umin = kInf, umax = -kInf, rmax = -kInf;
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
          if (valid[i]) {
            vmin = fmin(vmin, qs[i][0]); vmax = fmax(vmax, qs[i][0]);
            umin = fmin(umin, qs[i][1]); umax = fmax(umax, qs[i][1]);
            rmax = fmax(rmax, r[i]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
          vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
          umin = fmin(umin, __shfl_xor_sync(0xffffffffu, umin, o));
          umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
          rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        }
        const double slack = 8.881784197001252e-16;  // 2^-50
        int c_lo = 0, c_hi = 0;
        if (rmax >= 0.0) {   // otherwise nothing can be counted for this warp
          const int nch = ch_hi;
          const double lo_v = (vmin - rmax) - (fabs(vmin) + fabs(rmax)) * slack;
          const double hi_v = (vmax + rmax) + (fabs(vmax) + fabs(rmax)) * slack;
          if ((tid & 31) == 0) {
            c_lo = lower_bound_ge(a.cell_hi, nch, lo_v);    // first chunk whose coordinate-0 range reaches lo_v
            c_hi = upper_bound_gt(a.cell_lo, nch, hi_v);    // first chunk that starts beyond hi_v
          }
          c_lo = __shfl_sync(0xffffffffu, c_lo, 0);
          c_hi = __shfl_sync(0xffffffffu, c_hi, 0);
          y2lo = (umin - rmax) - (fabs(umin) + fabs(rmax)) * slack;
          y2hi = (umax + rmax) + (fabs(umax) + fabs(rmax)) * slack;
        }
        w_lo = c_lo; w_hi = max(c_lo, c_hi);
        if ((tid & 31) == 0) {
          wrange[2 * (tid >> 5)] = w_hi > w_lo ? w_lo : 0x7fffffff;
          wrange[2 * (tid >> 5) + 1] = w_hi > w_lo ? w_hi : 0;
        }
        __syncthreads();
        ch_lo = 0x7fffffff; ch_hi = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
          ch_lo = min(ch_lo, wrange[2 * w]);
          ch_hi = max(ch_hi, wrange[2 * w + 1]);
        }
        __syncthreads();
      }
    } else if (a.prune_b_row >= 0) {
      // the pruning coordinate is the first shared one (or the only extra one): host guarantees it
      // 1. regroup the tile's queries by radius so that warps are homogeneous
#pragma unroll
      for (int i = 0; i < QPT; ++i) {
        xch[tid + i * kThreads] = valid[i] ? r[i] : -kInf;
        sidx[tid + i * kThreads] = tid + i * kThreads;
      }
      __syncthreads();
      cta_sort_keys<kTileQ>(xch, sidx);
      int src[QPT];
#pragma unroll
      for (int i = 0; i < QPT; ++i) src[i] = sidx[QPT * tid + i];
      __syncthreads();
#pragma unroll
      for (int t = 0; t < C; ++t) {
        double v[QPT];
#pragma unroll
        for (int i = 0; i < QPT; ++i) v[i] = qs[i][t];
        cta_permute<QPT>(xch, src, v);
#pragma unroll
        for (int i = 0; i < QPT; ++i) qs[i][t] = v[i];
      }
#pragma unroll
      for (int t = 0; t < E; ++t) {
        double v[QPT];
#pragma unroll
        for (int i = 0; i < QPT; ++i) v[i] = qe[i][t];
        cta_permute<QPT>(xch, src, v);
#pragma unroll
        for (int i = 0; i < QPT; ++i) qe[i][t] = v[i];
      }
      cta_permute<QPT>(xch, src, r);
#pragma unroll
      for (int i = 0; i < QPT; ++i) {
        qslot[i] = src[i];
        valid[i] = src[i] < tile.q_n;
      }
      // 2. per warp: candidates outside [min(q) - max(r), max(q) + max(r)] in the sorted coordinate cannot
      //    be neighbours of any of its queries; the bounds are widened by 2^-50 relative so that the
      //    rounded subtraction in the exact test can never disagree with them.
      double vmin = kInf, vmax = -kInf, rmax = -kInf;
#pragma unroll
      for (int i = 0; i < QPT; ++i) {
        if (valid[i]) {
          const double v = (C > 0) ? qs[i][0] : qe[i][0];
          vmin = fmin(vmin, v);
          vmax = fmax(vmax, v);
          rmax = fmax(rmax, r[i]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      }
      int s_lo = 0, s_hi = 0;
      if ((tid & 31) == 0 && rmax >= 0.0) {    // rmax < 0: every radius negative (or no query), nothing to count
        const double slack = 8.881784197001252e-16;  // 2^-50
        const double lo_v = (vmin - rmax) - (fabs(vmin) + fabs(rmax)) * slack;
        const double hi_v = (vmax + rmax) + (fabs(vmax) + fabs(rmax)) * slack;
        const double* col = a.B + a.prune_b_row * a.bstride + tile.c_lo;
        s_lo = lower_bound_ge(col, tile.c_len, lo_v);
        s_hi = upper_bound_gt(col, tile.c_len, hi_v);
      }
      s_lo = __shfl_sync(0xffffffffu, s_lo, 0);
      s_hi = __shfl_sync(0xffffffffu, s_hi, 0);
      w_lo = s_lo / TC;
      w_hi = s_hi > s_lo ? (s_hi + TC - 1) / TC : w_lo;
      if ((tid & 31) == 0) {
        wrange[2 * (tid >> 5)] = w_hi > w_lo ? w_lo : 0x7fffffff;
        wrange[2 * (tid >> 5) + 1] = w_hi > w_lo ? w_hi : 0;
      }
      __syncthreads();
      ch_lo = 0x7fffffff; ch_hi = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        ch_lo = min(ch_lo, wrange[2 * w]);
        ch_hi = max(ch_hi, wrange[2 * w + 1]);
      }
      __syncthreads();
    }
    unsigned long long npairs = 0;
    const bool dense_hits = a.prune_b_row >= 0 || cells;
    if (a.split > 1 && ch_hi > ch_lo) {        // this CTA's share of the tile's chunk range
      const int span = ch_hi - ch_lo;
      const int lo = ch_lo + (int)((long long)span * part / a.split);
      ch_hi = ch_lo + (int)((long long)span * (part + 1) / a.split);
      ch_lo = lo;
    }

    for (int j = ch_lo; j < ch_hi; ++j) {
      const int c_off = j * TC;
      const int len = min(TC, len_pad - c_off);
      if (tid == 0) stage_chunk<D, TC>(sbuf, &bar, a.B, a.bstride, brows, tile.c_lo + c_off, len);
      mbar_wait(&bar, phase);
      phase ^= 1;
      if (j >= w_lo && j < w_hi) {
        int jlo = 0, jhi = len;
        if (cells) {     // slots of this chunk whose shared coordinate 1 (ascending inside the chunk) can matter
          const int lenv = min(TC, tile.c_len - c_off);
          const double* sy = sbuf + TC;
          jlo = lower_bound_ge(sy, lenv, y2lo) & ~(kGroup - 1);
          jhi = min(len, (upper_bound_gt(sy, lenv, y2hi) + kGroup - 1) & ~(kGroup - 1));
          if ((tid & 31) == 0) npairs += (unsigned long long)max(0, jhi - jlo) * 32 * QPT;
        } else {
          if ((tid & 31) == 0) npairs += (unsigned long long)min(TC, tile.c_len - c_off) * 32 * QPT;
        }
#pragma unroll 1
        for (int jj = jlo; jj < jhi; jj += kGroup) {
          double c[kGroup][D];
#pragma unroll
          for (int t = 0; t < D; ++t) {
#pragma unroll
            for (int u = 0; u < kGroup; u += 2) {
              const double2 v = *reinterpret_cast<const double2*>(&sbuf[t * TC + jj + u]);
              c[u][t] = v.x;
              c[u + 1][t] = v.y;
            }
          }
          if (C > 0 && dense_hits) {
            // windowed (pruned) scan: a neighbour turns up in nearly every group of some lane, so the
            // early-out branch below would be taken all the time; count everything branch-free instead
            if constexpr (C > 0) {
#pragma unroll
              for (int i = 0; i < QPT; ++i) {
#pragma unroll
                for (int u = 0; u < kGroup; ++u) {
                  bool in = fabs(qs[i][0] - c[u][0]) <= r[i];
#pragma unroll
                  for (int t = 1; t < C; ++t) in = in && (fabs(qs[i][t] - c[u][t]) <= r[i]);
                  ns[i] += (int)in;
                  if constexpr (E > 0) ne0[i] += (int)(in && fabs(qe[i][0] - c[u][C]) <= r[i]);
                  if constexpr (E > 1) ne1[i] += (int)(in && fabs(qe[i][1] - c[u][C + 1]) <= r[i]);
                }
              }
            }
          } else if constexpr (C > 0) {
            // the shared coordinates decide; the private ones are only looked at for pairs that pass
            bool h[QPT][kGroup];
            bool any = false;
#pragma unroll
            for (int i = 0; i < QPT; ++i) {
#pragma unroll
              for (int u = 0; u < kGroup; ++u) {
                bool in = fabs(qs[i][0] - c[u][0]) <= r[i];
#pragma unroll
                for (int t = 1; t < C; ++t) in = in && (fabs(qs[i][t] - c[u][t]) <= r[i]);
                h[i][u] = in;
                any = any | in;
              }
            }
            if (any) {
#pragma unroll
              for (int i = 0; i < QPT; ++i) {
#pragma unroll
                for (int u = 0; u < kGroup; ++u) {
                  ns[i] += (int)h[i][u];
                  if constexpr (E > 0) ne0[i] += (int)(h[i][u] && fabs(qe[i][0] - c[u][C]) <= r[i]);
                  if constexpr (E > 1) ne1[i] += (int)(h[i][u] && fabs(qe[i][1] - c[u][C + 1]) <= r[i]);
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < QPT; ++i) {
#pragma unroll
              for (int u = 0; u < kGroup; ++u) {
                if constexpr (E > 0) ne0[i] += (int)(fabs(qe[i][0] - c[u][0]) <= r[i]);
                if constexpr (E > 1) ne1[i] += (int)(fabs(qe[i][1] - c[u][1]) <= r[i]);
              }
            }
          }
        }
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      if (valid[i]) {
        const int slot = tile.q_lo + qslot[i];
        if (a.split > 1) {
          if constexpr (C > 0) { if (ns[i]) atomicAdd(&a.cnt_s[slot], ns[i]); }
          if constexpr (E > 0) { if (ne0[i]) atomicAdd(&a.cnt_e0[slot], ne0[i]); }
          if constexpr (E > 1) { if (ne1[i]) atomicAdd(&a.cnt_e1[slot], ne1[i]); }
        } else {
          if constexpr (C > 0) a.cnt_s[slot] = ns[i];
          if constexpr (E > 0) a.cnt_e0[slot] = ne0[i];
          if constexpr (E > 1) a.cnt_e1[slot] = ne1[i];
        }
      }
    }
    if (a.pairs) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
      if ((tid & 31) == 0 && npairs) atomicAdd(a.pairs, npairs);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Generic run-time-dimension fallback (d > kMaxDim, up to kMaxDimAny): brute force, one query per
// thread with its coordinates in local memory, candidates staged 32 at a time in shared memory,
// k-th distance through the global-memory heap.  Same exact tests; only used for spaces wider than
// the specialised kernels cover (rare: the estimators degrade long before).
// ----------------------------------------------------------------------------------------------
constexpr int kGenChunk = 32;

struct GenKnnArgs {
  const double* P;
  int64_t stride;
  RowSel rows;
  int d;
  const Tile* tiles;     // 256-row tiles
  int ntiles;
  int k;
  double* eps;
  double* heap;          // [k+1][gridDim.x * kThreads]
  unsigned long long* pairs;
};

struct GenCountArgs {
  const double* Q;
  int64_t qstride;
  const double* B;
  int64_t bstride;
  RowSel q_srow, b_srow;   // C shared rows
  RowSel q_erow, b_erow;   // E extra rows
  int C, E;
  const double* radius;
  const Tile* tiles;
  int ntiles;
  int* cnt_s;
  int* cnt_e0;
  int* cnt_e1;
  unsigned long long* pairs;
};

}  // namespace eb2
