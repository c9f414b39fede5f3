// ennemi_b200 — C ABI, per-device contexts and the orchestration of the hot path (sm_100a).
//
// Data layout in HBM (per call, on the device's stream, from the stream-ordered pool):
//   raw     d x n           dimension-major copy of the caller's coordinates
//   P       d x stride      "point set": the same rows permuted into processing order and padded
//                           with NaN; rows of one class form a segment that starts on a 16-slot
//                           boundary (TMA bulk copies need 16 B alignment); inside a segment the
//                           rows are ascending in one chosen coordinate (`sort_row`), which is what
//                           lets the all-pairs kernels skip candidate chunks exactly.  Single-segment
//                           sets of 2+ dimensions use the two-level layout: ascending in `sort_row`
//                           ACROSS chunks of chunk_len(d) slots, ascending in a second coordinate
//                           INSIDE each chunk (cell_sort_kernel), with the per-chunk range of
//                           `sort_row` in cell_lo / cell_hi
//   slot_row  stride        slot -> caller's row (-1 for padding)
//   eps, radius, counts     per slot
//   tiles                   one entry per 256- or 512-row query tile: query slots + candidate segment
//   cache / derived         per device, across calls: uploaded columns, noise vectors and prepared
//                           (rescaled + sorted) variables shared by the tasks of a call
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ennemi_b200.h"
#include "eb2_aux_kernels.cuh"
#include "eb2_launch.h"
#include "eb2_ksg2.h"

namespace {

using namespace eb2;

thread_local std::string g_err;
thread_local int g_data_flags = 0;   // bit0: NaN in the input columns, bit1: non-finite prepared data

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

struct CudaFail {
  cudaError_t e;
  const char* what;
  int line;
};
#define CU(x)                                            \
  do {                                                   \
    cudaError_t e_ = (x);                                \
    if (e_ != cudaSuccess) throw CudaFail{e_, #x, __LINE__}; \
  } while (0)

constexpr int EB2_ERR_RETRY_GENERAL = 1000;   // internal code, never returned through the C ABI
constexpr int kMaxDev = 64;
constexpr int kMaxLanes = 4;       // independent streams ("lanes") per device: `dev` = ordinal | lane << 8
constexpr int kNumEvents = 8;

// device-resident columns (raw observations, noise vectors) uploaded once and referenced by key;
// shared by the lanes of a device
// a prepared variable (cached column window after rescaling + noise) and its ascending order; shared by
// every task of a call that uses the same descriptor (pairwise_mi: each column 63 times per role)
struct Derived {
  double* vals = nullptr;     // n prepared values, caller's row order
  double* sorted = nullptr;   // ascending
  int* perm = nullptr;        // rank -> row
  int* dflag = nullptr;       // bit0: NaN among the inputs, bit1: non-finite prepared value
  cudaEvent_t ready = nullptr;
  uint64_t src_key = 0, noise_key = 0;
  int64_t n = 0;
};
struct DevShared {
  std::mutex mu;
  std::unordered_map<uint64_t, std::pair<double*, int64_t>> cache;
  std::unordered_map<std::string, Derived> derived;
};
DevShared g_shared[kMaxDev];

struct Ctx {
  int dev = -1;
  std::atomic<bool> ready{false};    // published with release order once every field below is set (get_ctx)
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;        // second stream: work that can overlap the k-NN kernel (marginal sort)
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t ev[kNumEvents] = {};
  int sm_count = 148;
  char* pinned = nullptr;   // host staging: tile tables up, results down
  char* bounce = nullptr;   // page-locked bounce buffer for uploads from pageable memory (kBounceSlots x kBounceBytes)
  cudaEvent_t bounce_ev[4] = {};
  size_t pinned_cap = 0;
  double last_ms[5] = {0, 0, 0, 0, 0};
  int last_launches = 0;
  int last_pipeline = 0;              // 1: the last KSG call on this lane went through the bivariate pipeline
  DevShared* shared = nullptr;
  // per-call workspace: one block per lane, bump-allocated by Scratch and kept across calls (calls on a lane are
  // serialised and stream-ordered, so the next call may reuse it); grown after a call that did not fit
  char* arena = nullptr;
  size_t arena_cap = 0;
  // tile tables of recent single-segment layouts (same n, same options => same table: every pair of a pairwise_mi call)
  struct TileEntry { int64_t key[8]; void* dev; int count; int64_t rows; };
  std::vector<TileEntry> tile_cache;
  double* psi_tab = nullptr;          // psi_ref(n), n < kPsiTab, filled on first use on this lane's stream
  // CUDA graphs of repeated resident estimates (EB2_GRAPH=0 turns them off): the lane workspace gives a repeated same-shape call the
  // same device addresses, so its ~45 launches can be replayed as one graph launch.  Entries die with the workspace
  // or the tile tables they reference.
  struct GraphEntry {
    int64_t key[10];
    cudaGraphExec_t exec = nullptr;
    void* host_result = nullptr;      // pinned block the graph's last node copies the result into
    int64_t rows = 0;
    int launches = 0;
    bool seen = false;                // ran eagerly once without outgrowing the workspace
    bool failed = false;              // capture did not work for this signature: stay eager
  };
  std::vector<GraphEntry> graphs;
  void drop_graphs() {
    for (auto& g : graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  // NumPy pairwise-summation tables of the most recent window lengths (device copies; allocated and used on `stream`)
  struct NpTables { int64_t n; void* leaves; void* children; int nleaves, ninner; NpLevels levels; };
  std::vector<NpTables> np_tables;
};

Ctx g_ctx[kMaxDev * kMaxLanes];
std::mutex g_init_mu;

Ctx& get_ctx(int dev_lane) {
  const int dev = dev_lane & 0xff, lane = (dev_lane >> 8) & 0xff;
  if (dev_lane < 0 || dev >= kMaxDev || lane >= kMaxLanes) throw CudaFail{cudaErrorInvalidDevice, "device index", __LINE__};
  Ctx& c = g_ctx[dev * kMaxLanes + lane];
  if (c.ready.load(std::memory_order_acquire)) return c;
  std::lock_guard<std::mutex> g(g_init_mu);
  if (c.ready.load(std::memory_order_acquire)) return c;
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (dev >= count) throw CudaFail{cudaErrorInvalidDevice, "device index beyond cudaGetDeviceCount", __LINE__};
  CU(cudaSetDevice(dev));
  CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming));
  for (auto& e : c.ev) CU(cudaEventCreate(&e));
  CU(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
  cudaMemPool_t pool;
  CU(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t keep = std::numeric_limits<uint64_t>::max();   // keep freed workspace cached in the pool
  CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (lane == 0) CU(k2::init());
  if (lane == 0) {
    // reserve workspace once: the pool keeps it (release threshold above), so the first large call does not
    // pay for growing the pool allocation by allocation
    void* warm = nullptr;
    size_t want = size_t(2) << 30;
    if (const char* e = getenv("EB2_POOL_MB")) want = size_t(atoll(e)) << 20;
    if (want && cudaMallocAsync(&warm, want, c.stream) == cudaSuccess) {
      cudaFreeAsync(warm, c.stream);
      cudaStreamSynchronize(c.stream);
    } else {
      cudaGetLastError();
    }
  }
  c.pinned_cap = 4 << 20;
  CU(cudaMallocHost(reinterpret_cast<void**>(&c.pinned), c.pinned_cap));
  c.dev = dev;
  c.shared = &g_shared[dev];
  c.ready.store(true, std::memory_order_release);
  return c;
}


// Host -> device copy of a large array on the lane's stream, complete on return.  Pageable memory is what users pass
// (NumPy arrays): cudaMemcpyAsync stages it internally at ~8 GB/s.  Here the calling thread copies 1 MiB pieces into a
// page-locked ring while the DMA engine moves the previous ones, which runs at the host's memcpy speed (and two lanes
// uploading two variables side by side double it).
constexpr size_t kBounceBytes = size_t(1) << 20;
constexpr int kBounceSlots = 4;
void upload(Ctx& c, void* dst, const void* src, size_t bytes) {
  cudaPointerAttributes attr;
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
  else cudaGetLastError();
  if (!pageable || bytes < 2 * kBounceBytes || getenv("EB2_NO_BOUNCE")) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    return;
  }
  if (!c.bounce) {
    CU(cudaMallocHost(reinterpret_cast<void**>(&c.bounce), kBounceBytes * kBounceSlots));
    for (auto& e : c.bounce_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  size_t done = 0;
  for (int i = 0; done < bytes; ++i) {
    const int slot = i % kBounceSlots;
    const size_t len = std::min(kBounceBytes, bytes - done);
    if (i >= kBounceSlots) CU(cudaEventSynchronize(c.bounce_ev[slot]));      // the slot's previous piece has left
    std::memcpy(c.bounce + slot * kBounceBytes, static_cast<const char*>(src) + done, len);
    CU(cudaMemcpyAsync(static_cast<char*>(dst) + done, c.bounce + slot * kBounceBytes, len, cudaMemcpyHostToDevice, c.stream));
    CU(cudaEventRecord(c.bounce_ev[slot], c.stream));
    done += len;
  }
  CU(cudaStreamSynchronize(c.stream));
}

// Per-call scratch: stream-ordered allocations, released when the call ends.
struct Scratch {
  Ctx& c;
  std::vector<void*> ptrs;
  size_t pinned_used = 0;
  int launches = 0;
  explicit Scratch(Ctx& ctx) : c(ctx) {}
  std::vector<void*> extra_pinned;
  bool used_side = false;     // work was forked to the second stream: it must finish before buffers are released
  size_t arena_used = 0, overflow = 0;
  double* result = nullptr;   // device block of the call: [0..3] reduction output, [4] pair counter (u64), [5] data flags (int)
  Ctx::GraphEntry* capture = nullptr;    // the call is being captured into this entry (stream capture is active)
  ~Scratch() {
    if (used_side) {
      cudaEventRecord(c.join, c.side);
      cudaStreamWaitEvent(c.stream, c.join, 0);
    }
    for (void* p : ptrs) cudaFreeAsync(p, c.stream);
    if (overflow) {
      // the call did not fit into the lane's workspace: replace it by one that would have held everything
      c.drop_graphs();                 // they reference the old workspace
      if (c.arena) cudaFreeAsync(c.arena, c.stream);
      c.arena = nullptr;
      c.arena_cap = 0;
      const size_t want = (arena_used + overflow) + (arena_used + overflow) / 4 + (1 << 20);
      void* p = nullptr;
      if (cudaMallocAsync(&p, want, c.stream) == cudaSuccess) {
        c.arena = static_cast<char*>(p);
        c.arena_cap = want;
      } else {
        cudaGetLastError();
      }
    }
    if (!extra_pinned.empty()) {
      cudaStreamSynchronize(c.stream);
      for (void* p : extra_pinned) cudaFreeHost(p);
    }
  }
  template <typename T>
  T* dev(size_t count) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) / 256 * 256;
    if (arena_used + bytes <= c.arena_cap) {
      T* p = reinterpret_cast<T*>(c.arena + arena_used);
      arena_used += bytes;
      return p;
    }
    if (capture) throw CudaFail{cudaErrorMemoryAllocation, "workspace too small during graph capture", __LINE__};
    void* p = nullptr;
    CU(cudaMallocAsync(&p, bytes, c.stream));
    ptrs.push_back(p);
    overflow += bytes;
    return static_cast<T*>(p);
  }
  // pinned staging slice (valid until the call ends)
  template <typename T>
  T* host(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    if (pinned_used + bytes > c.pinned_cap) {
      void* extra = nullptr;                       // rare: a dedicated block for this call
      CU(cudaMallocHost(&extra, bytes));
      extra_pinned.push_back(extra);
      return static_cast<T*>(extra);
    }
    T* p = reinterpret_cast<T*>(c.pinned + pinned_used);
    pinned_used += bytes;
    return p;
  }
};

inline int cdiv(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }
inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- the point set -----------------------------------------------------------------------------
struct PointSet {
  int qpt = kMaxQpt;          // query rows per thread of the all-pairs kernels: tiles are tile_rows(qpt) rows
  double* P = nullptr;
  int64_t stride = 0;
  int d = 0;
  int64_t n = 0;
  int* slot_row = nullptr;
  int sort_row = -1;
  std::vector<int> seg_slot, seg_len;   // per segment
  // two-level layout (see cell_sort_kernel): chunks of chunk_len(cell_dim) slots ordered by cell_row2 inside
  int cell_dim = 0;
  int cell_row2 = -1;
  double* cell_lo = nullptr;
  double* cell_hi = nullptr;
  const double* sorted_keys = nullptr;   // sort_row values in ascending order (n of them), when sorted
};

template <typename K, typename V>
void sort_pairs(Scratch& s, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit, int end_bit) {
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s.c.stream));
  void* tmp = s.dev<char>(tmp_bytes);
  CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s.c.stream));
}

// the same sort on the context's second stream, concurrent with whatever the main stream does next; every
// buffer must have been allocated (and its inputs produced) on the main stream before the call.  The caller
// joins with side_join() before the results are used.
template <typename K, typename V>
void sort_pairs_side(Scratch& s, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit, int end_bit) {
  Ctx& c = s.c;
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, n, begin_bit, end_bit, c.side));
  void* tmp = s.dev<char>(tmp_bytes);
  s.used_side = true;
  CU(cudaEventRecord(c.fork, c.stream));
  CU(cudaStreamWaitEvent(c.side, c.fork, 0));
  CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, n, begin_bit, end_bit, c.side));
  CU(cudaEventRecord(c.join, c.side));
}
void side_join(Scratch& s) { CU(cudaStreamWaitEvent(s.c.stream, s.c.join, 0)); }

void sort_keys(Scratch& s, const double* kin, double* kout, int n) {
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, kin, kout, n, 0, 64, s.c.stream));
  void* tmp = s.dev<char>(tmp_bytes);
  CU(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, kin, kout, n, 0, 64, s.c.stream));
}

// raw: d x n on the device.  cls: class id per row on the device (or NULL), class_size: rows per class.
// cell_dim > 0 asks for the two-level layout for a search space of that dimension whose coordinate 1 is
// row cell_row2 (single-segment sets only)
struct Presorted {
  const double* sorted_keys;   // row `sort_row` in ascending order
  const int* perm;             // rank -> row
};
PointSet build_point_set_rows(Scratch& s, const double* const* row_src, int d, int64_t n, const int* cls,
                              const std::vector<int>& class_size, int sort_row, int cell_dim, int cell_row2,
                              const Presorted* pre, const Presorted* pre2 = nullptr, bool pre2_on_side = false);

PointSet build_point_set(Scratch& s, const double* raw, int d, int64_t n, const int* cls,
                         const std::vector<int>& class_size, int sort_row, int cell_dim = 0, int cell_row2 = -1) {
  const double* rows[kMaxDimAny + 1];
  for (int t = 0; t < d; ++t) rows[t] = raw + static_cast<int64_t>(t) * n;
  return build_point_set_rows(s, rows, d, n, cls, class_size, sort_row, cell_dim, cell_row2, nullptr);
}

// pre / pre2: row `sort_row` / row `cell_row2` already sorted (ascending values + rank -> row).  With both, the
// two-level layout is built by PARTITION instead of per-chunk sorts: the rows are read in the order of coordinate 1
// and stably radix-sorted by their chunk (rank in coordinate 0 / chunk length): 9 key bits at N = 10^6.
PointSet build_point_set_rows(Scratch& s, const double* const* row_src, int d, int64_t n, const int* cls,
                              const std::vector<int>& class_size, int sort_row, int cell_dim, int cell_row2,
                              const Presorted* pre, const Presorted* pre2, bool pre2_on_side) {
  cudaStream_t st = s.c.stream;
  PointSet ps;
  ps.d = d;
  ps.n = n;
  ps.sort_row = sort_row;
  // 512-row tiles only when they still give every SM several CTAs; smaller sets use 256-row tiles
  ps.qpt = (n / tile_rows(kMaxQpt) >= 4 * static_cast<int64_t>(s.c.sm_count)) ? kMaxQpt : 1;
  if (const char* e = getenv("EB2_QPT")) ps.qpt = atoi(e) == 1 ? 1 : 2;   // tuning knob
  if (d > kMaxDim) ps.qpt = 1;      // the generic wide-space kernels use 256-row tiles
  const int nseg = cls ? static_cast<int>(class_size.size()) : 1;
  std::vector<int> seg_rank(nseg);
  ps.seg_slot.resize(nseg);
  ps.seg_len.resize(nseg);
  int64_t rank = 0, slot = 0;
  for (int c = 0; c < nseg; ++c) {
    const int len = cls ? class_size[c] : static_cast<int>(n);
    seg_rank[c] = static_cast<int>(rank);
    ps.seg_slot[c] = static_cast<int>(slot);
    ps.seg_len[c] = len;
    rank += len;
    slot += round_up(len, kSegAlign);
  }
  ps.stride = std::max<int64_t>(slot, kSegAlign);
  const int blocks_n = cdiv(n, 256);

  // rank -> input row
  const int* perm = nullptr;
  const int* cls_sorted = nullptr;
  if (sort_row >= 0) {
    const int* p1 = nullptr;
    if (pre) {
      p1 = pre->perm;
      ps.sorted_keys = pre->sorted_keys;
    } else {
      int* iota = s.dev<int>(n);
      iota_kernel<<<blocks_n, 256, 0, st>>>(iota, static_cast<int>(n));
      s.launches++;
      double* keys_out = s.dev<double>(n);
      int* pp = s.dev<int>(n);
      sort_pairs<double, int>(s, row_src[sort_row], keys_out, iota, pp, static_cast<int>(n), 0, 64);
      p1 = pp;
      ps.sorted_keys = keys_out;
    }
    perm = p1;
    if (cls) {
      // stable second pass by class keeps the coordinate order inside each class
      int* cls_by_rank = s.dev<int>(n);
      gather_int_kernel<<<blocks_n, 256, 0, st>>>(cls, p1, cls_by_rank, static_cast<int>(n));
      s.launches++;
      int* cls_out = s.dev<int>(n);
      int* p2 = s.dev<int>(n);
      int bits = 1;
      while ((1 << bits) < nseg && bits < 31) ++bits;
      sort_pairs<int, int>(s, cls_by_rank, cls_out, p1, p2, static_cast<int>(n), 0, bits);
      perm = p2;
      cls_sorted = cls_out;
    }
  } else if (cls) {
    int* iota = s.dev<int>(n);
    iota_kernel<<<blocks_n, 256, 0, st>>>(iota, static_cast<int>(n));
    s.launches++;
    int* cls_out = s.dev<int>(n);
    int* p2 = s.dev<int>(n);
    int bits = 1;
    while ((1 << bits) < nseg && bits < 31) ++bits;
    sort_pairs<int, int>(s, cls, cls_out, iota, p2, static_cast<int>(n), 0, bits);
    perm = p2;
    cls_sorted = cls_out;
  }

  const bool want_cells = cell_dim >= 2 && cell_dim <= kMaxDim && cell_row2 >= 0 && sort_row >= 0 && !cls && !getenv("EB2_NO_CELLS");
  bool cells_done = false;
  if (want_cells && pre2 && perm && !getenv("EB2_CELL_SORT")) {
    const int tc = chunk_len(cell_dim);
    const int nchunks = cdiv(n, tc);
    int* row_chunk = s.dev<int>(n);
    int* keys = s.dev<int>(n);
    int* keys_out = s.dev<int>(n);
    int* order = s.dev<int>(n);
    row_chunk_kernel<<<blocks_n, 256, 0, st>>>(perm, static_cast<int>(n), tc, row_chunk);
    if (pre2_on_side) side_join(s);
    chunk_key_kernel<<<blocks_n, 256, 0, st>>>(pre2->perm, row_chunk, static_cast<int>(n), keys);
    s.launches += 2;
    int bits = 1;
    while ((1 << bits) < nchunks && bits < 31) ++bits;
    sort_pairs<int, int>(s, keys, keys_out, pre2->perm, order, static_cast<int>(n), 0, bits);
    perm = order;
    ps.cell_dim = cell_dim;
    ps.cell_row2 = cell_row2;
    ps.cell_lo = s.dev<double>(nchunks);
    ps.cell_hi = s.dev<double>(nchunks);
    cell_bounds_kernel<<<cdiv(nchunks, 256), 256, 0, st>>>(ps.sorted_keys, n, tc, nchunks, ps.cell_lo, ps.cell_hi);
    s.launches++;
    cells_done = true;
  } else if (pre2 && pre2_on_side) {
    side_join(s);
  }

  ps.P = s.dev<double>(static_cast<size_t>(d) * ps.stride);
  ps.slot_row = s.dev<int>(ps.stride);
  if (cls) {     // (single-segment sets: the gather kernel writes the few padding slots itself)
    CU(cudaMemsetAsync(ps.P, 0xFF, sizeof(double) * d * ps.stride, st));     // all-ones = NaN
    CU(cudaMemsetAsync(ps.slot_row, 0xFF, sizeof(int) * ps.stride, st));     // -1
  }
  GatherArgs ga;
  for (int t = 0; t < d; ++t) ga.rows[t] = row_src[t];
  ga.n = n; ga.d = d; ga.perm = perm; ga.cls_sorted = cls_sorted;
  ga.seg_rank = nullptr; ga.seg_slot = nullptr;
  if (cls) {
    int* h = s.host<int>(2 * nseg);
    std::memcpy(h, seg_rank.data(), sizeof(int) * nseg);
    std::memcpy(h + nseg, ps.seg_slot.data(), sizeof(int) * nseg);
    int* dv = s.dev<int>(2 * nseg);
    CU(cudaMemcpyAsync(dv, h, sizeof(int) * 2 * nseg, cudaMemcpyHostToDevice, st));
    ga.seg_rank = dv;
    ga.seg_slot = dv + nseg;
  }
  ga.P = ps.P; ga.stride = ps.stride; ga.slot_row = ps.slot_row;
  gather_kernel<<<cdiv(cls ? n : ps.stride, 256), 256, 0, st>>>(ga);
  s.launches++;
  CU(cudaGetLastError());
  if (want_cells && !cells_done) {
    const int tc = chunk_len(cell_dim);
    const int nchunks = cdiv(n, tc);
    ps.cell_dim = cell_dim;
    ps.cell_row2 = cell_row2;
    ps.cell_lo = s.dev<double>(nchunks);
    ps.cell_hi = s.dev<double>(nchunks);
    CU(launch_cell_sort(cell_dim, ps.P, ps.stride, d, ps.slot_row, n, sort_row, cell_row2, ps.cell_lo, ps.cell_hi, nchunks, st));
    s.launches++;
  }
  return ps;
}

// Query tiles of a point set whose first row rank lies in [row_lo, row_hi); the candidate segment
// is the tile's own segment (self) or a fixed slice (c_lo, c_len) of another array.
struct TileSet {
  Tile* dev = nullptr;
  int count = 0;
  int64_t rows = 0;
};

TileSet make_tiles(Scratch& s, const PointSet& ps, int64_t row_lo, int64_t row_hi, bool self, int c_lo, int c_len) {
  std::vector<Tile> t;
  int64_t rank = 0, rows = 0;
  // two-level layout: the chunks at the two ends of the across-chunk coordinate lie in sparse regions, all their
  // queries search wide windows over several chunks and their CTAs would form the tail of the kernel: split them
  // into quarter tiles so that the work spreads over four times as many SMs
  int edge_chunks = 0;
  if (self && ps.cell_lo && ps.seg_slot.size() == 1) {
    edge_chunks = 1;
    if (const char* e = getenv("EB2_EDGE_CHUNKS")) edge_chunks = atoi(e);     // tuning knob
  }
  // the table of a single-segment layout depends only on these numbers: reuse the device copy of an earlier call
  int64_t ckey[8] = {ps.seg_len.empty() ? 0 : ps.seg_len[0], ps.seg_slot.empty() ? 0 : ps.seg_slot[0], ps.qpt, row_lo, row_hi,
                     (self ? 1 : 0) | (ps.sort_row >= 0 ? 2 : 0) | (static_cast<int64_t>(edge_chunks) << 8) |
                         (static_cast<int64_t>(ps.cell_lo ? ps.cell_dim : 0) << 40),
                     c_lo, c_len};
  const bool cacheable = ps.seg_slot.size() == 1;
  if (cacheable) {
    for (size_t i = 0; i < s.c.tile_cache.size(); ++i) {
      if (std::memcmp(s.c.tile_cache[i].key, ckey, sizeof ckey) == 0) {
        Ctx::TileEntry e = s.c.tile_cache[i];
        s.c.tile_cache.erase(s.c.tile_cache.begin() + i);
        s.c.tile_cache.push_back(e);                      // most recently used last
        TileSet hit;
        hit.dev = static_cast<Tile*>(e.dev); hit.count = e.count; hit.rows = e.rows;
        return hit;
      }
    }
  }
  for (size_t g = 0; g < ps.seg_slot.size(); ++g) {
    const int tq_full = tile_rows(ps.qpt);
    int tq = tq_full;
    for (int off = 0; off < ps.seg_len[g]; off += tq) {
      tq = tq_full;
      if (edge_chunks > 0) {
        const int tc = chunk_len(ps.cell_dim), ch = off / tc, nch = cdiv(ps.seg_len[g], tc);
        if (ch < edge_chunks || ch >= nch - edge_chunks) tq = tq_full / 4;
      }
      const int qn = std::min(tq, ps.seg_len[g] - off);
      if (rank >= row_lo && rank < row_hi) {
        Tile x;
        x.q_lo = ps.seg_slot[g] + off;
        x.q_n = qn;
        x.c_lo = self ? ps.seg_slot[g] : c_lo;
        x.c_len = self ? ps.seg_len[g] : c_len;
        t.push_back(x);
        rows += qn;
      }
      rank += qn;
    }
  }
  if (ps.sort_row >= 0 && t.size() > 2) {
    // tiles at the two ends of the sorted coordinate sit in sparse regions and search the widest
    // windows: launch them first (CTAs are dispatched in table order) so they do not form the tail
    std::vector<Tile> o;
    o.reserve(t.size());
    for (size_t a = 0, b = t.size() - 1; a <= b && b < t.size(); ++a, --b) {
      o.push_back(t[a]);
      if (b != a) o.push_back(t[b]);
    }
    t.swap(o);
  }
  TileSet ts;
  ts.count = static_cast<int>(t.size());
  ts.rows = rows;
  if (ts.count) {
    Tile* h = s.host<Tile>(t.size());
    std::memcpy(h, t.data(), sizeof(Tile) * t.size());
    if (s.capture) throw CudaFail{cudaErrorInvalidValue, "tile table not cached during graph capture", __LINE__};
    if (cacheable) {
      void* p = nullptr;
      CU(cudaMallocAsync(&p, sizeof(Tile) * t.size(), s.c.stream));      // outlives the call (freed on eviction / shutdown)
      ts.dev = static_cast<Tile*>(p);
    } else {
      ts.dev = s.dev<Tile>(t.size());
    }
    CU(cudaMemcpyAsync(ts.dev, h, sizeof(Tile) * t.size(), cudaMemcpyHostToDevice, s.c.stream));
    if (cacheable) {
      if (s.c.tile_cache.size() >= 12) {
        s.c.drop_graphs();             // a graph may reference the evicted table
        cudaFreeAsync(s.c.tile_cache.front().dev, s.c.stream);
        s.c.tile_cache.erase(s.c.tile_cache.begin());
      }
      Ctx::TileEntry e;
      std::memcpy(e.key, ckey, sizeof ckey);
      e.dev = ts.dev; e.count = ts.count; e.rows = ts.rows;
      s.c.tile_cache.push_back(e);
    }
  }
  return ts;
}

RowSel rows_range(int first, int count) {
  RowSel r{};
  for (int t = 0; t < count; ++t) r.row[t] = first + t;
  return r;
}

void run_knn(Scratch& s, const PointSet& ps, const RowSel& rows, int D, int k, const TileSet& ts, double* eps,
             unsigned long long* pairs) {
  if (!ts.count) return;
  if (D > kMaxDim) {          // wider than the specialised kernels: generic brute force
    GenKnnArgs g;
    g.P = ps.P; g.stride = ps.stride; g.rows = rows; g.d = D; g.tiles = ts.dev; g.ntiles = ts.count; g.k = k;
    g.eps = eps; g.pairs = pairs;
    const int grid = std::min(ts.count, s.c.sm_count * 4);
    g.heap = s.dev<double>(static_cast<size_t>(k + 1) * grid * kThreads);
    knn_generic_kernel<<<grid, kThreads, 0, s.c.stream>>>(g);
    CU(cudaGetLastError());
    s.launches++;
    return;
  }
  if (D == 1 && ps.sort_row >= 0 && rows.row[0] == ps.sort_row && !getenv("EB2_NO_KNN1D")) {
    // one dimension, ascending inside every segment: the k nearest neighbours are among the k predecessors and successors
    Knn1dArgs g;
    g.x = ps.P + static_cast<int64_t>(ps.sort_row) * ps.stride; g.tiles = ts.dev; g.ntiles = ts.count; g.k = k;
    g.eps = eps; g.pairs = pairs;
    knn1d_kernel<<<ts.count, kThreads, 0, s.c.stream>>>(g);
    CU(cudaGetLastError());
    s.launches++;
    return;
  }
  KnnArgs a;
  a.P = ps.P; a.stride = ps.stride; a.rows = rows; a.tiles = ts.dev; a.k = k;
  a.sort_row = -1;
  a.cell_lo = nullptr; a.cell_hi = nullptr;
  // pruning needs the sorted coordinate to be a coordinate of the search space; the kernel reads it
  // as the query's coordinate 0 (the order of coordinates is irrelevant to the max-norm)
  for (int t = 0; t < D && ps.sort_row >= 0; ++t)
    if (rows.row[t] == ps.sort_row) { std::swap(a.rows.row[0], a.rows.row[t]); a.sort_row = ps.sort_row; break; }
  if (a.sort_row >= 0 && ps.cell_lo && ps.cell_dim == D) {
    // two-level layout: the in-chunk coordinate must be coordinate 1 of the search space
    for (int t = 1; t < D; ++t)
      if (a.rows.row[t] == ps.cell_row2) { std::swap(a.rows.row[1], a.rows.row[t]); a.cell_lo = ps.cell_lo; a.cell_hi = ps.cell_hi; break; }
    if (!a.cell_lo) throw CudaFail{cudaErrorInvalidValue, "cell layout does not match the search space", __LINE__};
  } else if (ps.cell_lo) {
    throw CudaFail{cudaErrorInvalidValue, "cell layout used with a different search space", __LINE__};
  }
  a.eps = eps; a.heap = nullptr; a.pairs = pairs; a.ntiles = ts.count;
  a.defer_below = 0; a.left_list = nullptr; a.left_count = nullptr; a.left_best = nullptr;
  const int k1t = (k + 1 <= 4) ? 4 : 8;
  if (a.sort_row >= 0 && k + 1 <= 8) {
    a.defer_below = a.cell_lo ? (D <= 2 ? 4 : 8) : 16;      // measured optima (round 1 sweeps of EB2_DEFER)
    if (const char* e = getenv("EB2_DEFER")) a.defer_below = atoi(e);     // tuning knob
  }
  if (a.defer_below > 0) {
    a.left_list = s.dev<LeftEntry>(ps.stride);
    a.left_count = s.dev<unsigned int>(1);
    a.left_best = s.dev<double>(static_cast<size_t>(ps.stride) * k1t);
    CU(cudaMemsetAsync(a.left_count, 0, sizeof(unsigned int), s.c.stream));
  }
  int lane_dim = 5;
  if (const char* e = getenv("EB2_LANE_DIM")) lane_dim = atoi(e);                          // tuning knob
  a.lane_scan = (a.cell_lo && k + 1 <= 8 && D <= lane_dim && D <= 5) ? (1 << 30) : 0;
  if (const char* e = getenv("EB2_LANE_SCAN")) a.lane_scan = a.lane_scan ? atoi(e) : 0;   // tuning knob (0 = off)
  const int grid = knn_grid(k, ts.count, s.c.sm_count);
  if (k + 1 > 8) a.heap = s.dev<double>(static_cast<size_t>(k + 1) * grid * tile_rows(ps.qpt));
  CU(launch_knn(D, ps.qpt, a, grid, s.c.stream));
  s.launches++;
  if (a.defer_below > 0) {
    // entry count stays on the device: a fixed persistent grid walks the list
    CU(launch_knn_leftover(D, a, s.c.sm_count * 8, s.c.stream));
    s.launches++;
  }
}

struct CountOut {
  int* s = nullptr;
  int* e0 = nullptr;
  int* e1 = nullptr;
};

// queries from `qs` (tiles), candidates from `bs`; shared rows / extra rows are given per set
void run_count(Scratch& s, const PointSet& qs, const PointSet& bs, int C, int E, const RowSel& q_srow,
               const RowSel& b_srow, const RowSel& q_erow, const RowSel& b_erow, const double* radius,
               const TileSet& ts, bool prune, const CountOut& out, unsigned long long* pairs) {
  if (!ts.count) return;
  if (C + E > kMaxDim) {      // generic brute force for wide marginals
    GenCountArgs g;
    g.Q = qs.P; g.qstride = qs.stride; g.B = bs.P; g.bstride = bs.stride;
    g.q_srow = q_srow; g.b_srow = b_srow; g.q_erow = q_erow; g.b_erow = b_erow; g.C = C; g.E = E;
    g.radius = radius; g.tiles = ts.dev; g.ntiles = ts.count;
    g.cnt_s = out.s; g.cnt_e0 = out.e0; g.cnt_e1 = out.e1; g.pairs = pairs;
    count_generic_kernel<<<std::min(ts.count, s.c.sm_count * 8), kThreads, 0, s.c.stream>>>(g);
    CU(cudaGetLastError());
    s.launches++;
    return;
  }
  CountArgs a;
  a.Q = qs.P; a.qstride = qs.stride; a.B = bs.P; a.bstride = bs.stride;
  a.q_srow = q_srow; a.b_srow = b_srow; a.q_erow = q_erow; a.b_erow = b_erow;
  a.radius = radius; a.tiles = ts.dev; a.ntiles = ts.count;
  a.prune_q_row = -1; a.prune_b_row = -1;
  if (prune && bs.sort_row >= 0) {
    // the sorted coordinate of the candidate set must be a SHARED coordinate of the marginals (or the
    // only coordinate); the kernel reads it as the query's shared coordinate 0
    for (int t = 0; t < C; ++t)
      if (b_srow.row[t] == bs.sort_row) {
        std::swap(a.b_srow.row[0], a.b_srow.row[t]);
        std::swap(a.q_srow.row[0], a.q_srow.row[t]);
        a.prune_b_row = bs.sort_row; a.prune_q_row = a.q_srow.row[0];
        break;
      }
    if (C == 0 && E == 1 && b_erow.row[0] == bs.sort_row) { a.prune_b_row = bs.sort_row; a.prune_q_row = q_erow.row[0]; }
  }
  a.cell_lo = nullptr; a.cell_hi = nullptr;
  if (bs.cell_lo) {
    // two-level candidate layout: usable when queries come from the same set, the chunk length matches and
    // the ordering coordinates are shared coordinates 0 and 1 of the marginals; otherwise no pruning at all
    a.prune_q_row = -1; a.prune_b_row = -1;
    if (prune && &qs == &bs && C >= 2 && bs.cell_dim > 0 && chunk_len(bs.cell_dim) == chunk_len(C + E)) {
      int i0 = -1, i1 = -1;
      for (int t = 0; t < C; ++t) {
        if (a.b_srow.row[t] == bs.sort_row) i0 = t;
        if (a.b_srow.row[t] == bs.cell_row2) i1 = t;
      }
      if (i0 >= 0 && i1 >= 0) {
        std::swap(a.b_srow.row[0], a.b_srow.row[i0]); std::swap(a.q_srow.row[0], a.q_srow.row[i0]);
        if (i1 == 0) i1 = i0;
        std::swap(a.b_srow.row[1], a.b_srow.row[i1]); std::swap(a.q_srow.row[1], a.q_srow.row[i1]);
        a.cell_lo = bs.cell_lo; a.cell_hi = bs.cell_hi;
      }
    }
  }
  a.cnt_s = out.s; a.cnt_e0 = out.e0; a.cnt_e1 = out.e1; a.pairs = pairs;
  // Few tiles for the GPU (N = 2*10^5: 800 CTAs of very different lengths on 592 resident slots, the SMs idle half of
  // the launch): the chunk range of every tile is dealt to `split` CTAs that add their integer counts (exact, order-free)
  a.split = 1;
  if (a.prune_b_row >= 0 || a.cell_lo) {
    const int64_t want = static_cast<int64_t>(s.c.sm_count) * 24;
    a.split = static_cast<int>(std::min<int64_t>(4, std::max<int64_t>(1, want / std::max(1, ts.count))));
  }
  if (const char* e = getenv("EB2_COUNT_SPLIT")) a.split = std::max(1, std::min(8, atoi(e)));     // tuning knob
  if (a.split > 1) {
    const PointSet& qp = qs;
    if (out.s) CU(cudaMemsetAsync(out.s, 0, sizeof(int) * qp.stride, s.c.stream));
    if (out.e0) CU(cudaMemsetAsync(out.e0, 0, sizeof(int) * qp.stride, s.c.stream));
    if (out.e1) CU(cudaMemsetAsync(out.e1, 0, sizeof(int) * qp.stride, s.c.stream));
  }
  CU(launch_count(C, E, qs.qpt, a, ts.count * a.split, s.c.stream));
  s.launches++;
}

// one or (qcoord2 != NULL) two 1-D marginals searched by one launch
void run_search(Scratch& s, const double* qcoord, const double* radius, const double* sorted, const TileSet& ts, int* cnt,
                const double* qcoord2 = nullptr, const double* sorted2 = nullptr, int* cnt2 = nullptr, bool from_eps = false) {
  if (!ts.count) return;
  SearchArgs a;
  a.from_eps = from_eps ? 1 : 0;
  a.qcoord = qcoord; a.radius = radius; a.sorted = sorted; a.tiles = ts.dev; a.ntiles = ts.count; a.cnt = cnt;
  a.qcoord2 = qcoord2; a.sorted2 = sorted2; a.cnt2 = cnt2;
  search_kernel<<<dim3(ts.count, qcoord2 ? 2 : 1), kThreads, 0, s.c.stream>>>(a);
  CU(cudaGetLastError());
  s.launches++;
}

constexpr int kPsiTab = 1 << 16;
// psi_ref(n) for n < kPsiTab, filled once per lane by the same device function an evaluation would use
void ensure_psi_tab(Scratch& s) {
  if (s.c.psi_tab) return;
  void* p = nullptr;
  CU(cudaMallocAsync(&p, sizeof(double) * kPsiTab, s.c.stream));       // kept until eb2_shutdown
  s.c.psi_tab = static_cast<double*>(p);
  psi_table_kernel<<<cdiv(kPsiTab, 256), 256, 0, s.c.stream>>>(s.c.psi_tab, kPsiTab);
  s.launches++;
}

// per-tile digamma partials -> 4 doubles on the device
double* run_psi(Scratch& s, int mode, const int* ca, const int* cb, const int* cc, const double* dist, const TileSet& ts) {
  double* out4 = s.result;              // zeroed by begin_call
  if (!out4) {
    out4 = s.dev<double>(4);
    CU(cudaMemsetAsync(out4, 0, sizeof(double) * 4, s.c.stream));
  }
  if (!ts.count) return out4;
  PsiArgs a;
  a.cnt_a = ca; a.cnt_b = cb; a.cnt_c = cc; a.dist = dist; a.tiles = ts.dev; a.ntiles = ts.count; a.mode = mode;
  if (mode != LOG_DIST) ensure_psi_tab(s);
  a.tab = s.c.psi_tab; a.tab_n = s.c.psi_tab ? kPsiTab : 0;
  a.partial = s.dev<double>(static_cast<size_t>(ts.count) * 4);
  psi_kernel<<<ts.count, kThreads, 0, s.c.stream>>>(a);
  psi_final_kernel<<<1, kThreads, 0, s.c.stream>>>(a.partial, ts.count, out4);
  CU(cudaGetLastError());
  s.launches += 2;
  return out4;
}

// Leaves and inner nodes of NumPy's pairwise summation for a range of n elements.  Returns the node id of
// the range: leaves are >= 0 (index into `leaves`), inner nodes are -(index + 1) into `inner`.
struct NpNode { int left, right, height; };
static int np_plan(long long n, long long off, std::vector<NpLeaf>& leaves, std::vector<NpNode>& inner) {
  if (n <= 128) {
    leaves.push_back(NpLeaf{off, static_cast<int>(n)});
    return static_cast<int>(leaves.size()) - 1;
  }
  long long n2 = n / 2;
  n2 -= n2 % 8;
  const int l = np_plan(n2, off, leaves, inner);
  const int r = np_plan(n - n2, off + n2, leaves, inner);
  const int hl = l >= 0 ? 0 : inner[-l - 1].height, hr = r >= 0 ? 0 : inner[-r - 1].height;
  inner.push_back(NpNode{l, r, 1 + std::max(hl, hr)});
  return -static_cast<int>(inner.size());
}

struct StatsWindow {
  const double* src;
  int64_t off, stride;
  double* out4;          // device: sum, mean, sum of squared deviations, std
};

// mean (out4[1]) and standard deviation (out4[3]) of src[off + i*stride], i < n, for every window, with NumPy's
// association, left on the device; four small kernels per group of kNpCols windows on the lane's stream
void device_stats(Scratch& s, const StatsWindow* win, int nwin, int64_t n) {
  Ctx& c = s.c;
  const Ctx::NpTables* tb = nullptr;
  for (const auto& e : c.np_tables)
    if (e.n == n) tb = &e;
  if (!tb) {
    std::vector<NpLeaf> leaves;
    std::vector<NpNode> inner;
    leaves.reserve(static_cast<size_t>(n / 64 + 2));
    inner.reserve(static_cast<size_t>(n / 64 + 2));
    np_plan(n, 0, leaves, inner);
    const int nleaves = static_cast<int>(leaves.size()), ninner = static_cast<int>(inner.size());
    // inner nodes by height (counting sort; the post-order root is the highest and stays last)
    Ctx::NpTables e{};
    e.n = n; e.nleaves = nleaves; e.ninner = ninner;
    int maxh = 0;
    for (const auto& nd : inner) maxh = std::max(maxh, nd.height);
    if (maxh > kNpMaxLevels) throw CudaFail{cudaErrorInvalidValue, "summation tree too deep", __LINE__};
    e.levels.nlevels = maxh;
    std::vector<int> first(maxh + 2, 0);                      // first[h] = first slot of the nodes of height h
    for (const auto& nd : inner) first[nd.height + 1]++;
    for (int h = 1; h <= maxh + 1; ++h) first[h] += first[h - 1];
    for (int h = 1; h <= maxh; ++h) e.levels.start[h - 1] = first[h];
    e.levels.start[maxh] = ninner;
    std::vector<int> slot(ninner);
    std::vector<int> next(first);
    for (int i = 0; i < ninner; ++i) slot[i] = next[inner[i].height]++;
    int2* hc = s.host<int2>(std::max(ninner, 1));
    auto id = [&](int v) { return v >= 0 ? v : nleaves + slot[-v - 1]; };
    for (int i = 0; i < ninner; ++i) hc[slot[i]] = make_int2(id(inner[i].left), id(inner[i].right));
    NpLeaf* hl = s.host<NpLeaf>(leaves.size());
    std::memcpy(hl, leaves.data(), sizeof(NpLeaf) * leaves.size());
    if (c.np_tables.size() >= 4) {
      cudaFreeAsync(c.np_tables.front().leaves, c.stream);
      cudaFreeAsync(c.np_tables.front().children, c.stream);
      c.np_tables.erase(c.np_tables.begin());
    }
    CU(cudaMallocAsync(&e.leaves, sizeof(NpLeaf) * leaves.size(), c.stream));
    CU(cudaMallocAsync(&e.children, sizeof(int2) * std::max(ninner, 1), c.stream));
    CU(cudaMemcpyAsync(e.leaves, hl, sizeof(NpLeaf) * leaves.size(), cudaMemcpyHostToDevice, c.stream));
    CU(cudaMemcpyAsync(e.children, hc, sizeof(int2) * std::max(ninner, 1), cudaMemcpyHostToDevice, c.stream));
    c.np_tables.push_back(e);
    tb = &c.np_tables.back();
  }
  const NpLeaf* dl = static_cast<const NpLeaf*>(tb->leaves);
  const int2* dc = static_cast<const int2*>(tb->children);
  const int blocks = cdiv(static_cast<int64_t>(tb->nleaves) * 8, 256);
  for (int w0 = 0; w0 < nwin; w0 += kNpCols) {
    const int m = std::min(kNpCols, nwin - w0);
    NpCols a{};
    for (int w = 0; w < m; ++w) {
      a.src[w] = win[w0 + w].src; a.off[w] = win[w0 + w].off; a.stride[w] = win[w0 + w].stride;
      a.out[w] = win[w0 + w].out4;
      a.val[w] = s.dev<double>(static_cast<size_t>(tb->nleaves) + tb->ninner);
    }
    for (int mode = 0; mode < 2; ++mode) {
      np_leaf_sum_kernel<<<dim3(blocks, m), 256, 0, c.stream>>>(a, dl, tb->nleaves, mode);
      np_tree_kernel<<<m, 1024, 0, c.stream>>>(a, dc, tb->levels, tb->nleaves, n, mode);
    }
    CU(cudaGetLastError());
    s.launches += 4;
  }
}

// ---- input staging -------------------------------------------------------------------------------
// The d x n block the estimators work on comes from (a) a host block (one H2D copy), (b) a device
// block (EB2_FLAG_DEVICE_INPUT), or (c) cached device columns, sliced / strided / rescaled by
// prep_kernel (the reference's _rescale_data arithmetic, ennemi/_driver.py:871-902, on the device).
struct Input {
  const double* coords = nullptr;
  const eb2_col_t* cols = nullptr;
  uint32_t flags = 0;
};

const double* stage_input(Scratch& s, const Input& in, int d, int64_t n, int* nonfinite_flag, bool check_finite = true) {
  const double* raw = in.coords;
  const int64_t total = static_cast<int64_t>(d) * n;
  if (in.cols) {
    PrepArgs pa;
    pa.d = d; pa.n = n; pa.flags = nonfinite_flag;
    std::lock_guard<std::mutex> cache_guard(s.c.shared->mu);
    auto& cache = s.c.shared->cache;
    std::vector<StatsWindow> windows;
    for (int t = 0; t < d; ++t) {
      const eb2_col_t& c = in.cols[t];
      auto it = cache.find(c.key);
      if (it == cache.end()) throw CudaFail{cudaErrorInvalidValue, "column key not in the device cache", __LINE__};
      const int64_t last = c.off + (n - 1) * c.stride;
      if (c.off < 0 || last < 0 || c.off >= it->second.second || last >= it->second.second)
        throw CudaFail{cudaErrorInvalidValue, "column slice outside the cached column", __LINE__};
      PrepCol& pc = pa.col[t];
      pc.src = it->second.first; pc.off = c.off; pc.stride = c.stride; pc.mean = c.mean; pc.std = c.std;
      pc.noise = nullptr; pc.noff = c.noff; pc.nstride = c.nstride; pc.dstats = nullptr; pc.flag = nullptr; pc.dst = nullptr;
      if (c.nkey != 0) {
        auto nt = cache.find(c.nkey);
        if (nt == cache.end()) throw CudaFail{cudaErrorInvalidValue, "noise key not in the device cache", __LINE__};
        const int64_t nlast = c.noff + (n - 1) * c.nstride;
        if (c.noff < 0 || nlast < 0 || nlast >= nt->second.second)
          throw CudaFail{cudaErrorInvalidValue, "noise slice outside the cached vector", __LINE__};
        pc.noise = nt->second.first;
      }
      if ((in.flags & EB2_FLAG_DEVICE_STATS) && c.std != 0.0 && c.mean != c.mean) {
        // statistics of the window computed here, where the column is: no host round trip before the task
        double* st4 = s.dev<double>(4);
        windows.push_back(StatsWindow{pc.src, c.off, c.stride, st4});
        pc.dstats = st4;
      }
    }
    if (!windows.empty()) device_stats(s, windows.data(), static_cast<int>(windows.size()), n);
    double* dv = s.dev<double>(static_cast<size_t>(total));
    pa.raw = dv;
    prep_kernel<<<cdiv(total, 256), 256, 0, s.c.stream>>>(pa);
    s.launches++;
    raw = dv;
  } else if (!(in.flags & EB2_FLAG_DEVICE_INPUT)) {
    double* dv = s.dev<double>(static_cast<size_t>(total));
    if (s.capture) CU(cudaMemcpyAsync(dv, in.coords, sizeof(double) * total, cudaMemcpyHostToDevice, s.c.stream));
    else upload(s.c, dv, in.coords, sizeof(double) * total);
    raw = dv;
  }
  if (check_finite) {
    nonfinite_kernel<<<cdiv(total, 256), 256, 0, s.c.stream>>>(raw, total, nonfinite_flag);
    s.launches++;
  }
  return raw;
}

const double* stage_coords(Scratch& s, const double* coords, int d, int64_t n, uint32_t flags, int* nonfinite_flag) {
  Input in;
  in.coords = coords;
  in.flags = flags;
  return stage_input(s, in, d, n, nonfinite_flag);
}

// Prepared variable for a column descriptor, shared across the tasks (and stream lanes) of a call.
// Created on first use on the caller's stream; other lanes wait on its event.
Derived* get_derived(Scratch& s, const eb2_col_t& col, int64_t n) {
  struct DKey { eb2_col_t c; int64_t n; } k;
  std::memset(&k, 0, sizeof k);
  k.c = col; k.n = n;
  const std::string key(reinterpret_cast<const char*>(&k), sizeof k);
  DevShared& sh = *s.c.shared;
  std::lock_guard<std::mutex> guard(sh.mu);
  auto it = sh.derived.find(key);
  if (it == sh.derived.end()) {
    auto src = sh.cache.find(col.key);
    if (src == sh.cache.end()) throw CudaFail{cudaErrorInvalidValue, "column key not in the device cache", __LINE__};
    const int64_t last = col.off + (n - 1) * col.stride;
    if (col.off < 0 || last < 0 || col.off >= src->second.second || last >= src->second.second)
      throw CudaFail{cudaErrorInvalidValue, "column slice outside the cached column", __LINE__};
    PrepArgs pa;
    pa.d = 1; pa.n = n;
    PrepCol& pc = pa.col[0];
    pc.src = src->second.first; pc.off = col.off; pc.stride = col.stride; pc.mean = col.mean; pc.std = col.std;
    pc.noise = nullptr; pc.noff = col.noff; pc.nstride = col.nstride; pc.dstats = nullptr; pc.flag = nullptr; pc.dst = nullptr;
    if (col.nkey != 0) {
      auto nt = sh.cache.find(col.nkey);
      if (nt == sh.cache.end()) throw CudaFail{cudaErrorInvalidValue, "noise key not in the device cache", __LINE__};
      const int64_t nlast = col.noff + (n - 1) * col.nstride;
      if (col.noff < 0 || nlast < 0 || nlast >= nt->second.second)
        throw CudaFail{cudaErrorInvalidValue, "noise slice outside the cached vector", __LINE__};
      pc.noise = nt->second.first;
    }
    Derived d;
    cudaStream_t st = s.c.stream;
    d.n = n; d.src_key = col.key; d.noise_key = col.nkey;
    try {
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d.vals), sizeof(double) * n, st));
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d.sorted), sizeof(double) * n, st));
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d.perm), sizeof(int) * n, st));
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d.dflag), sizeof(int), st));
      CU(cudaMemsetAsync(d.dflag, 0, sizeof(int), st));
      pa.raw = d.vals; pa.flags = d.dflag;
      prep_kernel<<<cdiv(n, 256), 256, 0, st>>>(pa);
      nonfinite_kernel<<<cdiv(n, 256), 256, 0, st>>>(d.vals, n, d.dflag);
      int* iota = s.dev<int>(n);
      iota_kernel<<<cdiv(n, 256), 256, 0, st>>>(iota, static_cast<int>(n));
      s.launches += 3;
      sort_pairs<double, int>(s, d.vals, d.sorted, iota, d.perm, static_cast<int>(n), 0, 64);
      CU(cudaEventCreateWithFlags(&d.ready, cudaEventDisableTiming));
      CU(cudaEventRecord(d.ready, st));
    } catch (...) {        // nothing half-built stays behind
      if (d.vals) cudaFreeAsync(d.vals, st);
      if (d.sorted) cudaFreeAsync(d.sorted, st);
      if (d.perm) cudaFreeAsync(d.perm, st);
      if (d.dflag) cudaFreeAsync(d.dflag, st);
      if (d.ready) cudaEventDestroy(d.ready);
      throw;
    }
    it = sh.derived.emplace(key, d).first;
  }
  CU(cudaStreamWaitEvent(s.c.stream, it->second.ready, 0));
  return &it->second;
}

void drop_derived_locked(DevShared& sh, uint64_t key, cudaStream_t st) {
  for (auto it = sh.derived.begin(); it != sh.derived.end();) {
    if (key == 0 || it->second.src_key == key || it->second.noise_key == key) {
      cudaFreeAsync(it->second.vals, st); cudaFreeAsync(it->second.sorted, st);
      cudaFreeAsync(it->second.perm, st); cudaFreeAsync(it->second.dflag, st);
      cudaEventDestroy(it->second.ready);
      it = sh.derived.erase(it);
    } else {
      ++it;
    }
  }
}

struct Outputs {
  double* eps = nullptr;     // host, caller's row order
  int64_t* cnt[3] = {nullptr, nullptr, nullptr};
};

// copies slot-order results to the caller's arrays (row order)
void export_outputs(Scratch& s, const PointSet& ps, const double* eps, const int* const cnt[3], const Outputs& o) {
  cudaStream_t st = s.c.stream;
  const int blocks = cdiv(ps.stride, 256);
  if (o.eps) {
    double* tmp = s.dev<double>(ps.n);
    CU(cudaMemsetAsync(tmp, 0xFF, sizeof(double) * ps.n, st));
    scatter_f64_kernel<<<blocks, 256, 0, st>>>(eps, ps.slot_row, ps.stride, tmp);
    s.launches++;
    CU(cudaMemcpyAsync(o.eps, tmp, sizeof(double) * ps.n, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 3; ++i) {
    if (o.cnt[i] && cnt[i]) {
      long long* tmp = s.dev<long long>(ps.n);
      CU(cudaMemsetAsync(tmp, 0xFF, sizeof(long long) * ps.n, st));
      scatter_i64_kernel<<<blocks, 256, 0, st>>>(cnt[i], ps.slot_row, ps.stride, tmp);
      s.launches++;
      CU(cudaMemcpyAsync(o.cnt[i], tmp, sizeof(long long) * ps.n, cudaMemcpyDeviceToHost, st));
    }
  }
}

// gathers sums + work counter + non-finite flag, synchronises, fills the partial block and timings
thread_local bool g_skip_timing = false;   // set by batched calls for all but their last task

struct Res { double v[4]; unsigned long long pairs; int nonfinite; int pad; unsigned long long rows; double spare;
             unsigned long long fix_lo; long long fix_hi; };
static_assert(offsetof(Res, pairs) == 32 && offsetof(Res, nonfinite) == 40 && offsetof(Res, rows) == 48 &&
              offsetof(Res, fix_lo) == 64 && offsetof(Res, fix_hi) == 72 && sizeof(Res) == 80,
              "Res mirrors the device result block");
constexpr int kResDoubles = 16;

// the synchronised result block -> error code / partial block
int parse_result(const Res* h, int64_t rows, double* partial) {
  const bool fixed = rows < 0;
  if (fixed) rows = static_cast<int64_t>(h->rows);      // the bivariate pipeline counts the rows it reduced on the device
  if (h->nonfinite & k2::kFlagOverflow) return EB2_ERR_RETRY_GENERAL;     // internal: the caller repeats on the general path
  if (h->nonfinite & 8) {
    g_data_flags = h->nonfinite;
    return fail(EB2_ERR_CONSTANT, "a window with device-computed statistics is constant (std < 1e-20): "
                                  "repeat the task with host-checked statistics");
  }
  if (h->nonfinite) {
    g_data_flags = h->nonfinite;
    return fail(EB2_ERR_NONFINITE, "data must be finite, check for nan or inf values");
  }
  if (partial) {
    for (int i = 0; i < EB2_P_LEN; ++i) partial[i] = 0.0;
    partial[EB2_P_SUM] = h->v[0];
    partial[EB2_P_ZERO_A] = h->v[1];
    partial[EB2_P_ZERO_B] = h->v[2];
    partial[EB2_P_ZERO_C] = h->v[3];
    partial[EB2_P_ROWS] = static_cast<double>(rows);
    partial[EB2_P_PAIRS] = static_cast<double>(h->pairs);
    if (fixed) {
      // the bivariate pipeline's exact sum: 128-bit two's complement in units of 2^-48, as four 32-bit limbs that
      // doubles carry (and add, over ranks) without rounding
      partial[EB2_P_FIX0] = static_cast<double>(h->fix_lo & 0xffffffffull);
      partial[EB2_P_FIX0 + 1] = static_cast<double>(h->fix_lo >> 32);
      partial[EB2_P_FIX0 + 2] = static_cast<double>(static_cast<unsigned long long>(h->fix_hi) & 0xffffffffull);
      partial[EB2_P_FIX0 + 3] = static_cast<double>(h->fix_hi >> 32);
      partial[EB2_P_FIXED] = 1.0;
    }
  }
  return EB2_OK;
}

// phase times of the call that just completed on the lane's stream (events of a replayed graph are external
// record nodes; should a driver refuse the query, the previous figures simply stay)
void read_phase_times(Ctx& c) {
  static const int pairs[5][2] = {{0, 5}, {1, 2}, {2, 3}, {3, 4}, {0, 1}};
  for (int i = 0; i < 5; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c.ev[pairs[i][0]], c.ev[pairs[i][1]]) == cudaSuccess) c.last_ms[i] = ms;
    else cudaGetLastError();
  }
}

void record_event(Scratch& s, cudaEvent_t ev) {
  if (s.capture) CU(cudaEventRecordWithFlags(ev, s.c.stream, cudaEventRecordExternal));
  else CU(cudaEventRecord(ev, s.c.stream));
}

// gathers sums + work counter + non-finite flag, synchronises, fills the partial block and timings
int finish_call(Scratch& s, const double* out4, const unsigned long long* pairs, const int* nonfinite, int64_t rows,
                double* partial) {
  Ctx& c = s.c;
  Res* h = s.host<Res>(1);
  if (out4 == s.result && s.result) {
    CU(cudaMemcpyAsync(h, s.result, sizeof(Res), cudaMemcpyDeviceToHost, c.stream));     // the whole block in one copy
  } else {
    CU(cudaMemcpyAsync(h->v, out4, sizeof(double) * 4, cudaMemcpyDeviceToHost, c.stream));
    CU(cudaMemcpyAsync(&h->pairs, pairs, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
    CU(cudaMemcpyAsync(&h->nonfinite, nonfinite, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  }
  const bool timing = !g_skip_timing;
  if (timing) record_event(s, c.ev[5]);
  if (s.capture) {
    // everything above was recorded, not run: close the capture, keep the executable graph, run it once
    Ctx::GraphEntry* ge = s.capture;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamEndCapture(c.stream, &graph));
    s.capture = nullptr;
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) throw CudaFail{e, "cudaGraphInstantiate", __LINE__};
    ge->exec = exec;
    ge->host_result = h;
    ge->rows = rows;
    ge->launches = s.launches;
    CU(cudaGraphLaunch(exec, c.stream));
  }
  CU(cudaStreamSynchronize(c.stream));
  if (timing) read_phase_times(c);
  c.last_launches = s.launches;
  return parse_result(h, rows, partial);
}

// replays the captured estimate: one launch, one synchronisation
int graph_replay(Ctx& c, Ctx::GraphEntry& ge, double* partial) {
  CU(cudaSetDevice(c.dev));
  CU(cudaGraphLaunch(ge.exec, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  read_phase_times(c);
  c.last_launches = ge.launches;
  return parse_result(static_cast<const Res*>(ge.host_result), ge.rows, partial);
}

// closes a capture that went wrong (nothing was executed) and forgets the error
void abort_capture(Ctx& c) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(c.stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture(c.stream, &graph);
    if (graph) cudaGraphDestroy(graph);
  }
  cudaGetLastError();
}

bool graph_enabled() {
  const char* e = getenv("EB2_GRAPH");      // EB2_GRAPH=0: always launch kernel by kernel
  return !e || atoi(e) != 0;
}

// entry of this call signature in the lane's graph cache (created on first sight; most recently used last)
Ctx::GraphEntry* graph_entry(Ctx& c, const int64_t (&key)[10]) {
  for (size_t i = 0; i < c.graphs.size(); ++i) {
    if (std::memcmp(c.graphs[i].key, key, sizeof key) == 0) {
      if (i + 1 != c.graphs.size()) {
        Ctx::GraphEntry e = c.graphs[i];
        c.graphs.erase(c.graphs.begin() + i);
        c.graphs.push_back(e);
      }
      return &c.graphs.back();
    }
  }
  if (c.graphs.size() >= 4) {
    if (c.graphs.front().exec) cudaGraphExecDestroy(c.graphs.front().exec);
    c.graphs.erase(c.graphs.begin());
  }
  Ctx::GraphEntry e;
  std::memcpy(e.key, key, sizeof key);
  c.graphs.push_back(e);
  return &c.graphs.back();
}

struct CallInit {
  unsigned long long* pairs;
  int* nonfinite;
};
CallInit begin_call(Scratch& s) {
  Ctx& c = s.c;
  CU(cudaSetDevice(c.dev));
  if (!g_skip_timing) record_event(s, c.ev[0]);
  CallInit ci;
  s.result = s.dev<double>(kResDoubles);
  CU(cudaMemsetAsync(s.result, 0, sizeof(double) * kResDoubles, c.stream));
  ci.pairs = reinterpret_cast<unsigned long long*>(s.result + 4);
  ci.nonfinite = reinterpret_cast<int*>(s.result + 5);
  return ci;
}
void mark(Scratch& s, int ev) {
  if (!g_skip_timing) record_event(s, s.c.ev[ev]);
}

// host copy of the reference's _psi for scalars (_entropy_estimators.py:327-350)
double psi_host(double y) {
  if (y == 0.0) return std::numeric_limits<double>::infinity();
  if (y == 1.0) return -0.5772156649015331;
  const double y2 = y * y;
  return std::log(y) - std::pow(y, -6.0) * (y2 * (y2 * (y / 2 + 1.0 / 12) - 1.0 / 120) + 1.0 / 252);
}

// mean over rows of (psi(a) + psi(b) - psi(c)) given the finite sum and the zero counters:
// a zero count anywhere in an array turns that array's psi into a scalar +inf in the reference
// the digamma sum of a partial block: the bivariate pipeline's exact fixed-point total when present
double partial_sum(const double* partial) {
  if (!(partial[EB2_P_FIXED] > 0)) return partial[EB2_P_SUM];
  __int128 total = 0;
  for (int i = 3; i >= 0; --i) total = (total << 32) + static_cast<__int128>(std::llround(partial[EB2_P_FIX0 + i]));
  return std::ldexp(static_cast<double>(total), -48);
}

double psi_mean(const double* partial, int64_t n) {
  const double inf = std::numeric_limits<double>::infinity();
  const bool za = partial[EB2_P_ZERO_A] > 0, zb = partial[EB2_P_ZERO_B] > 0, zc = partial[EB2_P_ZERO_C] > 0;
  if (!(za || zb || zc)) return partial_sum(partial) / static_cast<double>(n);
  return (za ? inf : 0.0) + (zb ? inf : 0.0) - (zc ? inf : 0.0);   // inf - inf = nan, as numpy gives
}

int check_common(const double* coords, int64_t n, int d, int k) {
  if (!coords) return fail(EB2_ERR_ARG, "coords is NULL");
  if (n <= 0 || n >= (int64_t(1) << 31) - 1024) return fail(EB2_ERR_ARG, "n out of range");
  if (d < 1) return fail(EB2_ERR_ARG, "dimension must be positive");
  if (d > EB2_MAX_DIM) return fail(EB2_ERR_UNSUPPORTED, "dimension %d above EB2_MAX_DIM=%d", d, EB2_MAX_DIM);
  if (k <= 0) return fail(EB2_ERR_ARG, "k must be greater than zero");
  return EB2_OK;
}

// class ids (host or device) -> device copy + per-class sizes; false on an id outside [0, ncls)
bool stage_classes(Scratch& s, const int32_t* cls, int64_t n, int ncls, uint32_t flags, const int** cls_dev,
                   std::vector<int>* sizes) {
  std::vector<int32_t> hbuf;
  const int32_t* hc = cls;
  if (flags & EB2_FLAG_DEVICE_INPUT) {
    hbuf.resize(n);
    CU(cudaMemcpyAsync(hbuf.data(), cls, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.c.stream));
    CU(cudaStreamSynchronize(s.c.stream));
    hc = hbuf.data();
    *cls_dev = cls;
  } else {
    int* dv = s.dev<int>(n);
    CU(cudaMemcpyAsync(dv, cls, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s.c.stream));
    *cls_dev = dv;
  }
  sizes->assign(ncls, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (hc[i] < 0 || hc[i] >= ncls) return false;
    (*sizes)[hc[i]]++;
  }
  return true;
}

template <typename F>
int guarded(int dev, F&& body) {
  try {
    Ctx& c = get_ctx(dev);
    std::lock_guard<std::mutex> g(c.mu);
    return body(c);
  } catch (const CudaFail& f) {
    return fail(EB2_ERR_CUDA, "CUDA error %d (%s) in %s at eb2_lib.cu:%d", (int)f.e, cudaGetErrorString(f.e), f.what, f.line);
  } catch (const std::bad_alloc&) {
    return fail(EB2_ERR_CUDA, "host allocation failed");
  }
}

}  // namespace

// =================================================================================================
extern "C" {

const char* eb2_last_error(void) { return g_err.c_str(); }
const char* eb2_version(void) { return "ennemi_b200 0.1 (sm_100a)"; }

int eb2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int eb2_init(void) {
  const int n = eb2_device_count();
  if (n == 0) return fail(EB2_ERR_CUDA, "no CUDA device available");
  try {
    for (int d = 0; d < n && d < kMaxDev; ++d) get_ctx(d);
  } catch (const CudaFail& f) {
    return fail(EB2_ERR_CUDA, "CUDA error %d (%s) in %s", (int)f.e, cudaGetErrorString(f.e), f.what);
  }
  return EB2_OK;
}

int eb2_shutdown(void) {
  std::lock_guard<std::mutex> g(g_init_mu);
  for (int d = 0; d < kMaxDev * kMaxLanes; ++d) {
    Ctx& c = g_ctx[d];
    if (!c.ready.load(std::memory_order_acquire)) continue;
    std::lock_guard<std::mutex> g2(c.mu);
    cudaSetDevice(c.dev);
    cudaStreamSynchronize(c.stream);
    {
      std::lock_guard<std::mutex> cache_guard(c.shared->mu);
      drop_derived_locked(*c.shared, 0, c.stream);
      for (auto& kv : c.shared->cache) cudaFreeAsync(kv.second.first, c.stream);
      c.shared->cache.clear();
    }
    if (c.arena) cudaFreeAsync(c.arena, c.stream);
    c.arena = nullptr; c.arena_cap = 0;
    if (c.psi_tab) cudaFreeAsync(c.psi_tab, c.stream);
    c.psi_tab = nullptr;
    c.drop_graphs();
    for (auto& te : c.tile_cache) cudaFreeAsync(te.dev, c.stream);
    c.tile_cache.clear();
    for (auto& tb : c.np_tables) { cudaFreeAsync(tb.leaves, c.stream); cudaFreeAsync(tb.children, c.stream); }
    c.np_tables.clear();
    cudaStreamSynchronize(c.stream);
    for (auto& e : c.ev) cudaEventDestroy(e);
    cudaEventDestroy(c.fork); cudaEventDestroy(c.join);
    cudaStreamDestroy(c.side);
    cudaStreamDestroy(c.stream);
    cudaFreeHost(c.pinned);
    c.pinned = nullptr;
    if (c.bounce) {
      cudaFreeHost(c.bounce);
      c.bounce = nullptr;
      for (auto& e : c.bounce_ev) cudaEventDestroy(e);
    }
    c.ready.store(false, std::memory_order_release);
  }
  return EB2_OK;
}

int eb2_last_timing(int dev, double* ms, int* launches) {
  const int ord = dev & 0xff, lane = (dev >> 8) & 0xff;
  if (dev < 0 || ord >= kMaxDev || lane >= kMaxLanes || !g_ctx[ord * kMaxLanes + lane].ready)
    return fail(EB2_ERR_ARG, "device %d has no context", dev);
  const Ctx& c = g_ctx[ord * kMaxLanes + lane];
  if (ms) for (int i = 0; i < 5; ++i) ms[i] = c.last_ms[i];
  if (launches) *launches = c.last_launches;
  return EB2_OK;
}

int eb2_last_pipeline(int dev) {
  const int ord = dev & 0xff, lane = (dev >> 8) & 0xff;
  if (dev < 0 || ord >= kMaxDev || lane >= kMaxLanes || !g_ctx[ord * kMaxLanes + lane].ready) return -1;
  return g_ctx[ord * kMaxLanes + lane].last_pipeline;
}

// ---- device column cache ---------------------------------------------------------------------------
int eb2_cache_put(int dev, uint64_t key, const double* host, int64_t n) {
  if (key == 0 || !host || n <= 0) return fail(EB2_ERR_ARG, "eb2_cache_put: bad argument");
  return guarded(dev, [&](Ctx& c) {
    CU(cudaSetDevice(c.dev));
    double* p = nullptr;
    CU(cudaMallocAsync(reinterpret_cast<void**>(&p), sizeof(double) * n, c.stream));   // from the cached pool
    try {
      upload(c, p, host, sizeof(double) * n);                       // complete on return: the caller may reuse `host`
    } catch (...) {
      cudaFreeAsync(p, c.stream);
      throw;
    }
    std::lock_guard<std::mutex> cache_guard(c.shared->mu);
    auto it = c.shared->cache.find(key);
    if (it != c.shared->cache.end()) {
      cudaFreeAsync(it->second.first, c.stream);
      c.shared->cache.erase(it);
    }
    c.shared->cache[key] = {p, n};
    return EB2_OK;
  });
}

// One H2D copy of a row-major (n x ncols) block (row stride ld >= ncols elements) and a device-side
// de-interleave into ncols cached columns: replaces ncols strided gathers on the host (pairwise_mi on an
// (n, nvar) array spent more time transposing on the CPU than estimating on the GPU).
static int cache_put_block_impl(int dev, const uint64_t* keys, int ncols, const double* src, int64_t n, int64_t ld, bool on_device) {
  if (!keys || !src || ncols <= 0 || n <= 0 || ld < ncols) return fail(EB2_ERR_ARG, "eb2_cache_put_block: bad argument");
  for (int j = 0; j < ncols; ++j)
    if (keys[j] == 0) return fail(EB2_ERR_ARG, "eb2_cache_put_block: key 0 is reserved");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    const size_t count = static_cast<size_t>(n) * ncols;
    const double* block = src;
    int64_t bld = ld;
    if (!on_device) {
      double* staged = s.dev<double>(count);
      if (ld == ncols)
        upload(c, staged, src, sizeof(double) * count);
      else
        CU(cudaMemcpy2DAsync(staged, sizeof(double) * ncols, src, sizeof(double) * ld, sizeof(double) * ncols,
                             static_cast<size_t>(n), cudaMemcpyHostToDevice, c.stream));
      block = staged;
      bld = ncols;
    }
    std::vector<double*> cols(ncols, nullptr);
    try {
      for (int j = 0; j < ncols; ++j)
        CU(cudaMallocAsync(reinterpret_cast<void**>(&cols[j]), sizeof(double) * n, c.stream));
      double** table_h = s.host<double*>(ncols);
      std::memcpy(table_h, cols.data(), sizeof(double*) * ncols);
      double** table_d = s.dev<double*>(ncols);
      CU(cudaMemcpyAsync(table_d, table_h, sizeof(double*) * ncols, cudaMemcpyHostToDevice, c.stream));
      deinterleave_kernel<<<dim3(cdiv(n, 32), cdiv(ncols, 32)), dim3(32, 8), 0, c.stream>>>(block, n, ncols, bld, table_d);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(c.stream));      // the caller may reuse the source right away
    } catch (...) {
      for (double* p : cols)
        if (p) cudaFreeAsync(p, c.stream);
      throw;
    }
    c.last_launches = 1;
    std::lock_guard<std::mutex> cache_guard(c.shared->mu);
    for (int j = 0; j < ncols; ++j) {
      auto it = c.shared->cache.find(keys[j]);
      if (it != c.shared->cache.end()) {
        cudaFreeAsync(it->second.first, c.stream);
        c.shared->cache.erase(it);
      }
      c.shared->cache[keys[j]] = {cols[j], n};
    }
    return EB2_OK;
  });
}

int eb2_cache_put_block(int dev, const uint64_t* keys, int ncols, const double* host, int64_t n, int64_t ld) {
  return cache_put_block_impl(dev, keys, ncols, host, n, ld, false);
}

// the same from a block that already is in device memory on `dev` (e.g. row slices uploaded by the ranks of a job and
// all-gathered over NVLink: every rank pays for 1/G of the host-to-device copy)
int eb2_cache_put_block_dev(int dev, const uint64_t* keys, int ncols, const double* dev_block, int64_t n, int64_t ld) {
  return cache_put_block_impl(dev, keys, ncols, dev_block, n, ld, true);
}

int eb2_cache_drop(int dev, uint64_t key) {
  return guarded(dev, [&](Ctx& c) {
    CU(cudaSetDevice(c.dev));
    std::lock_guard<std::mutex> cache_guard(c.shared->mu);
    auto& cache = c.shared->cache;
    drop_derived_locked(*c.shared, key, c.stream);
    if (key == 0) {
      for (auto& kv : cache) cudaFreeAsync(kv.second.first, c.stream);
      cache.clear();
    } else {
      auto it = cache.find(key);
      if (it != cache.end()) { cudaFreeAsync(it->second.first, c.stream); cache.erase(it); }
    }
    return EB2_OK;
  });
}

int eb2_last_data_flags(void) { return g_data_flags; }

int eb2_cache_stats(int dev, uint64_t key, int64_t off, int64_t stride, int64_t n, double* mean, double* std_out) {
  if (key == 0 || n <= 0 || stride == 0 || !mean || !std_out) return fail(EB2_ERR_ARG, "eb2_cache_stats: bad argument");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    const double* src = nullptr;
    {
      std::lock_guard<std::mutex> cache_guard(c.shared->mu);
      auto it = c.shared->cache.find(key);
      if (it == c.shared->cache.end()) return fail(EB2_ERR_ARG, "eb2_cache_stats: column key not in the device cache");
      const int64_t last = off + (n - 1) * stride;
      if (off < 0 || last < 0 || off >= it->second.second || last >= it->second.second)
        return fail(EB2_ERR_ARG, "eb2_cache_stats: window outside the cached column");
      src = it->second.first;
    }
    double* out = s.dev<double>(4);
    const StatsWindow win{src, off, stride, out};
    device_stats(s, &win, 1, n);
    double* h = s.host<double>(4);
    CU(cudaMemcpyAsync(h, out, sizeof(double) * 4, cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    *mean = h[1];
    *std_out = h[3];
    c.last_launches = s.launches;
    return EB2_OK;
  });
}

// eb2_cache_stats for nwin windows of one length in one call (one group of launches, one read-back)
int eb2_cache_stats_many(int dev, const uint64_t* keys, const int64_t* offs, int nwin, int64_t stride, int64_t n,
                         double* means, double* stds) {
  if (!keys || !offs || nwin <= 0 || n <= 0 || stride == 0 || !means || !stds)
    return fail(EB2_ERR_ARG, "eb2_cache_stats_many: bad argument");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    double* out = s.dev<double>(static_cast<size_t>(nwin) * 4);
    std::vector<StatsWindow> wins(nwin);
    {
      std::lock_guard<std::mutex> cache_guard(c.shared->mu);
      for (int w = 0; w < nwin; ++w) {
        auto it = c.shared->cache.find(keys[w]);
        if (it == c.shared->cache.end()) return fail(EB2_ERR_ARG, "eb2_cache_stats_many: column key not in the device cache");
        const int64_t last = offs[w] + (n - 1) * stride;
        if (offs[w] < 0 || last < 0 || offs[w] >= it->second.second || last >= it->second.second)
          return fail(EB2_ERR_ARG, "eb2_cache_stats_many: window outside the cached column");
        wins[w] = StatsWindow{it->second.first, offs[w], stride, out + 4 * w};
      }
    }
    device_stats(s, wins.data(), nwin, n);
    double* h = s.host<double>(static_cast<size_t>(nwin) * 4);
    CU(cudaMemcpyAsync(h, out, sizeof(double) * 4 * nwin, cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    for (int w = 0; w < nwin; ++w) { means[w] = h[4 * w + 1]; stds[w] = h[4 * w + 3]; }
    c.last_launches = s.launches;
    return EB2_OK;
  });
}

// Many tasks of one shape in one call: no interpreter work and no GIL between tasks.  Task t uses
// cols[t * d .. t * d + d).  status[t] = 0, or EB2_ERR_* | data_flags << 8 with values[t] = NaN; the
// call itself fails only on bad arguments.
int eb2_mi_cols_batch(int dev, const eb2_col_t* cols, int64_t ntasks, int c_dim, int64_t n, int k, uint32_t flags,
                      double* values, int* status) {
  if (!cols || !values || !status || ntasks < 0 || c_dim < 0) return fail(EB2_ERR_ARG, "eb2_mi_cols_batch: bad argument");
  const int d = 2 + c_dim;
  for (int64_t t = 0; t < ntasks; ++t) {
    g_data_flags = 0;
    g_skip_timing = t + 1 < ntasks;      // eb2_last_timing reports the last task of the batch
    const int rc = c_dim == 0 ? eb2_ksg_mi_cols(dev, cols + t * d, n, k, flags, values + t)
                              : eb2_cmi_cols(dev, cols + t * d, n, c_dim, k, flags, values + t);
    status[t] = rc ? (rc | (g_data_flags << 8)) : 0;
    if (rc) values[t] = std::numeric_limits<double>::quiet_NaN();
  }
  g_skip_timing = false;
  return EB2_OK;
}

// Two ways to order the slots of every chunk by the in-chunk coordinate: one CTA-wide bitonic sort per chunk
// (cell_sort_kernel: 2 launches, ~40 us of a nearly empty GPU at N = 10^5) or a stable partition of the globally
// sorted in-chunk coordinate by chunk (7 small launches, cheaper on the device from a few 10^5 rows on).  Small
// estimates are bound by the host's launch rate, large ones by device time.
int64_t partition_min_rows() {
  if (const char* e = getenv("EB2_PARTITION_MIN")) return atoll(e);     // tuning knob
  return 300000;
}

// ---- the bivariate pipeline (eb2_ksg2.h) for ONE estimate --------------------------------------------------------
// descriptors travel as kernel parameters (not through a staging copy) so that a captured graph carries them
__global__ void k2_setup_kernel(const k2::Col c0, const k2::Col c1, const k2::Prob p, k2::Col* cols, k2::Prob* probs) {
  cols[0] = c0;
  cols[1] = c1;
  probs[0] = p;
  *c0.flag = 0;
  *c1.flag = 0;
}

int64_t k2_min_rows() {
  if (const char* e = getenv("EB2_K2_MIN")) return atoll(e);     // tuning / test knob
  return 2048;
}

// raw = [x ; y] on the device.  Everything up to the result copy is enqueued on the lane's stream.
double* run_k2(Scratch& s, const k2::Plan& plan, const double* raw, int64_t n, int k, int64_t row_lo, int64_t row_hi,
               double* eps_out, int64_t* nx_out, int64_t* ny_out) {
  Ctx& c = s.c;
  cudaStream_t st = c.stream;
  const int k1t = (k + 1 <= 4) ? 4 : 8;
  ensure_psi_tab(s);
  k2::Col hc0 = k2::carve_col(s.dev<char>(k2::col_bytes(n)), n, raw);
  k2::Col hc1 = k2::carve_col(s.dev<char>(k2::col_bytes(n)), n, raw + n);
  k2::Prob hp = k2::carve_prob(s.dev<char>(k2::prob_bytes(plan, k1t)), plan, k1t, 0, 1);
  hp.out = s.result;
  if (eps_out) { hp.eps_row = s.dev<double>(n); CU(cudaMemsetAsync(hp.eps_row, 0xFF, sizeof(double) * n, st)); }
  if (nx_out) { hp.nx_row = s.dev<long long>(n); CU(cudaMemsetAsync(hp.nx_row, 0xFF, sizeof(long long) * n, st)); }
  if (ny_out) { hp.ny_row = s.dev<long long>(n); CU(cudaMemsetAsync(hp.ny_row, 0xFF, sizeof(long long) * n, st)); }
  k2::Col* dcols = s.dev<k2::Col>(2);
  k2::Prob* dprob = s.dev<k2::Prob>(1);
  k2_setup_kernel<<<1, 1, 0, st>>>(hc0, hc1, hp, dcols, dprob);
  s.launches++;
  const k2::Shard sh{row_lo, row_hi};
  // the search needs the bucket structure of x only: the y column's grid and the fine cells of both columns (what the
  // marginal counts read) are built on the second stream while layout and search run
  // (the x buckets get the whole GPU first: they are on the critical path)
  s.used_side = true;
  CU(k2::colgrid_buckets(dcols, 1, plan, st, &s.launches));
  CU(cudaEventRecord(c.fork, st));
  CU(cudaStreamWaitEvent(c.side, c.fork, 0));
  CU(k2::colgrid_buckets(dcols + 1, 1, plan, c.side, &s.launches));
  CU(k2::colgrid_cells(dcols, 2, plan, c.side, &s.launches));
  CU(cudaEventRecord(c.join, c.side));
  CU(k2::layout(dcols, dprob, 1, plan, st, &s.launches));
  mark(s, 1);
  CU(k2::knn(dcols, dprob, 1, plan, k, sh, c.sm_count, st, &s.launches));
  mark(s, 2);
  side_join(s);
  CU(k2::count_psi(dcols, dprob, 1, plan, sh, c.psi_tab, kPsiTab, st, &s.launches));
  mark(s, 3);
  CU(k2::finalize(dcols, dprob, 1, plan, st, &s.launches));
  mark(s, 4);
  if (eps_out) CU(cudaMemcpyAsync(eps_out, hp.eps_row, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (nx_out) CU(cudaMemcpyAsync(nx_out, hp.nx_row, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  if (ny_out) CU(cudaMemcpyAsync(ny_out, hp.ny_row, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  return s.result;
}

// ---- three-level grid (eb2_ksg2.h) for one estimate in a space of three and more dimensions ----------------------------
__global__ void g3_setup_kernel(const k2::Col c, const k2::Grid3 g, k2::Col* dc, k2::Grid3* dg) {
  *dc = c;
  *dg = g;
  *c.flag = 0;
}
__global__ void widen_kernel(const int* src, long long* dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// Smallest input the three-level grid takes in a space of `dims` dimensions.  Measured on a B200 (tools/g3_time.py,
// k-NN entropy, grid vs general path): N = 5*10^5: 3-D 1.7 vs 2.3 ms, 4-D 7.8 vs 19.7 ms, 5-D 45 vs 29 ms; 4-D at
// N = 2*10^6: 25 vs 67 ms, at N = 10^5: 0.88 vs 0.67 ms.  The grid orders three coordinates; from the fifth dimension on
// too much of a window is rows that only differ in the coordinates it does not order.
int64_t g3_min_rows(int dims) {
  if (const char* e = getenv("EB2_G3_MIN")) return atoll(e);     // tuning / test knob (every dimension)
  return dims <= 4 ? 200000 : INT64_MAX;
}

// tiles over the rows [0, n) in the caller's order (the grid kernels leave their per-query results by ROW, so that the
// digamma / log reduction runs in a fixed order whatever order the grid put the rows in)
TileSet row_tiles(Scratch& s, int64_t n) {
  PointSet ps;
  ps.qpt = kMaxQpt; ps.d = 1; ps.n = n; ps.stride = n; ps.sort_row = -1;
  ps.seg_slot = {0};
  ps.seg_len = {static_cast<int>(n)};
  return make_tiles(s, ps, 0, n, true, 0, 0);
}

struct G3Run {
  k2::Grid3 host;        // device pointers of the grid
  TileSet rows;
};

// rows_src[d]: coordinate d of the space, d < D, on the device; coordinates 0 .. G-1 carry the grid
G3Run run_g3(Scratch& s, const k2::Plan& plan, const double* const* rows_src, int D, int G, int64_t n, int k, int count_c,
             const CallInit& ci) {
  Ctx& c = s.c;
  cudaStream_t st = c.stream;
  c.last_pipeline = 2;
  const int k1t = (k + 1 <= 4) ? 4 : 8;
  k2::Col hc = k2::carve_col(s.dev<char>(k2::col_bytes(n)), n, rows_src[0]);
  k2::Grid3 hg = k2::carve_grid3(s.dev<char>(k2::grid3_bytes(plan, D, k1t)), plan, D, G, k1t);
  for (int d = 0; d < D; ++d) hg.raw[d] = rows_src[d];
  hg.pairs = ci.pairs;
  // (a call that ends in a bucket overflow still runs its reduction kernels: they must not see uninitialised counts)
  CU(cudaMemsetAsync(hg.eps_row, 0, sizeof(double) * n, st));
  if (count_c > 0)
    for (int i = 0; i < 3; ++i) CU(cudaMemsetAsync(hg.cnt_row[i], 0, sizeof(int) * n, st));
  k2::Col* dcol = s.dev<k2::Col>(1);
  k2::Grid3* dg = s.dev<k2::Grid3>(1);
  g3_setup_kernel<<<1, 1, 0, st>>>(hc, hg, dcol, dg);
  s.launches++;
  CU(k2::colgrid_buckets(dcol, 1, plan, st, &s.launches));
  CU(k2::layout3(dcol, dg, plan, st, &s.launches));
  mark(s, 1);
  CU(k2::knn3(dcol, dg, hg, plan, k, c.sm_count, st, &s.launches));
  mark(s, 2);
  if (count_c > 0) CU(k2::count3(dcol, dg, hg, plan, count_c, st, &s.launches));
  or_flags_kernel<<<1, 1, 0, st>>>(ci.nonfinite, hc.flag, nullptr);          // a bucket overflow sends the call to the general path
  s.launches++;
  G3Run r;
  r.host = hg;
  r.rows = row_tiles(s, n);
  return r;
}

// ---- a1: KSG ------------------------------------------------------------------------------------
static int ksg_rows_once(int dev, const Input& in, int64_t n, int k, int64_t row_lo, int64_t row_hi,
                         double* partial, double* eps_out, int64_t* nx_out, int64_t* ny_out, bool general_only) {
  const uint32_t flags = in.flags;
  return guarded(dev, [&](Ctx& c) {
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    // the bivariate pipeline (sample-sorted columns, bucket layout, one warp per 32 queries) takes every pruned estimate
    // of a size it is built for; a bucket that outgrows its CTA (heavily tied data) sends the call back here with
    // general_only set
    bool plan_ok = false;
    const k2::Plan plan = k2::make_plan(n, &plan_ok);
    const bool use_k2 = plan_ok && !general_only && prune && !(flags & EB2_FLAG_BRUTE_COUNT) && k + 1 <= 8 &&
                        n >= k2_min_rows() && !getenv("EB2_NO_K2");
    // Repeated resident estimates of one shape (unless EB2_GRAPH=0): the second call runs eagerly once the lane workspace has
    // its final size, the third is captured into a CUDA graph, later ones replay it with one launch.
    const bool graphable = graph_enabled() && !in.cols && (flags & EB2_FLAG_DEVICE_INPUT) && prune &&
                           !(flags & EB2_FLAG_BRUTE_COUNT) &&
                           (use_k2 || (n >= partition_min_rows() && !getenv("EB2_NO_CELLS") && !getenv("EB2_CELL_SORT"))) &&
                           !eps_out && !nx_out && !ny_out && partial && !g_skip_timing;
    const int64_t gkey[10] = {static_cast<int64_t>(reinterpret_cast<intptr_t>(in.coords)), n, k,
                              static_cast<int64_t>(flags) | (use_k2 ? int64_t(1) << 40 : 0), row_lo, row_hi,
                              static_cast<int64_t>(reinterpret_cast<intptr_t>(c.arena)), static_cast<int64_t>(c.arena_cap),
                              0, 0};
    if (graphable) {
      Ctx::GraphEntry* ge = graph_entry(c, gkey);
      if (ge->exec) return graph_replay(c, *ge, partial);
    }
    auto run = [&](bool capture) -> int {
    Scratch s(c);
    if (capture) {
      CU(cudaSetDevice(c.dev));
      CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
      s.capture = graph_entry(c, gkey);
    }
    CallInit ci = begin_call(s);
    const double* raw = nullptr;
    c.last_pipeline = use_k2 ? 1 : 0;
    if (use_k2) {
      raw = stage_input(s, in, 2, n, ci.nonfinite, false);        // (the bucket count kernel checks for non-finite values)
      double* out4 = run_k2(s, plan, raw, n, k, row_lo, row_hi, eps_out, nx_out, ny_out);
      const int rc = finish_call(s, out4, ci.pairs, ci.nonfinite, -1, partial);
      if (graphable && !capture && rc == EB2_OK && s.overflow == 0 && reinterpret_cast<intptr_t>(c.arena) == static_cast<intptr_t>(gkey[6]))
        graph_entry(c, gkey)->seen = true;
      return rc;
    }
    const Derived* dx = nullptr;
    const Derived* dy = nullptr;
    const double* ys_direct = nullptr;
    PointSet ps;
    if (in.cols && prune && !(flags & (EB2_FLAG_BRUTE_COUNT | EB2_FLAG_SINGLE_USE | EB2_FLAG_DEVICE_STATS)) && !getenv("EB2_NO_DERIVED")) {
      // prepared variables (rescaled values + their ascending order) are shared by all tasks of the call
      dx = get_derived(s, in.cols[0], n);
      dy = get_derived(s, in.cols[1], n);
      or_flags_kernel<<<1, 1, 0, c.stream>>>(ci.nonfinite, dx->dflag, dy->dflag);
      s.launches++;
      const double* rows[2] = {dx->vals, dy->vals};
      const Presorted pre{dx->sorted, dx->perm};
      const Presorted pre2{dy->sorted, dy->perm};
      ps = build_point_set_rows(s, rows, 2, n, nullptr, {}, 0, 2, 1, &pre, n >= partition_min_rows() ? &pre2 : nullptr);
    } else {
      raw = stage_input(s, in, 2, n, ci.nonfinite);
      if (prune && n >= partition_min_rows() && !(flags & EB2_FLAG_BRUTE_COUNT) && !getenv("EB2_NO_CELLS") && !getenv("EB2_CELL_SORT")) {
        // x and y are sorted side by side on the two streams (value -> row); the two-level layout then is a stable
        // partition of the y order by x chunk, and both ascending arrays serve the marginal searches afterwards
        int* iota = s.dev<int>(n);
        iota_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(iota, static_cast<int>(n));
        s.launches++;
        double* xk = s.dev<double>(n); int* xp = s.dev<int>(n);
        double* yk = s.dev<double>(n); int* yp = s.dev<int>(n);
        sort_pairs_side<double, int>(s, raw + n, yk, iota, yp, static_cast<int>(n), 0, 64);
        sort_pairs<double, int>(s, raw, xk, iota, xp, static_cast<int>(n), 0, 64);
        const double* rows[2] = {raw, raw + n};
        const Presorted pre{xk, xp};
        const Presorted pre2{yk, yp};
        ps = build_point_set_rows(s, rows, 2, n, nullptr, {}, 0, 2, 1, &pre, &pre2, true);
        ys_direct = yk;
      } else {
        ps = build_point_set(s, raw, 2, n, nullptr, {}, prune ? 0 : -1, 2, 1);   // x across chunks, y inside
      }
    }
    TileSet self = make_tiles(s, ps, row_lo, row_hi, true, 0, 0);
    // the ascending y needed by the n_y search: a by-product of the layout above, or (layouts without it) sorted on
    // the second stream while the k-NN kernel runs
    const double* ys = dy ? dy->sorted : ys_direct;
    bool forked = false;
    if (!ys && !(flags & EB2_FLAG_BRUTE_COUNT)) {
      double* t = s.dev<double>(n);
      size_t tmp_bytes = 0;
      CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, raw + n, t, (int)n, 0, 64, c.side));
      void* tmp = s.dev<char>(tmp_bytes);
      s.used_side = true;
      CU(cudaEventRecord(c.fork, c.stream));
      CU(cudaStreamWaitEvent(c.side, c.fork, 0));
      CU(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, raw + n, t, (int)n, 0, 64, c.side));
      CU(cudaEventRecord(c.join, c.side));
      ys = t;
      forked = true;
    }
    mark(s, 1);
    double* eps = s.dev<double>(ps.stride);
    double* radius = s.dev<double>(ps.stride);
    run_knn(s, ps, rows_range(0, 2), 2, k, self, eps, ci.pairs);
    mark(s, 2);
    if (flags & EB2_FLAG_BRUTE_COUNT) {      // (the sort + search path takes the radius from eps itself)
      radius_kernel<<<cdiv(ps.stride, 256), 256, 0, c.stream>>>(eps, radius, ps.stride);
      s.launches++;
    }
    if (forked) CU(cudaStreamWaitEvent(c.stream, c.join, 0));
    int* nx = s.dev<int>(ps.stride);
    int* ny = s.dev<int>(ps.stride);
    if (flags & EB2_FLAG_BRUTE_COUNT) {
      CountOut out; out.e0 = nx; out.e1 = ny;
      run_count(s, ps, ps, 0, 2, RowSel{}, RowSel{}, rows_range(0, 2), rows_range(0, 2), radius, self, false, out, ci.pairs);
    } else {
      const double* xs = ps.sorted_keys;             // x in ascending order, by-product of the layout sort
      if (!prune) { double* t = s.dev<double>(n); sort_keys(s, raw, t, (int)n); xs = t; }
      TileSet all = make_tiles(s, ps, row_lo, row_hi, false, 0, (int)n);
      run_search(s, ps.P, eps, xs, all, nx, ps.P + ps.stride, ys, ny, true);
    }
    mark(s, 3);
    double* out4 = run_psi(s, PSI_AB, nx, ny, nullptr, nullptr, self);
    mark(s, 4);
    const int* cnts[3] = {nx, ny, nullptr};
    Outputs o; o.eps = eps_out; o.cnt[0] = nx_out; o.cnt[1] = ny_out;
    export_outputs(s, ps, eps, cnts, o);
    const int rc = finish_call(s, out4, ci.pairs, ci.nonfinite, self.rows, partial);
    if (graphable && !capture && rc == EB2_OK && s.overflow == 0 && reinterpret_cast<intptr_t>(c.arena) == static_cast<intptr_t>(gkey[6]))
      graph_entry(c, gkey)->seen = true;       // (looked up again: tile-table eviction may have emptied the cache)
    return rc;
    };
    if (graphable) {
      Ctx::GraphEntry* ge = graph_entry(c, gkey);
      if (ge->seen && !ge->failed) {
        try {
          return run(true);
        } catch (const CudaFail&) {         // capture is best effort: stay eager for this signature
          abort_capture(c);
          Ctx::GraphEntry* g2 = graph_entry(c, gkey);
          if (g2->exec) { cudaGraphExecDestroy(g2->exec); g2->exec = nullptr; }
          g2->failed = true;
        } catch (...) {
          abort_capture(c);
          throw;
        }
      }
    }
    return run(false);
  });
}

static int ksg_rows_impl(int dev, const Input& in, int64_t n, int k, int64_t row_lo, int64_t row_hi,
                         double* partial, double* eps_out, int64_t* nx_out, int64_t* ny_out) {
  const int rc = ksg_rows_once(dev, in, n, k, row_lo, row_hi, partial, eps_out, nx_out, ny_out, false);
  if (rc != EB2_ERR_RETRY_GENERAL) return rc;
  return ksg_rows_once(dev, in, n, k, row_lo, row_hi, partial, eps_out, nx_out, ny_out, true);
}

int eb2_ksg_mi_rows(int dev, const double* coords, int64_t n, int k, uint32_t flags, int64_t row_lo, int64_t row_hi,
                    double* partial, double* eps_out, int64_t* nx_out, int64_t* ny_out) {
  if (int rc0 = check_common(coords, n, 2, k)) return rc0;
  if (!partial) return fail(EB2_ERR_ARG, "partial is NULL");
  Input in; in.coords = coords; in.flags = flags;
  return ksg_rows_impl(dev, in, n, k, row_lo, row_hi, partial, eps_out, nx_out, ny_out);
}

int eb2_ksg_mi_cols_rows(int dev, const eb2_col_t* cols, int64_t n, int k, uint32_t flags, int64_t row_lo,
                         int64_t row_hi, double* partial) {
  if (!cols || !partial) return fail(EB2_ERR_ARG, "cols/partial is NULL");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, 2, k)) return rc0;
  Input in; in.cols = cols; in.flags = flags & ~EB2_FLAG_DEVICE_INPUT;
  return ksg_rows_impl(dev, in, n, k, row_lo, row_hi, partial, nullptr, nullptr, nullptr);
}

int eb2_ksg_mi_cols(int dev, const eb2_col_t* cols, int64_t n, int k, uint32_t flags, double* value) {
  if (!cols || !value) return fail(EB2_ERR_ARG, "cols/value is NULL");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, 2, k)) return rc0;
  Input in; in.cols = cols; in.flags = flags & ~EB2_FLAG_DEVICE_INPUT;
  double partial[EB2_P_LEN];
  const int rc = ksg_rows_impl(dev, in, n, k, 0, n, partial, nullptr, nullptr, nullptr);
  if (rc) return rc;
  return eb2_ksg_mi_finish(partial, n, k, value);
}

int eb2_ksg_mi_finish(const double* partial, int64_t n, int k, double* value) {
  if (!partial || !value || n <= 0 || k <= 0) return fail(EB2_ERR_ARG, "bad argument");
  *value = psi_host((double)n) + psi_host((double)k) - psi_mean(partial, n);    // :113
  return EB2_OK;
}

int eb2_ksg_mi(int dev, const double* coords, int64_t n, int k, uint32_t flags, double* value, double* eps_out,
               int64_t* nx_out, int64_t* ny_out) {
  if (!value) return fail(EB2_ERR_ARG, "value is NULL");
  double partial[EB2_P_LEN];
  const int rc = eb2_ksg_mi_rows(dev, coords, n, k, flags, 0, n, partial, eps_out, nx_out, ny_out);
  if (rc) return rc;
  return eb2_ksg_mi_finish(partial, n, k, value);
}


// ---- configs[1] from ONE process: the query rows of a single estimate sharded over the GPUs of the box -----------------
// (the multi-process form of the same - one rank per GPU, torch.distributed, NCCL all-reduce of the partial blocks - is
// ennemi_b200/distributed.py).  Every GPU receives the point set, builds its grid, searches and counts the buckets of its
// shard; the partial blocks are summed here on the host.  Their digamma sums are exact integers (EB2_P_FIX0), so the
// value is bit-identical for every ngpu.
int eb2_sharded_ksg_mi(int ngpu, const double* coords, int64_t n, int k, uint32_t flags, double* value) {
  if (int rc0 = check_common(coords, n, 2, k)) return rc0;
  if (!value) return fail(EB2_ERR_ARG, "value is NULL");
  if (flags & EB2_FLAG_DEVICE_INPUT) return fail(EB2_ERR_ARG, "eb2_sharded_ksg_mi takes host buffers only");
  const int avail = eb2_device_count();
  if (ngpu < 1 || ngpu > avail) return fail(EB2_ERR_ARG, "ngpu = %d, but %d CUDA device(s) are usable", ngpu, avail);
  std::vector<std::vector<double>> parts(ngpu, std::vector<double>(EB2_P_LEN, 0.0));
  std::vector<int> rcs(ngpu, 0);
  std::vector<std::string> errs(ngpu);
  std::vector<int> dflags(ngpu, 0);
  auto work = [&](int g) {
    const int64_t lo = n * g / ngpu, hi = n * (g + 1) / ngpu;
    rcs[g] = eb2_ksg_mi_rows(g, coords, n, k, flags, lo, hi, parts[g].data(), nullptr, nullptr, nullptr);
    if (rcs[g]) { errs[g] = g_err; dflags[g] = g_data_flags; }       // (both are thread-local)
  };
  std::vector<std::thread> threads;
  for (int g = 1; g < ngpu; ++g) threads.emplace_back(work, g);
  work(0);
  for (auto& t : threads) t.join();
  for (int g = 0; g < ngpu; ++g)
    if (rcs[g]) { g_err = errs[g]; g_data_flags = dflags[g]; return rcs[g]; }
  std::vector<double> total(EB2_P_LEN, 0.0);
  for (int g = 0; g < ngpu; ++g)
    for (int i = 0; i < EB2_P_LEN; ++i) total[i] += parts[g][i];
  return eb2_ksg_mi_finish(total.data(), n, k, value);
}

// ---- a7 for pairwise_mi: all pairs of a set of prepared variables in one call -------------------------------------
// (the reference builds the same task list at ennemi/_driver.py:703-707 and maps it over a thread pool, :736-785).
// Every variable is rescaled and sample-sorted ONCE; the pairs then go through the bivariate pipeline in batches,
// each stage one launch for the whole batch (blockIdx.y / .z = pair).
int eb2_ksg_mi_pairs(int dev, const eb2_col_t* cols, int nvar, const int32_t* pairs, int64_t npairs, int64_t n, int k,
                     uint32_t flags, double* values, int* status) {
  if (!cols || !pairs || !values || !status || nvar < 1 || npairs < 0) return fail(EB2_ERR_ARG, "eb2_ksg_mi_pairs: bad argument");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, 2, k)) return rc0;
  for (int64_t t = 0; t < 2 * npairs; ++t)
    if (pairs[t] < 0 || pairs[t] >= nvar) return fail(EB2_ERR_ARG, "eb2_ksg_mi_pairs: variable index out of range");
  bool plan_ok = false;
  const k2::Plan plan = k2::make_plan(n, &plan_ok);
  auto one_by_one = [&](int64_t t, bool general_only) {
    const eb2_col_t two[2] = {cols[pairs[2 * t]], cols[pairs[2 * t + 1]]};
    Input in; in.cols = two; in.flags = (flags & ~EB2_FLAG_DEVICE_INPUT) | EB2_FLAG_SINGLE_USE;
    double partial[EB2_P_LEN];
    g_data_flags = 0;
    int rc = ksg_rows_once(dev, in, n, k, 0, n, partial, nullptr, nullptr, nullptr, general_only);
    if (rc == EB2_ERR_RETRY_GENERAL) rc = ksg_rows_once(dev, in, n, k, 0, n, partial, nullptr, nullptr, nullptr, true);
    status[t] = rc ? (rc | (g_data_flags << 8)) : 0;
    values[t] = std::numeric_limits<double>::quiet_NaN();
    if (!rc) eb2_ksg_mi_finish(partial, n, k, values + t);
  };
  if (!plan_ok || k + 1 > 8 || n < k2_min_rows() || getenv("EB2_NO_K2") || npairs == 0) {
    for (int64_t t = 0; t < npairs; ++t) one_by_one(t, false);
    return EB2_OK;
  }
  std::vector<Res> res(static_cast<size_t>(npairs));
  const int rc = guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    record_event(s, c.ev[0]);
    const int k1t = (k + 1 <= 4) ? 4 : 8;
    ensure_psi_tab(s);
    // 1. prepared variables (the reference's _rescale_data, ennemi/_driver.py:871-902) and their sample sorts
    double* vals = s.dev<double>(static_cast<size_t>(nvar) * n);
    int* colflags = s.dev<int>(nvar);
    CU(cudaMemsetAsync(colflags, 0, sizeof(int) * nvar, st));
    const size_t cbytes = k2::col_bytes(n);
    char* cscratch = s.dev<char>(cbytes * nvar);
    k2::Col* hcols = s.host<k2::Col>(nvar);
    {
      std::lock_guard<std::mutex> cache_guard(c.shared->mu);
      auto& cache = c.shared->cache;
      for (int v0 = 0; v0 < nvar; v0 += kMaxDimAny) {
        const int m = std::min(kMaxDimAny, nvar - v0);
        PrepArgs pa;
        pa.d = m; pa.n = n; pa.raw = nullptr; pa.flags = nullptr;
        for (int t = 0; t < m; ++t) {
          const eb2_col_t& col = cols[v0 + t];
          auto it = cache.find(col.key);
          if (it == cache.end()) throw CudaFail{cudaErrorInvalidValue, "column key not in the device cache", __LINE__};
          const int64_t last = col.off + (n - 1) * col.stride;
          if (col.off < 0 || last < 0 || col.off >= it->second.second || last >= it->second.second)
            throw CudaFail{cudaErrorInvalidValue, "column slice outside the cached column", __LINE__};
          PrepCol& pc = pa.col[t];
          pc.src = it->second.first; pc.off = col.off; pc.stride = col.stride; pc.mean = col.mean; pc.std = col.std;
          pc.noise = nullptr; pc.noff = col.noff; pc.nstride = col.nstride; pc.dstats = nullptr;
          if (col.nkey != 0) {
            auto nt = cache.find(col.nkey);
            if (nt == cache.end()) throw CudaFail{cudaErrorInvalidValue, "noise key not in the device cache", __LINE__};
            const int64_t nlast = col.noff + (n - 1) * col.nstride;
            if (col.noff < 0 || nlast < 0 || nlast >= nt->second.second)
              throw CudaFail{cudaErrorInvalidValue, "noise slice outside the cached vector", __LINE__};
            pc.noise = nt->second.first;
          }
          pc.flag = colflags + v0 + t;
          pc.dst = vals + static_cast<int64_t>(v0 + t) * n;
          hcols[v0 + t] = k2::carve_col(cscratch + cbytes * (v0 + t), n, pc.dst);
          hcols[v0 + t].flag = pc.flag;
        }
        prep_kernel<<<cdiv(static_cast<int64_t>(m) * n, 256), 256, 0, st>>>(pa);
        s.launches++;
      }
    }
    k2::Col* dcols = s.dev<k2::Col>(nvar);
    CU(cudaMemcpyAsync(dcols, hcols, sizeof(k2::Col) * nvar, cudaMemcpyHostToDevice, st));
    CU(k2::colgrid(dcols, nvar, plan, st, &s.launches));
    record_event(s, c.ev[1]);
    // 2. the pairs, in batches sized to a workspace budget
    const size_t pbytes = k2::prob_bytes(plan, k1t);
    size_t budget = size_t(3) << 30;
    if (const char* e = getenv("EB2_PAIR_BATCH_MB")) budget = size_t(atoll(e)) << 20;     // tuning knob
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(npairs, static_cast<int64_t>(budget / pbytes)));
    char* pscratch = s.dev<char>(pbytes * batch);
    double* outs = s.dev<double>(static_cast<size_t>(npairs) * kResDoubles);
    CU(cudaMemsetAsync(outs, 0, sizeof(double) * kResDoubles * npairs, st));
    k2::Prob* dprobs = s.dev<k2::Prob>(batch);
    const k2::Shard whole{0, n};
    for (int64_t p0 = 0; p0 < npairs; p0 += batch) {
      const int m = static_cast<int>(std::min<int64_t>(batch, npairs - p0));
      k2::Prob* hp = s.host<k2::Prob>(m);
      for (int t = 0; t < m; ++t) {
        hp[t] = k2::carve_prob(pscratch + pbytes * t, plan, k1t, pairs[2 * (p0 + t)], pairs[2 * (p0 + t) + 1]);
        hp[t].out = outs + kResDoubles * (p0 + t);
      }
      CU(cudaMemcpyAsync(dprobs, hp, sizeof(k2::Prob) * m, cudaMemcpyHostToDevice, st));
      CU(k2::layout(dcols, dprobs, m, plan, st, &s.launches));
      CU(k2::knn(dcols, dprobs, m, plan, k, whole, c.sm_count, st, &s.launches));
      CU(k2::count_psi(dcols, dprobs, m, plan, whole, c.psi_tab, kPsiTab, st, &s.launches));
      CU(k2::finalize(dcols, dprobs, m, plan, st, &s.launches));
    }
    record_event(s, c.ev[2]);
    record_event(s, c.ev[3]);
    record_event(s, c.ev[4]);
    static_assert(sizeof(Res) <= kResDoubles * sizeof(double), "result block");
    std::vector<double> hout(static_cast<size_t>(npairs) * kResDoubles);
    CU(cudaMemcpyAsync(hout.data(), outs, sizeof(double) * kResDoubles * npairs, cudaMemcpyDeviceToHost, st));
    record_event(s, c.ev[5]);
    CU(cudaStreamSynchronize(st));
    read_phase_times(c);
    c.last_launches = s.launches;
    for (int64_t t = 0; t < npairs; ++t) std::memcpy(&res[t], hout.data() + kResDoubles * t, sizeof(Res));
    return EB2_OK;
  });
  if (rc) return rc;
  for (int64_t t = 0; t < npairs; ++t) {
    const Res& h = res[t];
    values[t] = std::numeric_limits<double>::quiet_NaN();
    if (h.nonfinite & k2::kFlagOverflow) {
      one_by_one(t, true);                 // heavily tied data: the general path takes this pair
    } else if (h.nonfinite) {
      status[t] = EB2_ERR_NONFINITE | ((h.nonfinite & 0xff) << 8);
    } else {
      status[t] = 0;
      double partial[EB2_P_LEN];
      parse_result(&h, -1, partial);
      eb2_ksg_mi_finish(partial, n, k, values + t);
    }
  }
  return EB2_OK;
}

// ---- a2: Frenzel-Pompe ----------------------------------------------------------------------------
static int cmi_rows_once(int dev, const Input& in, int64_t n, int c_dim, int k, int64_t row_lo, int64_t row_hi,
                         double* partial, double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out, bool general_only) {
  const uint32_t flags = in.flags;
  const int d = 2 + c_dim;
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    const double* raw = stage_input(s, in, d, n, ci.nonfinite);
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    bool plan_ok = false;
    const k2::Plan plan = k2::make_plan3(n, d, &plan_ok);
    if (plan_ok && !general_only && prune && c_dim >= 2 && d <= k2::kG3MaxD && k + 1 <= 8 && row_lo == 0 && row_hi == n &&
        n >= g3_min_rows(d) && !getenv("EB2_NO_G3") && getenv("EB2_G3_CMI")) {
      // (opt-in: measured at N = 2*10^5, c = 3 the grid examines 6 x fewer candidates than the general path but walks them
      //  one thread per query and comes out slower, 12.5 against 5.7 ms; the counts tie.  Kept for the parity tests.)
      // three-level grid over the condition: every space on this path (xyz, xz, yz, z) contains z, so ONE grid on
      // (z_0, z_1[, z_2]) serves the search in the joint space and the three counts; x and y travel as payload
      const double* rows[k2::kG3MaxD];
      for (int t = 0; t < c_dim; ++t) rows[t] = raw + static_cast<int64_t>(2 + t) * n;
      rows[c_dim] = raw;
      rows[c_dim + 1] = raw + n;
      G3Run g = run_g3(s, plan, rows, d, c_dim >= 3 ? 3 : 2, n, k, c_dim, ci);
      mark(s, 3);
      double* out4 = run_psi(s, PSI_AB_MINUS_C, g.host.cnt_row[1], g.host.cnt_row[2], g.host.cnt_row[0], nullptr, g.rows);
      mark(s, 4);
      if (eps_out) CU(cudaMemcpyAsync(eps_out, g.host.eps_row, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
      int64_t* outs[3] = {nz_out, nxz_out, nyz_out};
      for (int i = 0; i < 3; ++i) {
        if (!outs[i]) continue;
        long long* wide = s.dev<long long>(n);
        widen_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(g.host.cnt_row[i], wide, n);
        s.launches++;
        CU(cudaMemcpyAsync(outs[i], wide, sizeof(long long) * n, cudaMemcpyDeviceToHost, c.stream));
      }
      return finish_call(s, out4, ci.pairs, ci.nonfinite, n, partial);
    }
    // every space on this path (xyz, xz, yz, z) contains z_0: sort by it
    // ... and, with a condition of 2+ dimensions, order the slots inside each chunk by z_1 (two-level layout)
    c.last_pipeline = 0;
    PointSet ps = build_point_set(s, raw, d, n, nullptr, {}, prune ? 2 : -1, c_dim >= 2 ? d : 0, c_dim >= 2 ? 3 : -1);
    TileSet self = make_tiles(s, ps, row_lo, row_hi, true, 0, 0);
    mark(s, 1);
    double* eps = s.dev<double>(ps.stride);
    double* radius = s.dev<double>(ps.stride);
    run_knn(s, ps, rows_range(0, d), d, k, self, eps, ci.pairs);
    mark(s, 2);
    radius_kernel<<<cdiv(ps.stride, 256), 256, 0, c.stream>>>(eps, radius, ps.stride);
    s.launches++;
    int* nz = s.dev<int>(ps.stride);
    int* nxz = s.dev<int>(ps.stride);
    int* nyz = s.dev<int>(ps.stride);
    CountOut out; out.s = nz; out.e0 = nxz; out.e1 = nyz;
    run_count(s, ps, ps, c_dim, 2, rows_range(2, c_dim), rows_range(2, c_dim), rows_range(0, 2), rows_range(0, 2),
              radius, self, prune, out, ci.pairs);
    mark(s, 3);
    double* out4 = run_psi(s, PSI_AB_MINUS_C, nxz, nyz, nz, nullptr, self);
    mark(s, 4);
    const int* cnts[3] = {nxz, nyz, nz};
    Outputs o; o.eps = eps_out; o.cnt[0] = nxz_out; o.cnt[1] = nyz_out; o.cnt[2] = nz_out;
    export_outputs(s, ps, eps, cnts, o);
    return finish_call(s, out4, ci.pairs, ci.nonfinite, self.rows, partial);
  });
}

static int cmi_rows_impl(int dev, const Input& in, int64_t n, int c_dim, int k, int64_t row_lo, int64_t row_hi,
                         double* partial, double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out) {
  const int rc = cmi_rows_once(dev, in, n, c_dim, k, row_lo, row_hi, partial, eps_out, nxz_out, nyz_out, nz_out, false);
  if (rc != EB2_ERR_RETRY_GENERAL) return rc;
  return cmi_rows_once(dev, in, n, c_dim, k, row_lo, row_hi, partial, eps_out, nxz_out, nyz_out, nz_out, true);
}

int eb2_cmi_rows(int dev, const double* coords, int64_t n, int c_dim, int k, uint32_t flags, int64_t row_lo,
                 int64_t row_hi, double* partial, double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out) {
  if (c_dim < 1) return fail(EB2_ERR_ARG, "condition needs at least one dimension");
  if (int rc0 = check_common(coords, n, 2 + c_dim, k)) return rc0;
  if (!partial) return fail(EB2_ERR_ARG, "partial is NULL");
  Input in; in.coords = coords; in.flags = flags;
  return cmi_rows_impl(dev, in, n, c_dim, k, row_lo, row_hi, partial, eps_out, nxz_out, nyz_out, nz_out);
}

int eb2_cmi_cols_rows(int dev, const eb2_col_t* cols, int64_t n, int c_dim, int k, uint32_t flags, int64_t row_lo,
                      int64_t row_hi, double* partial) {
  if (!cols || !partial) return fail(EB2_ERR_ARG, "cols/partial is NULL");
  if (c_dim < 1) return fail(EB2_ERR_ARG, "condition needs at least one dimension");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, 2 + c_dim, k)) return rc0;
  Input in; in.cols = cols; in.flags = flags & ~EB2_FLAG_DEVICE_INPUT;
  return cmi_rows_impl(dev, in, n, c_dim, k, row_lo, row_hi, partial, nullptr, nullptr, nullptr, nullptr);
}

int eb2_cmi_cols(int dev, const eb2_col_t* cols, int64_t n, int c_dim, int k, uint32_t flags, double* value) {
  if (!cols || !value) return fail(EB2_ERR_ARG, "cols/value is NULL");
  if (c_dim < 1) return fail(EB2_ERR_ARG, "condition needs at least one dimension");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, 2 + c_dim, k)) return rc0;
  Input in; in.cols = cols; in.flags = flags & ~EB2_FLAG_DEVICE_INPUT;
  double partial[EB2_P_LEN];
  const int rc = cmi_rows_impl(dev, in, n, c_dim, k, 0, n, partial, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  return eb2_cmi_finish(partial, n, k, value);
}

int eb2_cmi_finish(const double* partial, int64_t n, int k, double* value) {
  if (!partial || !value || n <= 0 || k <= 0) return fail(EB2_ERR_ARG, "bad argument");
  *value = psi_host((double)k) - psi_mean(partial, n);    // :156
  return EB2_OK;
}

int eb2_cmi(int dev, const double* coords, int64_t n, int c_dim, int k, uint32_t flags, double* value, double* eps_out,
            int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out) {
  if (!value) return fail(EB2_ERR_ARG, "value is NULL");
  double partial[EB2_P_LEN];
  const int rc = eb2_cmi_rows(dev, coords, n, c_dim, k, flags, 0, n, partial, eps_out, nxz_out, nyz_out, nz_out);
  if (rc) return rc;
  return eb2_cmi_finish(partial, n, k, value);
}

// ---- a3: Ross -------------------------------------------------------------------------------------
int eb2_ross_mi(int dev, const double* coords, const int32_t* cls, int64_t n, int ncls, int k, uint32_t flags,
                double* value, double* eps_out, int64_t* nfull_out) {
  if (int rc0 = check_common(coords, n, 1, k)) return rc0;
  if (!cls || ncls < 1 || !value) return fail(EB2_ERR_ARG, "cls/ncls/value invalid");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    const double* raw = stage_coords(s, coords, 1, n, flags, ci.nonfinite);
    const int* cls_dev = nullptr;
    std::vector<int> sizes;
    if (!stage_classes(s, cls, n, ncls, flags, &cls_dev, &sizes)) return fail(EB2_ERR_ARG, "class id outside [0, ncls)");
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    PointSet ps = build_point_set(s, raw, 1, n, cls_dev, sizes, prune ? 0 : -1);
    TileSet self = make_tiles(s, ps, 0, n, true, 0, 0);
    mark(s, 1);
    double* eps = s.dev<double>(ps.stride);
    double* radius = s.dev<double>(ps.stride);
    run_knn(s, ps, rows_range(0, 1), 1, k, self, eps, ci.pairs);      // :194 within the class
    mark(s, 2);
    radius_kernel<<<cdiv(ps.stride, 256), 256, 0, c.stream>>>(eps, radius, ps.stride);
    s.launches++;
    int* nfull = s.dev<int>(ps.stride);
    if (flags & EB2_FLAG_BRUTE_COUNT) {
      TileSet all = make_tiles(s, ps, 0, n, false, 0, (int)ps.stride);
      CountOut out; out.e0 = nfull;
      run_count(s, ps, ps, 0, 1, RowSel{}, RowSel{}, rows_range(0, 1), rows_range(0, 1), radius, all, false, out, ci.pairs);
    } else {
      double* xs = s.dev<double>(n);
      sort_keys(s, raw, xs, (int)n);
      TileSet all = make_tiles(s, ps, 0, n, false, 0, (int)n);
      run_search(s, ps.P, radius, xs, all, nfull);                     // :196 over all x
    }
    mark(s, 3);
    double* out4 = run_psi(s, PSI_A, nfull, nullptr, nullptr, nullptr, self);
    mark(s, 4);
    const int* cnts[3] = {nfull, nullptr, nullptr};
    Outputs o; o.eps = eps_out; o.cnt[0] = nfull_out;
    export_outputs(s, ps, eps, cnts, o);
    double partial[EB2_P_LEN];
    const int rc = finish_call(s, out4, ci.pairs, ci.nonfinite, self.rows, partial);
    if (rc) return rc;
    double weighted = 0.0;                                             // :199
    for (int g = 0; g < ncls; ++g)
      if (sizes[g] > 0) weighted += psi_host((double)sizes[g]) * ((double)sizes[g] / (double)n);
    *value = psi_host((double)n) + psi_host((double)k) - psi_mean(partial, n) - weighted;   // :200
    return EB2_OK;
  });
}

// ---- a4: conditional Ross ---------------------------------------------------------------------------
int eb2_ross_cmi(int dev, const double* coords, const int32_t* cls, int64_t n, int c_dim, int ncls, int k,
                 uint32_t flags, double* value, double* eps_out, int64_t* nxz_out, int64_t* nyz_out, int64_t* nz_out) {
  if (c_dim < 1) return fail(EB2_ERR_ARG, "condition needs at least one dimension");
  const int d = 1 + c_dim;
  if (int rc0 = check_common(coords, n, d, k)) return rc0;
  if (!cls || ncls < 1 || !value) return fail(EB2_ERR_ARG, "cls/ncls/value invalid");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    const double* raw = stage_coords(s, coords, d, n, flags, ci.nonfinite);
    const int* cls_dev = nullptr;
    std::vector<int> sizes;
    if (!stage_classes(s, cls, n, ncls, flags, &cls_dev, &sizes)) return fail(EB2_ERR_ARG, "class id outside [0, ncls)");
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    // A: grouped by class (and ascending in z_0 inside a class); B: every row, ascending in z_0
    PointSet A = build_point_set(s, raw, d, n, cls_dev, sizes, prune ? 1 : -1);
    PointSet B = A;
    if (prune) B = build_point_set(s, raw, d, n, nullptr, {}, 1);
    TileSet self = make_tiles(s, A, 0, n, true, 0, 0);
    TileSet all = make_tiles(s, A, 0, n, false, 0, prune ? (int)n : (int)A.stride);
    mark(s, 1);
    double* eps = s.dev<double>(A.stride);
    double* radius = s.dev<double>(A.stride);
    run_knn(s, A, rows_range(0, d), d, k, self, eps, ci.pairs);        // :240 (x,z) within the class
    mark(s, 2);
    radius_kernel<<<cdiv(A.stride, 256), 256, 0, c.stream>>>(eps, radius, A.stride);
    s.launches++;
    int* nz = s.dev<int>(A.stride);
    int* nxz = s.dev<int>(A.stride);
    int* nyz = s.dev<int>(A.stride);
    CountOut o_all; o_all.s = nz; o_all.e0 = nxz;                      // :243, :245 over all rows
    run_count(s, A, B, c_dim, 1, rows_range(1, c_dim), rows_range(1, c_dim), rows_range(0, 1), rows_range(0, 1),
              radius, all, prune, o_all, ci.pairs);
    CountOut o_cls; o_cls.s = nyz;                                     // :244 z within the class
    run_count(s, A, A, c_dim, 0, rows_range(1, c_dim), rows_range(1, c_dim), RowSel{}, RowSel{}, radius, self, prune,
              o_cls, ci.pairs);
    mark(s, 3);
    double* out4 = run_psi(s, PSI_AB_MINUS_C, nxz, nyz, nz, nullptr, self);
    mark(s, 4);
    const int* cnts[3] = {nxz, nyz, nz};
    Outputs o; o.eps = eps_out; o.cnt[0] = nxz_out; o.cnt[1] = nyz_out; o.cnt[2] = nz_out;
    export_outputs(s, A, eps, cnts, o);
    double partial[EB2_P_LEN];
    const int rc = finish_call(s, out4, ci.pairs, ci.nonfinite, self.rows, partial);
    if (rc) return rc;
    *value = psi_host((double)k) - psi_mean(partial, n);               // :247
    return EB2_OK;
  });
}

// ---- a5: k-NN entropy ---------------------------------------------------------------------------------
static int entropy_rows_once(int dev, const Input& in, int64_t n, int m, int k, int64_t row_lo,
                             int64_t row_hi, double* partial, double* dist_out, bool general_only) {
  const uint32_t flags = in.flags;
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    const double* raw = stage_input(s, in, m, n, ci.nonfinite);
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    bool plan_ok = false;
    const k2::Plan plan = k2::make_plan3(n, m, &plan_ok);
    if (plan_ok && !general_only && prune && m >= 3 && m <= k2::kG3MaxD && k + 1 <= 8 && row_lo == 0 && row_hi == n &&
        n >= g3_min_rows(m) && !getenv("EB2_NO_G3")) {
      // three-level grid: buckets of coordinate 0, cells over coordinates 1 and 2
      const double* rows[k2::kG3MaxD];
      for (int t = 0; t < m; ++t) rows[t] = raw + static_cast<int64_t>(t) * n;
      G3Run g = run_g3(s, plan, rows, m, 3, n, k, 0, ci);
      mark(s, 3);
      double* out4 = run_psi(s, LOG_DIST, nullptr, nullptr, nullptr, g.host.eps_row, g.rows);     // :42, rows in the caller's order
      mark(s, 4);
      if (dist_out) CU(cudaMemcpyAsync(dist_out, g.host.eps_row, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
      return finish_call(s, out4, ci.pairs, ci.nonfinite, n, partial);
    }
    c.last_pipeline = 0;
    PointSet ps = build_point_set(s, raw, m, n, nullptr, {}, prune ? 0 : -1, m, m >= 2 ? 1 : -1);
    TileSet self = make_tiles(s, ps, row_lo, row_hi, true, 0, 0);
    mark(s, 1);
    double* dist = s.dev<double>(ps.stride);
    run_knn(s, ps, rows_range(0, m), m, k, self, dist, ci.pairs);      // :39
    mark(s, 2);
    mark(s, 3);
    double* out4 = run_psi(s, LOG_DIST, nullptr, nullptr, nullptr, dist, self);
    mark(s, 4);
    const int* cnts[3] = {nullptr, nullptr, nullptr};
    Outputs o; o.eps = dist_out;
    export_outputs(s, ps, dist, cnts, o);
    return finish_call(s, out4, ci.pairs, ci.nonfinite, self.rows, partial);
  });
}

int eb2_entropy_rows(int dev, const double* coords, int64_t n, int m, int k, uint32_t flags, int64_t row_lo,
                     int64_t row_hi, double* partial, double* dist_out) {
  if (int rc0 = check_common(coords, n, m, k)) return rc0;
  if (!partial) return fail(EB2_ERR_ARG, "partial is NULL");
  Input in; in.coords = coords; in.flags = flags;
  const int rc = entropy_rows_once(dev, in, n, m, k, row_lo, row_hi, partial, dist_out, false);
  if (rc != EB2_ERR_RETRY_GENERAL) return rc;
  return entropy_rows_once(dev, in, n, m, k, row_lo, row_hi, partial, dist_out, true);
}

// the same estimate on device-resident columns (H(X | C) = H(X, C) - H(C), _driver.py:202-220, uploads
// every column of X and C once and names them in both terms)
int eb2_entropy_cols(int dev, const eb2_col_t* cols, int64_t n, int m, int k, uint32_t flags, double* value) {
  if (!cols || !value) return fail(EB2_ERR_ARG, "cols/value is NULL");
  if (int rc0 = check_common(reinterpret_cast<const double*>(cols), n, m, k)) return rc0;
  Input in; in.cols = cols; in.flags = flags & ~EB2_FLAG_DEVICE_INPUT;
  double partial[EB2_P_LEN];
  int rc = entropy_rows_once(dev, in, n, m, k, 0, n, partial, nullptr, false);
  if (rc == EB2_ERR_RETRY_GENERAL) rc = entropy_rows_once(dev, in, n, m, k, 0, n, partial, nullptr, true);
  if (rc) return rc;
  return eb2_entropy_finish(partial, n, m, k, value);
}

int eb2_entropy_finish(const double* partial, int64_t n, int m, int k, double* value) {
  if (!partial || !value || n <= 0 || k <= 0 || m <= 0) return fail(EB2_ERR_ARG, "bad argument");
  const double mean_log = partial[EB2_P_SUM] / (double)n;
  *value = psi_host((double)n) - psi_host((double)k) + m * (mean_log + std::log(2.0));   // :42
  return EB2_OK;
}

int eb2_entropy(int dev, const double* coords, int64_t n, int m, int k, uint32_t flags, double* value, double* dist_out) {
  if (!value) return fail(EB2_ERR_ARG, "value is NULL");
  double partial[EB2_P_LEN];
  const int rc = eb2_entropy_rows(dev, coords, n, m, k, flags, 0, n, partial, dist_out);
  if (rc) return rc;
  return eb2_entropy_finish(partial, n, m, k, value);
}

// ---- a6: digamma on the device ---------------------------------------------------------------------------
int eb2_psi(int dev, const int64_t* counts, int64_t n, double* out) {
  if (!counts || !out || n <= 0) return fail(EB2_ERR_ARG, "bad argument");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    long long* dc = s.dev<long long>(n);
    double* dout = s.dev<double>(n);
    CU(cudaMemcpyAsync(dc, counts, sizeof(long long) * n, cudaMemcpyHostToDevice, c.stream));
    psi_array_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(dc, n, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    c.last_launches = 1;
    return EB2_OK;
  });
}

// ---- roofline denominator: measured FP64 (DADD) issue rate of this device -----------------------------
int eb2_measure_fp64_peak(int dev, double* tera_instr_per_s) {
  if (!tera_instr_per_s) return fail(EB2_ERR_ARG, "output is NULL");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CU(cudaSetDevice(c.dev));
    const int blocks = c.sm_count * 32, threads = 256, iters = 4096;
    double* out = s.dev<double>(static_cast<size_t>(blocks) * threads);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      CU(cudaEventRecord(c.ev[6], c.stream));
      fp64_peak_kernel<<<blocks, threads, 0, c.stream>>>(out, 1e-9, iters);
      CU(cudaEventRecord(c.ev[7], c.stream));
      CU(cudaStreamSynchronize(c.stream));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, c.ev[6], c.ev[7]));
      if (rep > 0 && ms < best) best = ms;
    }
    *tera_instr_per_s = 8.0 * iters * static_cast<double>(blocks) * threads / (best * 1e-3) * 1e-12;
    return EB2_OK;
  });
}

// ---- primitives ---------------------------------------------------------------------------------------------
int eb2_kth_distance(int dev, const double* coords, const int32_t* cls, int64_t n, int d, int ncls, int k,
                     uint32_t flags, double* out) {
  if (int rc0 = check_common(coords, n, d, k)) return rc0;
  if (!out) return fail(EB2_ERR_ARG, "out is NULL");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    const double* raw = stage_coords(s, coords, d, n, flags, ci.nonfinite);
    const int* cls_dev = nullptr;
    std::vector<int> sizes;
    if (cls && !stage_classes(s, cls, n, ncls, flags, &cls_dev, &sizes)) return fail(EB2_ERR_ARG, "class id outside [0, ncls)");
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    PointSet ps = build_point_set(s, raw, d, n, cls_dev, sizes, prune ? 0 : -1);
    TileSet self = make_tiles(s, ps, 0, n, true, 0, 0);
    mark(s, 1);
    double* eps = s.dev<double>(ps.stride);
    run_knn(s, ps, rows_range(0, d), d, k, self, eps, ci.pairs);
    mark(s, 2); mark(s, 3); mark(s, 4);
    const int* cnts[3] = {nullptr, nullptr, nullptr};
    Outputs o; o.eps = out;
    export_outputs(s, ps, eps, cnts, o);
    double* out4 = s.dev<double>(4);
    CU(cudaMemsetAsync(out4, 0, sizeof(double) * 4, c.stream));
    return finish_call(s, out4, ci.pairs, ci.nonfinite, n, nullptr);
  });
}

int eb2_ball_count(int dev, const double* coords, const int32_t* cls, int64_t n, int d, int ncls, int within_class,
                   const double* radius, uint32_t flags, int64_t* out) {
  if (int rc0 = check_common(coords, n, d, 1)) return rc0;
  if (!out || !radius) return fail(EB2_ERR_ARG, "out/radius is NULL");
  if (flags & EB2_FLAG_DEVICE_INPUT) return fail(EB2_ERR_ARG, "eb2_ball_count takes host buffers only");
  return guarded(dev, [&](Ctx& c) {
    Scratch s(c);
    CallInit ci = begin_call(s);
    // the radius travels as one more row so that it is permuted with the points
    double* raw = s.dev<double>(static_cast<size_t>(d + 1) * n);
    CU(cudaMemcpyAsync(raw, coords, sizeof(double) * d * n, cudaMemcpyHostToDevice, c.stream));
    CU(cudaMemcpyAsync(raw + static_cast<int64_t>(d) * n, radius, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
    nonfinite_kernel<<<cdiv(static_cast<int64_t>(d) * n, 256), 256, 0, c.stream>>>(raw, static_cast<int64_t>(d) * n, ci.nonfinite);
    s.launches++;
    const int* cls_dev = nullptr;
    std::vector<int> sizes;
    const bool seg = cls && within_class;
    if (seg && !stage_classes(s, cls, n, ncls, flags, &cls_dev, &sizes)) return fail(EB2_ERR_ARG, "class id outside [0, ncls)");
    const bool prune = !(flags & EB2_FLAG_NO_PRUNE);
    PointSet ps = build_point_set(s, raw, d + 1, n, seg ? cls_dev : nullptr, sizes, prune ? 0 : -1);
    TileSet self = make_tiles(s, ps, 0, n, true, 0, 0);
    mark(s, 1); mark(s, 2);
    int* cnt = s.dev<int>(ps.stride);
    const double* rad = ps.P + static_cast<int64_t>(d) * ps.stride;
    if (d == 1 && !(flags & EB2_FLAG_BRUTE_COUNT) && prune) {
      // segment slices of row 0 are ascending: exact binary search
      TileSet sl = make_tiles(s, ps, 0, n, true, 0, 0);
      // search_kernel indexes `sorted + c_lo` with c_len valid entries, which is exactly the segment
      run_search(s, ps.P, rad, ps.P, sl, cnt);
    } else {
      CountOut co; co.s = cnt;
      run_count(s, ps, ps, d, 0, rows_range(0, d), rows_range(0, d), RowSel{}, RowSel{}, rad, self, prune, co, ci.pairs);
    }
    mark(s, 3); mark(s, 4);
    const int* cnts[3] = {cnt, nullptr, nullptr};
    Outputs o; o.cnt[0] = out;
    export_outputs(s, ps, nullptr, cnts, o);
    double* out4 = s.dev<double>(4);
    CU(cudaMemsetAsync(out4, 0, sizeof(double) * 4, c.stream));
    return finish_call(s, out4, ci.pairs, ci.nonfinite, n, nullptr);
  });
}

}  // extern "C"
