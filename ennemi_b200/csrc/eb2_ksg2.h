// ennemi_b200 — the bivariate KSG pipeline ("k2"): sample-sorted columns, bucket layout, warp-per-32-queries
// neighbour search, fused marginal counts + digamma.  Host-side launch API; kernels in eb2_ksg2.cu.
//
// One *column* is a prepared variable (n doubles in the caller's row order).  One *problem* is a pair of columns
// (x, y) whose KSG estimate (_entropy_estimators.py:69-113) is wanted.  Every kernel takes arrays of columns /
// problems and is launched once for all of them (blockIdx.y / .z = column / problem): a single estimate is a batch
// of one, pairwise_mi a batch of hundreds of pairs that share 64 sorted columns.
//
// Per column (k2_colsort: four launches for any number of columns)
//   splitters  B-1 values from a sorted regular sample (8 per bucket): bucket b holds split[b-1] < v <= split[b]
//   sorted     the column in ascending order (bucket-major, every bucket sorted by one CTA in shared memory)
//   perm       rank -> row;  bid: row -> bucket;  count / boff / soff: rows per bucket, first rank, first slot
//   lo / hi    value range of every bucket
// Per problem
//   px, py, slot_row   the point set in SLOT order: the rows of x-bucket b ("chunk" b) occupy the slots
//                      [soff[b], soff[b] + count[b]) in ascending y (ties by row), padded with NaN / -1 to a
//                      multiple of 32 slots, so that a warp's 32 queries are 32 y-neighbours of one chunk
//   eps                (k+1)-th neighbour distance per slot
//   partial            per 256-slot block: sum of psi(n_x) [+ zero count], sum of psi(n_y) [+ zero count]
//   out                4 doubles (sum, zeros_x, zeros_y, 0) + pair counter (u64) + flags (int) + rows reduced (u64 at [6])
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb2 {
namespace k2 {

constexpr int kBucketMean = 1024;    // target rows per bucket
constexpr int kBucketCap = 4096;     // most rows one bucket may hold (one CTA sorts it in shared memory)
constexpr int kMaxBuckets = 1024;
constexpr int kOversample = 8;       // sample values per bucket the splitters are chosen from
constexpr int kBlockSlots = 256;     // slots per reduction block (and per CTA of the count kernel)

// flag bits (columns and problems)
constexpr int kFlagNaN = 1;          // NaN among the inputs of a prepared column (set by prep_kernel)
constexpr int kFlagNonFinite = 2;    // non-finite prepared value
constexpr int kFlagOverflow = 16;    // a bucket outgrew kBucketCap: the caller repeats the estimate on the general path

struct Plan {
  int64_t n = 0;
  int B = 1;              // buckets per column
  int over = 1;           // samples per bucket
  int64_t smax = 0;       // upper bound of the slots of a problem (multiple of kBlockSlots)
  int nblk = 0;           // smax / kBlockSlots
};
// n rows -> bucket count etc.; ok == false when the pipeline does not take this size
Plan make_plan(int64_t n, bool* ok);

struct Col {
  const double* vals;     // [n] row order
  double* sorted;         // [n]
  int* perm;              // [n] rank -> row
  unsigned short* bid;    // [n] row -> bucket
  double* st_val;         // [n] scatter staging
  int* st_row;            // [n]
  double* split;          // [kMaxBuckets]
  int* count;             // [kMaxBuckets]
  int* fill;              // [kMaxBuckets]
  int* boff;              // [kMaxBuckets + 1] first rank of bucket b
  int* soff;              // [kMaxBuckets + 1] first slot of chunk b (multiples of 32); soff[B] = slots in use
  double* lo;             // [kMaxBuckets]
  double* hi;             // [kMaxBuckets]
  int* flag;              // kFlag* bits of this column
};
// bytes of per-column scratch besides vals (everything a Col points to), for n rows
size_t col_bytes(int64_t n);
// carves a Col out of one allocation of col_bytes(n) bytes (256-byte aligned base)
Col carve_col(char* base, int64_t n, const double* vals);

struct LeftEnt {
  int slot;
  int rstart;     // chunks [rstart, B) are still to be examined on the right (B: none)
  int lend;       // chunks [0, lend) on the left (0: none)
};

struct Prob {
  int cx, cy;               // columns
  double* px;               // [smax]
  double* py;               // [smax]
  int* slot_row;            // [smax]
  double* eps;              // [smax]
  LeftEnt* left;            // [left_cap]
  double* left_best;        // [left_cap][K1T]
  unsigned int* left_count; // [1]
  double* partial;          // [2][nblk][2]
  double* out;              // [8] result block: 0..3 sums, [4] pairs (u64), [5] flags (int), [6] rows reduced (u64)
  // optional per-row outputs (device, row order), or NULL
  double* eps_row;
  long long* nx_row;
  long long* ny_row;
};
size_t prob_bytes(const Plan& p, int k1t);
Prob carve_prob(char* base, const Plan& p, int k1t, int cx, int cy);

// shard of the slot blocks a call works on: blocks whose index maps into rows [row_lo, row_hi) of n
struct Shard {
  int64_t row_lo, row_hi, n;
};

cudaError_t init();       // opt-in shared memory sizes (once per process)

// cols / probs: DEVICE arrays.  Launch counts are returned through *launches.
cudaError_t colsort(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches);
cudaError_t layout(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches);
cudaError_t knn(const Col* cols, const Prob* probs, int nprob, const Plan& p, int k, const Shard& sh, int sm_count,
                cudaStream_t s, int* launches);
cudaError_t count_psi(const Col* cols, const Prob* probs, int nprob, const Plan& p, const Shard& sh, const double* psi_tab,
                      int tab_n, cudaStream_t s, int* launches);
// folds the block partials of every problem in a fixed order into probs[p].out[0..3] and ORs the column flags into out[5]
cudaError_t finalize(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches);

}  // namespace k2
}  // namespace eb2
