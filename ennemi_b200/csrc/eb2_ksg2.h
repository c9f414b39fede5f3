// ennemi_b200 — the bivariate KSG pipeline ("k2"): sort-free adaptive grid.  Host-side launch API; kernels in
// eb2_ksg2.cu.
//
// One *column* is a prepared variable (n doubles in the caller's row order).  One *problem* is a pair of columns
// (x, y) whose KSG estimate (_entropy_estimators.py:69-113) is wanted.  Every kernel takes arrays of columns /
// problems and is launched once for all of them (blockIdx.y = column / problem): a single estimate is a batch of
// one, pairwise_mi a batch of hundreds of pairs that share their columns.
//
// Nothing is sorted.  A column is cut into BUCKETS by a monotone function of the value (sampled quantiles, each
// quantile stretch cut linearly into kSub parts); the rows of a bucket are contiguous after one counting scatter.
// Inside a bucket, values are grouped into CELLS by a second monotone (linear) function over the bucket's own
// range, again by counting.  Because both functions are monotone and the SAME device function is used to build and
// to query, "every value in a cell below cell(t) is below t" holds exactly, and the few values in boundary cells are
// decided by the reference's exact predicates - counts and distances are bit-identical to an all-pairs evaluation.
//
// Per column (k2::colgrid: six launches for any number of columns)
//   split / clo / csc   the bucket function;  bkt: row -> bucket;  count / boff: rows per bucket, first slot
//   srow / sval         rows and values grouped by bucket
//   fval / fstart       values grouped by (bucket, fine cell) and the first slot of every fine cell: what the marginal
//                       neighbour counts (query_ball_point(..., return_length=True), :109-110) are read from
// Per problem (x = column cx, y = column cy)
//   px, py, prow, pbkt  the point set in SLOT order: the rows of x-bucket b occupy slots [boff[b], boff[b+1]), grouped
//                       into cells of the bucket's own y range (about one row per cell); cstart: first slot per cell
//   eps                 (k+1)-th neighbour distance per slot
//   acc                 the digamma sum as a 128-bit fixed-point integer (2^-48 units) + zero-count counters: integer
//                       addition is associative, so the sum does not depend on slot order, launch geometry or the
//                       number of GPUs that shared the rows
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb2 {
namespace k2 {

constexpr int kSub = 1;              // linear parts per quantile stretch (1: buckets are the quantile stretches themselves)
constexpr int kMaxCoarse = 1024;     // quantile stretches per column (at most)
constexpr int kMaxBuckets = kSub * kMaxCoarse;
constexpr int kCoarseRows = 2048;    // target rows per quantile stretch
constexpr int kOversample = 8;       // sample values per quantile stretch (at most; 4,096 samples in all)
constexpr int kBucketCap = 8192;     // most rows a bucket may hold (cells are built by one CTA in shared memory)
constexpr int kMaxCells = 4096;      // most cells per bucket
constexpr int kFixedBits = 48;       // digamma terms are accumulated in units of 2^-48

// flag bits (columns and problems)
constexpr int kFlagNaN = 1;          // NaN among the inputs of a prepared column (set by prep_kernel)
constexpr int kFlagNonFinite = 2;    // non-finite prepared value
constexpr int kFlagOverflow = 16;    // a bucket outgrew kBucketCap: the caller repeats the estimate on the general path

struct Plan {
  int64_t n = 0;
  int Bc = 1;             // quantile stretches
  int NB = kSub;          // buckets
  int over = 1;           // samples per stretch
};
// n rows -> grid sizes; ok == false when the pipeline does not take this size
Plan make_plan(int64_t n, bool* ok);

struct FineGrid {         // fine-cell map of one bucket, packed: one 32-byte load per threshold
  double vlo, fsc;
  int boff, ncell;
};

struct Col {
  const double* vals;     // [n] row order
  double* split;          // [kMaxCoarse] upper ends of the quantile stretches (Bc - 1 used)
  double* clo;            // [kMaxCoarse] lower end of the linear map of stretch c
  double* csc;            // [kMaxCoarse] its scale (kSub / width, 0 for a degenerate stretch)
  double* ssort;          // [kMaxCoarse * kOversample] the sample values in ascending order
  double* sraw;           // [kMaxCoarse * kOversample] ... as gathered
  int* count;             // [kMaxBuckets + 1]
  int* fill;              // [kMaxBuckets + 1]
  int* boff;              // [kMaxBuckets + 1] first slot of bucket b; boff[NB] = n
  double* vlo;            // [kMaxBuckets] smallest value of the bucket (NaN: empty)
  double* vhi;            // [kMaxBuckets] largest value
  double* fsc;            // [kMaxBuckets] scale of the fine-cell map
  int* ncell;             // [kMaxBuckets] fine cells of the bucket
  FineGrid* fg;           // [kMaxBuckets] (vlo, fsc, boff, ncell) packed
  unsigned short* bkt;    // [n] row -> bucket
  int* srow;              // [n] rows grouped by bucket
  double* sval;           // [n] their values
  double* fval;           // [n] values grouped by (bucket, fine cell)
  int* fstart;            // [n + 2] first slot of fine cell boff[b] + g; unused entries and the tail hold the next start
  int* flag;              // kFlag* bits of this column
};
size_t col_bytes(int64_t n);                                   // bytes of everything a Col points to besides vals
Col carve_col(char* base, int64_t n, const double* vals);      // base: 256-byte aligned block of col_bytes(n) bytes

struct LeftEnt {
  int slot;
  int rstart;     // buckets [rstart, NB) are still to be examined on the right (NB: none)
  int lend;       // buckets [0, lend) on the left (0: none)
  int skip_a;     // slots [skip_a, skip_a + 8) were examined already (the seeds; -16: none are among the remaining buckets)
};

struct Prob {
  int cx, cy;               // columns
  double* px;               // [n]
  double* py;               // [n]
  int* prow;                // [n] slot -> row
  unsigned short* pbkt;     // [n] slot -> x bucket
  int* cstart;              // [n + 2] first slot of cell boff[b] + f
  double* bymin;            // [kMaxBuckets] y-cell map of x-bucket b: cell = floor((y - bymin) * bysc)
  double* bysc;             // [kMaxBuckets]
  double* eps;              // [n] per slot
  LeftEnt* left;            // [left_cap]
  double* left_best;        // [left_cap][K1T]
  unsigned int* left_count; // [1]
  unsigned long long* acc;  // [4] fixed-point sum (lo, hi), zero counts of n_x and n_y
  double* out;              // [16] result block: 0 sum (double, informative), 1 zeros_x, 2 zeros_y, [4] pairs (u64),
                            //      [5] flags (int), [6] rows reduced (u64), [8] sum lo (u64), [9] sum hi (i64)
  // optional per-row outputs (device, row order), or NULL
  double* eps_row;
  long long* nx_row;
  long long* ny_row;
};
size_t prob_bytes(const Plan& p, int k1t);
Prob carve_prob(char* base, const Plan& p, int k1t, int cx, int cy);

// shard of the x-buckets a call works on: buckets whose first slot lies in [row_lo, row_hi)
struct Shard {
  int64_t row_lo, row_hi;
};

cudaError_t init();       // opt-in shared memory sizes (once per device)

// cols / probs: DEVICE arrays.  Launch counts are added to *launches.
cudaError_t colgrid(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches);
// ... in two steps: the bucket structure (all a layout / search needs of the x column) and the fine cells (counts only)
cudaError_t colgrid_buckets(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches);
cudaError_t colgrid_cells(const Col* cols, int ncol, const Plan& p, cudaStream_t s, int* launches);
cudaError_t layout(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches);
cudaError_t knn(const Col* cols, const Prob* probs, int nprob, const Plan& p, int k, const Shard& sh, int sm_count,
                cudaStream_t s, int* launches);
cudaError_t count_psi(const Col* cols, const Prob* probs, int nprob, const Plan& p, const Shard& sh, const double* psi_tab,
                      int tab_n, cudaStream_t s, int* launches);
// fills probs[p].out from the accumulators and ORs the column flags into out[5]
cudaError_t finalize(const Col* cols, const Prob* probs, int nprob, const Plan& p, cudaStream_t s, int* launches);


// ---- three-level grid for spaces of three and more dimensions (k-NN entropy :21-42, Frenzel-Pompe :116-156) ----------
// Same construction one level deeper: buckets of coordinate 0 (a Col built by colgrid_buckets), and inside every bucket
// cells over the bucket's own ranges of coordinates 1 and 2 (C1 x C2 <= 4,096 cells, about one row each), built by
// counting.  All D coordinates travel with the rows; a window in coordinates 0..2 is a few runs of consecutive slots.
constexpr int kG3MaxD = 8;

struct Grid3 {
  int D;                      // coordinates of the space (3 .. kG3MaxD); 0, 1, 2 are the grid coordinates
  int G;                      // grid levels in use: 2 (cells on coordinate 1 only) or 3
  const double* raw[kG3MaxD]; // coordinate rows, caller's row order (raw[0] = the bucket column's vals)
  double* pc[kG3MaxD];        // [n] coordinates in slot order
  int* prow;                  // [n] slot -> row
  unsigned short* pbkt;       // [n] slot -> bucket
  int* cstart;                // [n + 2] first slot of cell boff[b] + c1 * C2 + c2
  double* g1min; double* g1sc; double* g2min; double* g2sc;   // [kMaxBuckets] cell maps per bucket
  int* c1n; int* c2n;         // [kMaxBuckets] cells per bucket along coordinates 1 and 2
  double* eps_row;            // [n] k-th distance per ROW
  double* eps;                // [n] per slot
  LeftEnt* left; double* left_best; unsigned int* left_count;
  int* heavy; unsigned int* heavy_count;   // [n] slots whose counts are taken by a warp (wide radius), their number
  int* cnt_row[3];            // per row: n_z, n_xz, n_yz (Frenzel-Pompe), or NULL
  unsigned long long* pairs;  // work counter
  int* flag;                  // ORed with the bucket column's flag by the kernels' callers
};
size_t grid3_bytes(const Plan& p, int D, int k1t);
Grid3 carve_grid3(char* base, const Plan& p, int D, int G, int k1t);
// plan for a D-dimensional space (bucket size differs from the bivariate one)
Plan make_plan3(int64_t n, int D, bool* ok);

cudaError_t layout3(const Col* col0, const Grid3* g, const Plan& p, cudaStream_t s, int* launches);
cudaError_t knn3(const Col* col0, const Grid3* g, const Grid3& host_copy, const Plan& p, int k, int sm_count, cudaStream_t s,
                 int* launches);
// Frenzel-Pompe counts: coordinates [0, C) are the condition, C and C + 1 are x and y
cudaError_t count3(const Col* col0, const Grid3* g, const Grid3& host_copy, const Plan& p, int C, cudaStream_t s, int* launches);

}  // namespace k2
}  // namespace eb2
