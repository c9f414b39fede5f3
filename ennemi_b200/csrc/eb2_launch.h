// ennemi_b200 — host-side launchers of the templated kernels (one translation unit each so the
// instantiations build in parallel).
#pragma once
#include "eb2_kernels.cuh"

namespace eb2 {
// D = dimension of the search space (1..kMaxDim); k+1 <= 8 uses the register top-k variants,
// larger k the heap variant (args.heap must then hold (k+1) * grid * tile_rows(qpt) doubles).
// qpt = query rows per thread (1 or 2): tiles are tile_rows(qpt) rows
cudaError_t launch_knn(int D, int qpt, const KnnArgs& args, int grid, cudaStream_t stream);
int knn_grid(int k, int ntiles, int sm_count);
// finishes the queries knn_kernel deferred (register top-k variants only)
cudaError_t launch_knn_leftover(int D, const KnnArgs& args, int grid, cudaStream_t stream);
// two-level layout: per chunk of chunk_len(D) slots record the range of row1 and reorder by row2
cudaError_t launch_cell_sort(int D, double* P, int64_t stride, int d, int* slot_row, int64_t n, int row1, int row2,
                             double* cell_lo, double* cell_hi, int nchunks, cudaStream_t stream);
// C shared + E private coordinates
cudaError_t launch_count(int C, int E, int qpt, const CountArgs& args, int grid, cudaStream_t stream);
}  // namespace eb2
