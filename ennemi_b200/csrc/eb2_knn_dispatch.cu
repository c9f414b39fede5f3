// ennemi_b200 — instantiations of knn_kernel<D, K1T, QPT> / knn_leftover_kernel<D, K1T>.
// Compiled once per dimension D (-DEB2_KNN_D=1 .. 12) so that the instantiations build in parallel;
// the D == 0 object holds the run-time dispatch, the grid rule and the cell-sort launcher.
#include "eb2_launch.h"

#ifndef EB2_KNN_D
#define EB2_KNN_D 0
#endif

namespace eb2 {

#define EB2_DECLARE(D)                                                                              \
  cudaError_t launch_knn_d##D(int qpt, const KnnArgs& a, int grid, cudaStream_t s);                 \
  cudaError_t launch_knn_leftover_d##D(const KnnArgs& a, int grid, cudaStream_t s);
EB2_DECLARE(1) EB2_DECLARE(2) EB2_DECLARE(3) EB2_DECLARE(4) EB2_DECLARE(5) EB2_DECLARE(6)
EB2_DECLARE(7) EB2_DECLARE(8) EB2_DECLARE(9) EB2_DECLARE(10) EB2_DECLARE(11) EB2_DECLARE(12)
#undef EB2_DECLARE

#if EB2_KNN_D > 0

template <int D, int QPT>
static cudaError_t launch_dq(const KnnArgs& a, int grid, cudaStream_t s) {
  const int k1 = a.k + 1;
  if (k1 <= 4) knn_kernel<D, 4, QPT><<<grid, kThreads, 0, s>>>(a);
  else if (k1 <= 8) knn_kernel<D, 8, QPT><<<grid, kThreads, 0, s>>>(a);
  else knn_kernel<D, 0, QPT><<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

#define EB2_CAT2(a, b) a##b
#define EB2_CAT(a, b) EB2_CAT2(a, b)

cudaError_t EB2_CAT(launch_knn_d, EB2_KNN_D)(int qpt, const KnnArgs& a, int grid, cudaStream_t s) {
  return qpt == 1 ? launch_dq<EB2_KNN_D, 1>(a, grid, s) : launch_dq<EB2_KNN_D, 2>(a, grid, s);
}

cudaError_t EB2_CAT(launch_knn_leftover_d, EB2_KNN_D)(const KnnArgs& a, int grid, cudaStream_t s) {
  const int k1 = a.k + 1;
  if (k1 <= 4) knn_leftover_kernel<EB2_KNN_D, 4><<<grid, kThreads, 0, s>>>(a);
  else if (k1 <= 8) knn_leftover_kernel<EB2_KNN_D, 8><<<grid, kThreads, 0, s>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

#else  // EB2_KNN_D == 0: dispatch

int knn_grid(int k, int ntiles, int sm_count) {
  if (k + 1 <= 8) return ntiles;                 // one CTA per query tile, hardware scheduler balances
  const int cap = sm_count * 2;                  // heap variant is persistent: scratch is per CTA
  return ntiles < cap ? ntiles : cap;
}

cudaError_t launch_knn(int D, int qpt, const KnnArgs& a, int grid, cudaStream_t s) {
  switch (D) {
#define EB2_CASE(D) case D: return launch_knn_d##D(qpt, a, grid, s);
    EB2_CASE(1) EB2_CASE(2) EB2_CASE(3) EB2_CASE(4) EB2_CASE(5) EB2_CASE(6)
    EB2_CASE(7) EB2_CASE(8) EB2_CASE(9) EB2_CASE(10) EB2_CASE(11) EB2_CASE(12)
#undef EB2_CASE
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_knn_leftover(int D, const KnnArgs& a, int grid, cudaStream_t s) {
  switch (D) {
#define EB2_CASE(D) case D: return launch_knn_leftover_d##D(a, grid, s);
    EB2_CASE(1) EB2_CASE(2) EB2_CASE(3) EB2_CASE(4) EB2_CASE(5) EB2_CASE(6)
    EB2_CASE(7) EB2_CASE(8) EB2_CASE(9) EB2_CASE(10) EB2_CASE(11) EB2_CASE(12)
#undef EB2_CASE
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_cell_sort(int D, double* P, int64_t stride, int d, int* slot_row, int64_t n, int row1, int row2,
                             double* cell_lo, double* cell_hi, int nchunks, cudaStream_t s) {
  switch (chunk_len(D)) {
    case 2048: cell_sort_kernel<2048><<<nchunks, 1024, 0, s>>>(P, stride, d, slot_row, n, row1, row2, cell_lo, cell_hi); break;
    case 1024: cell_sort_kernel<1024><<<nchunks, 512, 0, s>>>(P, stride, d, slot_row, n, row1, row2, cell_lo, cell_hi); break;
    case 512: cell_sort_kernel<512><<<nchunks, kThreads, 0, s>>>(P, stride, d, slot_row, n, row1, row2, cell_lo, cell_hi); break;
    case 256: cell_sort_kernel<256><<<nchunks, kThreads, 0, s>>>(P, stride, d, slot_row, n, row1, row2, cell_lo, cell_hi); break;
    case 128: cell_sort_kernel<128><<<nchunks, kThreads, 0, s>>>(P, stride, d, slot_row, n, row1, row2, cell_lo, cell_hi); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

#endif

}  // namespace eb2
