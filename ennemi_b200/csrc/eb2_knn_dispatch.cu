// ennemi_b200 — instantiations of knn_kernel<D, K1T>.
#include "eb2_launch.h"

namespace eb2 {

int knn_grid(int k, int ntiles, int sm_count) {
  if (k + 1 <= 8) return ntiles;                 // one CTA per query tile, hardware scheduler balances
  const int cap = sm_count * 2;                  // heap variant is persistent: scratch is per CTA
  return ntiles < cap ? ntiles : cap;
}

template <int D, int QPT>
static cudaError_t launch_dq(const KnnArgs& a, int grid, cudaStream_t s) {
  const int k1 = a.k + 1;
  if (k1 <= 4) knn_kernel<D, 4, QPT><<<grid, kThreads, 0, s>>>(a);
  else if (k1 <= 8) knn_kernel<D, 8, QPT><<<grid, kThreads, 0, s>>>(a);
  else knn_kernel<D, 0, QPT><<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}
template <int D>
static cudaError_t launch_d(int qpt, const KnnArgs& a, int grid, cudaStream_t s) {
  return qpt == 1 ? launch_dq<D, 1>(a, grid, s) : launch_dq<D, 2>(a, grid, s);
}

template <int D>
static cudaError_t launch_left_d(const KnnArgs& a, int grid, cudaStream_t s) {
  const int k1 = a.k + 1;
  if (k1 <= 4) knn_leftover_kernel<D, 4><<<grid, kThreads, 0, s>>>(a);
  else if (k1 <= 8) knn_leftover_kernel<D, 8><<<grid, kThreads, 0, s>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_knn_leftover(int D, const KnnArgs& a, int grid, cudaStream_t s) {
  switch (D) {
    case 1: return launch_left_d<1>(a, grid, s);
    case 2: return launch_left_d<2>(a, grid, s);
    case 3: return launch_left_d<3>(a, grid, s);
    case 4: return launch_left_d<4>(a, grid, s);
    case 5: return launch_left_d<5>(a, grid, s);
    case 6: return launch_left_d<6>(a, grid, s);
    case 7: return launch_left_d<7>(a, grid, s);
    case 8: return launch_left_d<8>(a, grid, s);
    case 9: return launch_left_d<9>(a, grid, s);
    case 10: return launch_left_d<10>(a, grid, s);
    case 11: return launch_left_d<11>(a, grid, s);
    case 12: return launch_left_d<12>(a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_knn(int D, int qpt, const KnnArgs& a, int grid, cudaStream_t s) {
  switch (D) {
    case 1: return launch_d<1>(qpt, a, grid, s);
    case 2: return launch_d<2>(qpt, a, grid, s);
    case 3: return launch_d<3>(qpt, a, grid, s);
    case 4: return launch_d<4>(qpt, a, grid, s);
    case 5: return launch_d<5>(qpt, a, grid, s);
    case 6: return launch_d<6>(qpt, a, grid, s);
    case 7: return launch_d<7>(qpt, a, grid, s);
    case 8: return launch_d<8>(qpt, a, grid, s);
    case 9: return launch_d<9>(qpt, a, grid, s);
    case 10: return launch_d<10>(qpt, a, grid, s);
    case 11: return launch_d<11>(qpt, a, grid, s);
    case 12: return launch_d<12>(qpt, a, grid, s);
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

}  // namespace eb2
