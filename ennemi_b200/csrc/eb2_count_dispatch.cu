// ennemi_b200 — instantiations of count_kernel<C, E>.
#include "eb2_launch.h"

namespace eb2 {

template <int C, int E>
static cudaError_t go(int qpt, const CountArgs& a, int grid, cudaStream_t s) {
  if constexpr (C + E >= 1 && C + E <= kMaxDim) {
    if (qpt == 1) count_kernel<C, E, 1><<<grid, kThreads, 0, s>>>(a);
    else count_kernel<C, E, 2><<<grid, kThreads, 0, s>>>(a);
    return cudaGetLastError();
  } else {
    return cudaErrorInvalidValue;
  }
}

template <int C>
static cudaError_t by_e(int E, int qpt, const CountArgs& a, int grid, cudaStream_t s) {
  switch (E) {
    case 0: return go<C, 0>(qpt, a, grid, s);
    case 1: return go<C, 1>(qpt, a, grid, s);
    case 2: return go<C, 2>(qpt, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_count(int C, int E, int qpt, const CountArgs& a, int grid, cudaStream_t s) {
  switch (C) {
    case 0: return by_e<0>(E, qpt, a, grid, s);
    case 1: return by_e<1>(E, qpt, a, grid, s);
    case 2: return by_e<2>(E, qpt, a, grid, s);
    case 3: return by_e<3>(E, qpt, a, grid, s);
    case 4: return by_e<4>(E, qpt, a, grid, s);
    case 5: return by_e<5>(E, qpt, a, grid, s);
    case 6: return by_e<6>(E, qpt, a, grid, s);
    case 7: return by_e<7>(E, qpt, a, grid, s);
    case 8: return by_e<8>(E, qpt, a, grid, s);
    case 9: return by_e<9>(E, qpt, a, grid, s);
    case 10: return by_e<10>(E, qpt, a, grid, s);
    case 11: return by_e<11>(E, qpt, a, grid, s);
    case 12: return by_e<12>(E, qpt, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace eb2
