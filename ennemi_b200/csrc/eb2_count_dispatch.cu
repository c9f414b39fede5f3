// ennemi_b200 — instantiations of count_kernel<C, E, QPT>.
// Compiled once per number of shared coordinates C (-DEB2_COUNT_C=0 .. 12, E = 0..2 inside) so that the
// instantiations build in parallel; the object built with -DEB2_COUNT_C=-1 holds the run-time dispatch.
#include "eb2_launch.h"

#ifndef EB2_COUNT_C
#define EB2_COUNT_C -1
#endif

namespace eb2 {

#define EB2_DECLARE(C) cudaError_t launch_count_c##C(int E, int qpt, const CountArgs& a, int grid, cudaStream_t s);
EB2_DECLARE(0) EB2_DECLARE(1) EB2_DECLARE(2) EB2_DECLARE(3) EB2_DECLARE(4) EB2_DECLARE(5) EB2_DECLARE(6)
EB2_DECLARE(7) EB2_DECLARE(8) EB2_DECLARE(9) EB2_DECLARE(10) EB2_DECLARE(11) EB2_DECLARE(12)
#undef EB2_DECLARE

#if EB2_COUNT_C >= 0

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

#define EB2_CAT2(a, b) a##b
#define EB2_CAT(a, b) EB2_CAT2(a, b)

cudaError_t EB2_CAT(launch_count_c, EB2_COUNT_C)(int E, int qpt, const CountArgs& a, int grid, cudaStream_t s) {
  switch (E) {
    case 0: return go<EB2_COUNT_C, 0>(qpt, a, grid, s);
    case 1: return go<EB2_COUNT_C, 1>(qpt, a, grid, s);
    case 2: return go<EB2_COUNT_C, 2>(qpt, a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

#else  // dispatch

cudaError_t launch_count(int C, int E, int qpt, const CountArgs& a, int grid, cudaStream_t s) {
  switch (C) {
#define EB2_CASE(C) case C: return launch_count_c##C(E, qpt, a, grid, s);
    EB2_CASE(0) EB2_CASE(1) EB2_CASE(2) EB2_CASE(3) EB2_CASE(4) EB2_CASE(5) EB2_CASE(6)
    EB2_CASE(7) EB2_CASE(8) EB2_CASE(9) EB2_CASE(10) EB2_CASE(11) EB2_CASE(12)
#undef EB2_CASE
    default: return cudaErrorInvalidValue;
  }
}

#endif

}  // namespace eb2
