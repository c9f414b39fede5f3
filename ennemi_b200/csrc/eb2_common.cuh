// ennemi_b200 — shared device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb2 {

constexpr int kMaxDim = 12;        // dimensions with specialised (register-resident) kernels
constexpr int kMaxDimAny = 32;     // EB2_MAX_DIM: beyond kMaxDim a generic run-time-dimension brute-force path is used
constexpr int kThreads = 256;      // threads per CTA in the all-pairs kernels
constexpr int kMaxQpt = 2;         // query rows per thread: 2 (512-row tiles) for large sets, 1 (256-row tiles)
                                   // when the set is too small to fill the 148 SMs with 512-row tiles
__host__ __device__ constexpr int tile_rows(int qpt) { return kThreads * qpt; }
constexpr int kSegAlign = 16;      // segments start on 16-slot (128 B) boundaries: TMA bulk copies need 16 B

// Candidate-chunk length (slots) staged in shared memory per step, by dimension of the space.
#ifndef EB2_CHUNK2
#define EB2_CHUNK2 2048
#endif
#ifndef EB2_CHUNK5
#define EB2_CHUNK5 1024
#endif
__host__ __device__ constexpr int chunk_len(int d) { return d <= 2 ? EB2_CHUNK2 : (d <= 5 ? EB2_CHUNK5 : (d <= 8 ? 256 : 128)); }

// One tile of query rows and the candidate segment it is compared against.
struct Tile {
  int q_lo;      // first query slot
  int q_n;       // valid query rows in the tile (<= tile_rows(qpt)); slots past it are NaN padding
  int c_lo;      // first candidate slot of the segment (multiple of kSegAlign)
  int c_len;     // valid candidate rows in the segment (the padded tail up to the next multiple of 16 is NaN)
};

// ---- mbarrier / TMA bulk-copy PTX (cp.async.bulk -> SASS UBLKCP) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- the reference's digamma (_entropy_estimators.py:327-350), n >= 1 ---------------------------
// psi(1) is the literal at :343; everything else the expansion at :348, same operation order.
__device__ __forceinline__ double psi_ref(double y) {
  if (y == 1.0) return -0.5772156649015331;
  const double y2 = y * y;
  return log(y) - pow(y, -6.0) * (y2 * (y2 * (y / 2 + 1.0 / 12) - 1.0 / 120) + 1.0 / 252);
}

// deterministic block-wide sum: warp shuffle tree, then a fixed-order tree over the warps
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* smem /* THREADS/32 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < THREADS / 32) ? smem[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

template <int THREADS>
__device__ __forceinline__ double block_max_bcast(double v, double* smem /* THREADS/32 + 1 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  double r = smem[0];
#pragma unroll
  for (int i = 1; i < THREADS / 32; ++i) r = fmax(r, smem[i]);
  return r;  // valid in every thread
}

template <int THREADS>
__device__ __forceinline__ double block_min_bcast(double v, double* smem) {
  return -block_max_bcast<THREADS>(-v, smem);
}

}  // namespace eb2
