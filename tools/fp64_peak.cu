// FP64 CUDA-core issue-rate microbenchmark for sm_100a (B200).
// Gives the roofline denominator for the brute-force Chebyshev kernels:
// how many FP64-pipe warp instructions (DADD / DSETP / DFMA) the chip retires per second,
// alone and mixed with the ALU work the real inner loop carries.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

// Pure DADD: CHAINS independent dependent-chains per thread.
__global__ void k_dadd(double* out, double c) {
  double a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = __dadd_rn(a[i], c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Pure DFMA.
__global__ void k_dfma(double* out, double c) {
  double a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = __fma_rn(a[i], c, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DADD + DSETP 1:1 — the shape of the Chebyshev inner loop in one dimension:
// d = q - c ; p = p && (|d| <= thr).  The predicate is folded into an int counter once per
// CHAINS pairs so that the ALU pipe sees little work.
__global__ void k_dadd_dsetp(double* out, double c, double thr) {
  double q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) q[i] = threadIdx.x * 1e-3 + i;
  int cnt = 0;
  double cc = c;
  for (int it = 0; it < ITERS; ++it) {
    int hit;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 d;\n\t"
        "sub.rn.f64 d, %1, %9;\n\t abs.f64 d, d;\n\t setp.le.f64 p, d, %10;\n\t"
        "sub.rn.f64 d, %2, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %3, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %4, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %5, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %6, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %7, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "sub.rn.f64 d, %8, %9;\n\t abs.f64 d, d;\n\t setp.le.and.f64 p, d, %10, p;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(hit)
        : "d"(q[0]), "d"(q[1]), "d"(q[2]), "d"(q[3]), "d"(q[4]), "d"(q[5]), "d"(q[6]), "d"(q[7]),
          "d"(cc), "d"(thr));
    cnt += hit;
    cc = __longlong_as_double(__double_as_longlong(cc) + 1);  // new candidate each iter (ALU, cheap)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = cnt + cc;
}

// Pure DSETP chain.
__global__ void k_dsetp(double* out, double c, double thr) {
  double q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) q[i] = threadIdx.x * 1e-3 + i;
  int cnt = 0;
  double cc = c;
  for (int it = 0; it < ITERS; ++it) {
    int hit;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.le.f64 p, %1, %9;\n\t"
        "setp.le.and.f64 p, %2, %9, p;\n\t"
        "setp.le.and.f64 p, %3, %9, p;\n\t"
        "setp.le.and.f64 p, %4, %9, p;\n\t"
        "setp.le.and.f64 p, %5, %9, p;\n\t"
        "setp.le.and.f64 p, %6, %9, p;\n\t"
        "setp.le.and.f64 p, %7, %9, p;\n\t"
        "setp.le.and.f64 p, %8, %9, p;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(hit)
        : "d"(q[0]), "d"(q[1]), "d"(q[2]), "d"(q[3]), "d"(q[4]), "d"(q[5]), "d"(q[6]), "d"(q[7]),
          "d"(cc), "d"(thr));
    cnt += hit;
    cc = __longlong_as_double(__double_as_longlong(cc) + 1);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = cnt + cc;
}

// DADD + integer compare of the high word (|d|.hi <= thr.hi): FP64 pipe sees only the subtract.
__global__ void k_dadd_isetp(double* out, double c, double thr) {
  double q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) q[i] = threadIdx.x * 1e-3 + i;
  int cnt = 0;
  double cc = c;
  const unsigned thr_hi = (unsigned)__double2hiint(thr);
  for (int it = 0; it < ITERS; ++it) {
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      double d = __dadd_rn(q[i], -cc);
      unsigned h = ((unsigned)__double2hiint(d)) & 0x7fffffffu;
      m = max(m, h);
    }
    cnt += (m <= thr_hi);
    cc = __longlong_as_double(__double_as_longlong(cc) + 1);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = cnt + cc;
}

// Broadcast LDS.128 feed rate together with DADD+DSETP (2 coords per candidate from smem).
__global__ void k_lds_pair(double* out, const double2* __restrict__ cand, int ncand, double thr) {
  extern __shared__ double2 sc[];
  for (int i = threadIdx.x; i < ncand; i += blockDim.x) sc[i] = cand[i];
  __syncthreads();
  const double qx0 = threadIdx.x * 1e-3, qy0 = threadIdx.x * 2e-3;
  const double qx1 = qx0 + 0.5, qy1 = qy0 + 0.5;
  int cnt = 0;
  for (int rep = 0; rep < 16; ++rep) {
#pragma unroll 8
    for (int j = 0; j < ncand; ++j) {
      const double2 c = sc[j];
      bool p0 = (fabs(__dadd_rn(qx0, -c.x)) <= thr) & (fabs(__dadd_rn(qy0, -c.y)) <= thr);
      bool p1 = (fabs(__dadd_rn(qx1, -c.x)) <= thr) & (fabs(__dadd_rn(qy1, -c.y)) <= thr);
      if (p0 | p1) cnt++;
    }
    thr *= 0.999;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = cnt;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int threads = 256, blocks = sms * 8 * 4;   // 8 CTAs/SM resident, 4 waves
  double* out; CK(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", prop.name, sms, prop.clockRate);
  const double total_thread_iters = (double)blocks * threads * ITERS;
  {
    float ms = time_ms([&] { k_dadd<<<blocks, threads>>>(out, 1e-9); });
    printf(", \"dadd_Tinstr_s\": %.3f", total_thread_iters * CHAINS / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001); });
    printf(", \"dfma_Tinstr_s\": %.3f", total_thread_iters * CHAINS / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dsetp<<<blocks, threads>>>(out, 0.5, 0.25); });
    printf(", \"dsetp_Tinstr_s\": %.3f", total_thread_iters * CHAINS / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dadd_dsetp<<<blocks, threads>>>(out, 0.5, 0.25); });
    printf(", \"dadd_dsetp_pairs_T_s\": %.3f, \"dadd_dsetp_fp64_Tinstr_s\": %.3f",
           total_thread_iters * CHAINS / ms * 1e-9, 2 * total_thread_iters * CHAINS / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dadd_isetp<<<blocks, threads>>>(out, 0.5, 0.25); });
    printf(", \"dadd_isetp_pairs_T_s\": %.3f", total_thread_iters * CHAINS / ms * 1e-9);
  }
  {
    const int ncand = 2048;
    double2* cand; CK(cudaMalloc(&cand, sizeof(double2) * ncand));
    CK(cudaMemset(cand, 0, sizeof(double2) * ncand));
    const int b2 = sms * 4 * 4;
    float ms = time_ms([&] { k_lds_pair<<<b2, threads, ncand * sizeof(double2)>>>(out, cand, ncand, 0.3); });
    const double pairs = (double)b2 * threads * 2.0 * ncand * 16;
    printf(", \"lds_pair2d_Tpairs_s\": %.3f, \"lds_pair2d_fp64_Tinstr_s\": %.3f", pairs / ms * 1e-9,
           4 * pairs / ms * 1e-9);
  }
  printf("}\n");
  return 0;
}
