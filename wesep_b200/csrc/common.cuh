// Shared device/host helpers for the wesep_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/wesep_b200.h"

namespace wb {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
extern int g_gemm_mode;
extern int g_gemm_backend;

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
#define WB_LAUNCH_CHECK(what)            \
  do {                                   \
    int _rc = ::wb::check_launch(what);  \
    if (_rc) return _rc;                 \
  } while (0)
#define WB_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      snprintf(::wb::g_err, sizeof(::wb::g_err), "%s: %s", #expr, cudaGetErrorString(_e)); \
      return -3;                                                               \
    }                                                                          \
  } while (0)
#define WB_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) return ::wb::fail(-1, msg); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV floats per thread; result valid in thread 0. `red` = smem float[NV*32].
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float x = lane < nw ? red[i * 32 + lane] : 0.f;
      v[i] = warp_sum(x);
    }
  }
}

__device__ __forceinline__ float prelu_f(float x, float a) { return x > 0.f ? x : a * x; }

// gLN row statistics: stats = (sum, sumsq) over `count` elements -> mean, rstd (biased var, eps in sqrt)
// fp64 division / sqrt expand to long instruction sequences on a narrow pipe (measured: dropping three of them per
// thread per tile from a GEMM epilogue saved 16 % of that kernel), so: one fp64 reciprocal, fp64 only for the
// cancellation-prone E[x^2] - mean^2, and a Newton-refined fp32 rsqrt.
__device__ __forceinline__ void gln_mean_rstd(const double* st, double count, float eps, float& mean, float& rstd) {
  const double ic = __drcp_rn(count);
  const double m = st[0] * ic;
  double var = fma(st[1], ic, -m * m);
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  const float v = (float)var + eps;
  float r = rsqrtf(v);
  r = r * fmaf(-0.5f * v * r, r, 1.5f);
  rstd = r;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace wb
