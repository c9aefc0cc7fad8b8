// TF-GridNet building blocks (wesep/models/tfgridnet.py:197-302,
// wesep/modules/tfgridnet/gridnet_block.py:118-284).  The BLSTMs run on the cluster recurrence kernel (lstm_rec.cu), every
// Linear / 1x1 Conv2d and both attention products on the tcgen05 pointwise GEMMs, LayerNorm over channels on the cLN
// kernel (misc.cu); this file holds what was missing:
//   * "head layer norm": PReLU (one slope per head) then LayerNorm over (E, F) of every (batch, head, frame) with a
//     per-(head, e, f) affine — AllHeadPReLULayerNormalization4DCF (gridnet_block.py:255-284) and, with one head,
//     PReLU + LayerNormalization4DCF of attn_concat_proj (gridnet_block.py:103-110,229-252)
//   * row softmax with the 1/sqrt(d) scale of the attention matrix (gridnet_block.py:213-214)
//   * the per-utterance (unbiased) standard deviation of the RMS normalisation (tfgridnet.py:217-218)
// Maps are act tensors [B][H*E][T*F] (F contiguous).
#include "common.cuh"

namespace wb {

constexpr int HL_THREADS = 256;
constexpr int HL_TT = 8;          // frames per CTA

// block-wide sum of two doubles -> every thread (red: 2 x 8 doubles)
__device__ __forceinline__ void hl_block_sum2(double& a, double& b, double (*red)[8]) {
  a = warp_sum(a); b = warp_sum(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  a = 0.0; b = 0.0;
#pragma unroll
  for (int w = 0; w < HL_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
}

// grid (ceil(T / HL_TT), H, B)
__global__ void __launch_bounds__(HL_THREADS) hln_fwd_kernel(WesepHeadLnArgs a) {
  __shared__ double red[2][8];
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int EF = a.E * a.F;
  const float alpha = __ldg(a.alpha + (a.alpha_per_head ? h : 0));
  const int64_t base = ((int64_t)b * a.H + h) * a.E * a.ld;
  for (int t = blockIdx.x * HL_TT; t < min((blockIdx.x + 1) * HL_TT, a.T); ++t) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = tid; i < EF; i += HL_THREADS) {
      const int e = i / a.F, f = i - e * a.F;
      float v = __ldg(a.x + base + (int64_t)e * a.ld + (int64_t)t * a.F + f);
      v = v >= 0.f ? v : alpha * v;
      s0 += (double)v; s1 = fma((double)v, (double)v, s1);
    }
    hl_block_sum2(s0, s1, red);
    const double mean = s0 / EF, var = fmax(s1 / EF - mean * mean, 0.0);
    const float mu = (float)mean, rs = (float)(1.0 / sqrt(var + (double)a.eps));
    if (tid == 0) {
      const int64_t si = (((int64_t)b * a.H + h) * a.T + t) * 2;
      a.mr[si] = mu; a.mr[si + 1] = rs;
    }
    for (int i = tid; i < EF; i += HL_THREADS) {
      const int e = i / a.F, f = i - e * a.F;
      const int64_t o = base + (int64_t)e * a.ld + (int64_t)t * a.F + f;
      float v = __ldg(a.x + o);
      v = v >= 0.f ? v : alpha * v;
      a.y[o] = (v - mu) * rs * __ldg(a.gamma + (int64_t)h * EF + i) + __ldg(a.beta + (int64_t)h * EF + i);
    }
  }
}

// backward: gx, and per CTA partial (dgamma, dbeta) in shared memory + dalpha in a register, flushed with atomics.
// dynamic shared memory: 2 * E * F floats
__global__ void __launch_bounds__(HL_THREADS) hln_bwd_kernel(WesepHeadLnArgs a) {
  extern __shared__ float hl_acc[];
  __shared__ double red[2][8];
  __shared__ float red_a[8];
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int EF = a.E * a.F;
  float* dg = hl_acc;
  float* db = hl_acc + EF;
  for (int i = tid; i < EF; i += HL_THREADS) { dg[i] = 0.f; db[i] = 0.f; }
  const float alpha = __ldg(a.alpha + (a.alpha_per_head ? h : 0));
  const int64_t base = ((int64_t)b * a.H + h) * a.E * a.ld;
  float dal = 0.f;
  for (int t = blockIdx.x * HL_TT; t < min((blockIdx.x + 1) * HL_TT, a.T); ++t) {
    const int64_t si = (((int64_t)b * a.H + h) * a.T + t) * 2;
    const float mu = __ldg(a.mr + si), rs = __ldg(a.mr + si + 1);
    double s0 = 0.0, s1 = 0.0;
    for (int i = tid; i < EF; i += HL_THREADS) {
      const int e = i / a.F, f = i - e * a.F;
      const int64_t o = base + (int64_t)e * a.ld + (int64_t)t * a.F + f;
      const float x = __ldg(a.x + o);
      const float v = x >= 0.f ? x : alpha * x;
      const float xh = (v - mu) * rs;
      const float g = __ldg(a.gy + o);
      const float gh = g * __ldg(a.gamma + (int64_t)h * EF + i);
      dg[i] += g * xh;                    // each (thread, i) pair is private: no race inside the CTA
      db[i] += g;
      s0 += (double)gh; s1 = fma((double)gh, (double)xh, s1);
    }
    hl_block_sum2(s0, s1, red);
    const float m0 = (float)(s0 / EF), m1 = (float)(s1 / EF);
    for (int i = tid; i < EF; i += HL_THREADS) {
      const int e = i / a.F, f = i - e * a.F;
      const int64_t o = base + (int64_t)e * a.ld + (int64_t)t * a.F + f;
      const float x = __ldg(a.x + o);
      const float v = x >= 0.f ? x : alpha * x;
      const float xh = (v - mu) * rs;
      const float gh = __ldg(a.gy + o) * __ldg(a.gamma + (int64_t)h * EF + i);
      const float gv = rs * (gh - m0 - xh * m1);
      a.gx[o] = x > 0.f ? gv : alpha * gv;          // nn.PReLU: slope alpha for x <= 0
      if (x <= 0.f) dal = fmaf(gv, x, dal);
    }
  }
  __syncthreads();
  for (int i = tid; i < EF; i += HL_THREADS) {
    atomicAdd(a.dgamma + (int64_t)h * EF + i, dg[i]);
    atomicAdd(a.dbeta + (int64_t)h * EF + i, db[i]);
  }
  dal = warp_sum(dal);
  if ((tid & 31) == 0) red_a[tid >> 5] = dal;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < HL_THREADS / 32; ++w) s += red_a[w];
    atomicAdd(a.dalpha + (a.alpha_per_head ? h : 0), s);
  }
}

// ------------------------------------------------------------------------------------------------ row softmax
// y[r][c] = softmax_c(scale * x[r][c]) for c < C; columns [C, ld) of y are zeroed (the matrix is then a GEMM operand)
__global__ void __launch_bounds__(256) softmax_fwd_kernel(WesepSoftmaxArgs a) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= a.rows) return;
  const float* x = a.x + r * a.ld;
  float* y = a.y + r * a.ld;
  float mx = -INFINITY;
  for (int c = lane; c < a.C; c += 32) mx = fmaxf(mx, __ldg(x + c) * a.scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int c = lane; c < a.C; c += 32) s += __expf(__ldg(x + c) * a.scale - mx);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int c = lane; c < a.ld; c += 32) y[c] = c < a.C ? __expf(__ldg(x + c) * a.scale - mx) * inv : 0.f;
}
// gx = scale * y * (gy - sum_c gy y)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(WesepSoftmaxArgs a) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= a.rows) return;
  const float* y = a.y + r * a.ld;
  const float* gy = a.gy + r * a.ld;
  float* gx = a.gx + r * a.ld;
  float d = 0.f;
  for (int c = lane; c < a.C; c += 32) d = fmaf(__ldg(gy + c), __ldg(y + c), d);
  d = warp_sum(d);
  for (int c = lane; c < a.ld; c += 32) gx[c] = c < a.C ? a.scale * __ldg(y + c) * (__ldg(gy + c) - d) : 0.f;
}

// ------------------------------------------------------------------------------------------------ row std (unbiased)
__global__ void __launch_bounds__(256) rowstd_kernel(const float* __restrict__ x, int64_t ld, int L, float* __restrict__ std_out,
                                                     float* __restrict__ inv_out) {
  __shared__ double red[2][8];
  const int64_t r = blockIdx.x;
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < L; i += 256) {
    const double v = (double)__ldg(x + r * ld + i);
    s0 += v; s1 = fma(v, v, s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    const double var = fmax((b - a * a / L) / (double)(L > 1 ? L - 1 : 1), 0.0);
    const float sd = (float)sqrt(var);
    std_out[r] = sd;
    inv_out[r] = 1.f / sd;
  }
}

}  // namespace wb

using namespace wb;

static int check_hln(const WesepHeadLnArgs* a) {
  if (!a || a->B <= 0 || a->B > 65535 || a->H <= 0 || a->H > 65535 || a->E <= 0 || a->T <= 0 || a->F <= 0) return fail(-1, "head_ln: bad shape");
  if (a->ld < (int64_t)a->T * a->F) return fail(-1, "head_ln: row stride < T * F");
  if ((int64_t)a->E * a->F > 24576) return fail(-2, "head_ln: E * F > 24576 (shared-memory accumulators)");
  if (!a->x || !a->alpha || !a->gamma || !a->beta || !a->mr) return fail(-1, "head_ln: null buffer");
  return 0;
}
extern "C" int wesep_b200_head_ln_fwd(const WesepHeadLnArgs* a, void* stream) {
  if (int rc = check_hln(a)) return rc;
  if (!a->y) return fail(-1, "head_ln: null output");
  hln_fwd_kernel<<<dim3(cdiv(a->T, HL_TT), a->H, a->B), HL_THREADS, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("head_ln_fwd");
  return 0;
}
extern "C" int wesep_b200_head_ln_bwd(const WesepHeadLnArgs* a, void* stream) {
  if (int rc = check_hln(a)) return rc;
  if (!a->gy || !a->gx || !a->dgamma || !a->dbeta || !a->dalpha) return fail(-1, "head_ln: null gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t EF = (size_t)a->E * a->F;
  WB_CUDA(cudaMemsetAsync(a->dgamma, 0, (size_t)a->H * EF * sizeof(float), st));
  WB_CUDA(cudaMemsetAsync(a->dbeta, 0, (size_t)a->H * EF * sizeof(float), st));
  WB_CUDA(cudaMemsetAsync(a->dalpha, 0, (size_t)(a->alpha_per_head ? a->H : 1) * sizeof(float), st));
  const size_t smem = 2 * EF * sizeof(float);
  WB_CUDA(cudaFuncSetAttribute(hln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hln_bwd_kernel<<<dim3(cdiv(a->T, HL_TT), a->H, a->B), HL_THREADS, smem, st>>>(*a);
  WB_LAUNCH_CHECK("head_ln_bwd");
  return 0;
}

static int check_sm(const WesepSoftmaxArgs* a) {
  if (!a || a->rows <= 0 || a->C <= 0 || a->ld < a->C) return fail(-1, "softmax: bad shape");
  return 0;
}
extern "C" int wesep_b200_softmax_fwd(const WesepSoftmaxArgs* a, void* stream) {
  if (int rc = check_sm(a)) return rc;
  if (!a->x || !a->y) return fail(-1, "softmax: null buffer");
  softmax_fwd_kernel<<<cdiv(a->rows, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("softmax_fwd");
  return 0;
}
extern "C" int wesep_b200_softmax_bwd(const WesepSoftmaxArgs* a, void* stream) {
  if (int rc = check_sm(a)) return rc;
  if (!a->y || !a->gy || !a->gx) return fail(-1, "softmax: null buffer");
  softmax_bwd_kernel<<<cdiv(a->rows, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("softmax_bwd");
  return 0;
}

extern "C" int wesep_b200_rowstd(const WesepRowStdArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->L <= 0 || a->ld < a->L || !a->x || !a->std || !a->inv_std) return fail(-1, "rowstd: bad arguments");
  rowstd_kernel<<<a->n, 256, 0, (cudaStream_t)stream>>>(a->x, a->ld, a->L, a->std, a->inv_std);
  WB_LAUNCH_CHECK("rowstd");
  return 0;
}
