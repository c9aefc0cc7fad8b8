// Channel-wise LayerNorm, framing / overlap-add and small elementwise kernels around the
// encoder / decoder of Spex+ (wesep/modules/tasnet/{encoder,decoder}.py, common/norm.py:51-66).
#include "common.cuh"

namespace wb {

// ------------------------------------------------------------------------------------ cLN
// block (32, 8): x <-> frame t, y <-> channel group. grid (cdiv(T,32), n).
__global__ void __launch_bounds__(256) cln_fwd_kernel(WesepClnFwdArgs a) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = blockIdx.x * 32 + tx, n = blockIdx.y;
  const bool ok = t < a.T;
  const float* x = a.x + (int64_t)n * a.C * a.ldx + t;
  float s = 0.f;
  if (ok)
    for (int c = ty; c < a.C; c += 8) s += __ldg(x + (int64_t)c * a.ldx);
  red[ty][tx] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) mean += red[k][tx];
  mean /= (float)a.C;
  __syncthreads();
  float q = 0.f;
  if (ok)
    for (int c = ty; c < a.C; c += 8) {
      float d = __ldg(x + (int64_t)c * a.ldx) - mean;
      q = fmaf(d, d, q);
    }
  red[ty][tx] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) var += red[k][tx];
  const float rstd = rsqrtf(var / (float)a.C + a.eps);
  if (ok) {
    float* y = a.y + (int64_t)n * a.C * a.ldy + t;
    for (int c = ty; c < a.C; c += 8)
      y[(int64_t)c * a.ldy] = (__ldg(x + (int64_t)c * a.ldx) - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
    if (ty == 0) {
      a.mean[(int64_t)n * a.T + t] = mean;
      a.rstd[(int64_t)n * a.T + t] = rstd;
    }
  }
}

// grid (cdiv(T,1024), n): each block sweeps 32 sub-tiles of 32 frames and keeps its per-channel (dgamma, dbeta)
// partials in shared memory, so global atomics are 2*C per CTA (not per sub-tile).
constexpr int CLN_BWD_FRAMES = 512;
__global__ void __launch_bounds__(256) cln_bwd_kernel(WesepClnBwdArgs a) {
  extern __shared__ float cacc[];  // [2][C]
  __shared__ float red[2][8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, n = blockIdx.y;
  const int nch = (a.C + 7) / 8;  // channels per ty
  for (int i = ty * 32 + tx; i < 2 * a.C; i += 256) cacc[i] = 0.f;
  __syncthreads();
  for (int sub = 0; sub < CLN_BWD_FRAMES / 32; ++sub) {
    const int t = blockIdx.x * CLN_BWD_FRAMES + sub * 32 + tx;
    if (blockIdx.x * CLN_BWD_FRAMES + sub * 32 >= a.T) break;   // block-uniform
    const bool ok = t < a.T;
    const float* x = a.x + (int64_t)n * a.C * a.ldx + t;
    const float* gy = a.gy + (int64_t)n * a.C * a.ldg + t;
    float mean = 0.f, rstd = 0.f;
    if (ok) { mean = a.mean[(int64_t)n * a.T + t]; rstd = a.rstd[(int64_t)n * a.T + t]; }
    float s1 = 0.f, s2 = 0.f;
    if (ok)
      for (int c = ty; c < a.C; c += 8) {
        float g = __ldg(gy + (int64_t)c * a.ldg) * __ldg(a.gamma + c);
        float xh = (__ldg(x + (int64_t)c * a.ldx) - mean) * rstd;
        s1 += g;
        s2 = fmaf(g, xh, s2);
      }
    __syncthreads();
    red[0][ty][tx] = s1;
    red[1][ty][tx] = s2;
    __syncthreads();
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { m1 += red[0][k][tx]; m2 += red[1][k][tx]; }
    m1 /= (float)a.C;
    m2 /= (float)a.C;
    float* dx = a.dx + (int64_t)n * a.C * a.lddx + t;
    for (int i = 0; i < nch; ++i) {
      const int c = ty + 8 * i;
      float dg = 0.f, db = 0.f;
      if (ok && c < a.C) {
        const float gyv = __ldg(gy + (int64_t)c * a.ldg);
        const float xh = (__ldg(x + (int64_t)c * a.ldx) - mean) * rstd;
        dx[(int64_t)c * a.lddx] = rstd * (gyv * __ldg(a.gamma + c) - m1 - xh * m2);
        dg = gyv * xh;
        db = gyv;
      }
      dg = warp_sum(dg);   // warp == fixed ty: reduces over the 32 frames; channel c is owned by this warp only
      db = warp_sum(db);
      if (tx == 0 && c < a.C) {
        cacc[c] += dg;
        cacc[a.C + c] += db;
      }
    }
  }
  __syncthreads();
  for (int c = ty * 32 + tx; c < a.C; c += 256) {
    atomicAdd(a.dgamma + c, cacc[c]);
    atomicAdd(a.dbeta + c, cacc[a.C + c]);
  }
}

// ------------------------------------------------------------------------------------ framing
__global__ void frames_kernel(WesepFrameArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, n = blockIdx.z;
  if (k >= a.K) return;
  const int64_t s = (int64_t)k * a.hop + j;
  float v = s < a.S ? __ldg(a.x + (int64_t)n * a.ldx + s) : 0.f;
  a.F[((int64_t)n * a.J + j) * a.ldf + k] = v;
}

__global__ void overlap_add_kernel(WesepOlaArgs a) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (s >= a.S) return;
  const int j0 = (int)(s % a.hop);
  const int64_t k0 = s / a.hop;
  float acc = a.bias ? __ldg(a.bias) : 0.f;
  const float* F = a.F + (int64_t)n * a.J * a.ldf;
  for (int i = 0; j0 + i * a.hop < a.J; ++i) {
    const int64_t k = k0 - i;
    if (k < 0) break;
    if (k < a.K) acc += __ldg(F + (int64_t)(j0 + i * a.hop) * a.ldf + k);
  }
  a.y[(int64_t)n * a.ldy + s] = acc;
}

// ------------------------------------------------------------------------------------ elementwise
__global__ void mask_bwd_kernel(WesepMaskBwdArgs a) {
  const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const int64_t row = blockIdx.x;  // n*C rows
  if (t >= a.T) return;
  const int64_t o = row * a.ld + t;
  float4 gs = *reinterpret_cast<const float4*>(a.gS + o);
  float4 w = *reinterpret_cast<const float4*>(a.w + o);
  float4 m = *reinterpret_cast<const float4*>(a.m + o);
  float4 gw, gm;
  gw.x = gs.x * m.x; gw.y = gs.y * m.y; gw.z = gs.z * m.z; gw.w = gs.w * m.w;
  gm.x = m.x > 0.f ? gs.x * w.x : 0.f; gm.y = m.y > 0.f ? gs.y * w.y : 0.f;
  gm.z = m.z > 0.f ? gs.z * w.z : 0.f; gm.w = m.w > 0.f ? gs.w * w.w : 0.f;
  if (a.acc_w) {
    float4 o4 = *reinterpret_cast<const float4*>(a.gw + o);
    gw.x += o4.x; gw.y += o4.y; gw.z += o4.z; gw.w += o4.w;
  }
  *reinterpret_cast<float4*>(a.gw + o) = gw;   // columns in [T, ld) inside the last vector are padding
  *reinterpret_cast<float4*>(a.gm + o) = gm;
}

__global__ void relu_bwd_kernel(WesepReluBwdArgs a) {
  const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const int64_t row = blockIdx.x;
  if (t >= a.T) return;
  const int64_t o = row * a.ld + t;
  float4 y = *reinterpret_cast<const float4*>(a.y + o);
  float4 g = *reinterpret_cast<const float4*>(a.gy + o);
  g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f;
  g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
  *reinterpret_cast<float4*>(a.gx + o) = g;
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_cln_fwd(const WesepClnFwdArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "cln: empty shape");
  cln_fwd_kernel<<<dim3(cdiv(a->T, 32), a->n), dim3(32, 8), 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("cln_fwd");
  return 0;
}
extern "C" int wesep_b200_cln_bwd(const WesepClnBwdArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "cln: empty shape");
  cln_bwd_kernel<<<dim3(cdiv(a->T, CLN_BWD_FRAMES), a->n), dim3(32, 8), 2 * a->C * sizeof(float), (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("cln_bwd");
  return 0;
}
extern "C" int wesep_b200_frames(const WesepFrameArgs* a, void* stream) {
  if (a->n <= 0 || a->J <= 0 || a->K <= 0 || a->hop <= 0) return fail(-1, "frames: empty shape");
  frames_kernel<<<dim3(cdiv(a->K, 256), a->J, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("frames");
  return 0;
}
extern "C" int wesep_b200_overlap_add(const WesepOlaArgs* a, void* stream) {
  if (a->n <= 0 || a->J <= 0 || a->K <= 0 || a->hop <= 0 || a->S <= 0) return fail(-1, "overlap_add: empty shape");
  overlap_add_kernel<<<dim3(cdiv(a->S, 256), a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("overlap_add");
  return 0;
}
extern "C" int wesep_b200_mask_bwd(const WesepMaskBwdArgs* a, void* stream) {
  if ((a->ld & 3) || !aligned16(a->gS) || !aligned16(a->w) || !aligned16(a->m) || !aligned16(a->gw) || !aligned16(a->gm))
    return fail(-1, "mask_bwd: alignment");
  mask_bwd_kernel<<<dim3(a->n * a->C, cdiv(a->T, 1024)), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("mask_bwd");
  return 0;
}
extern "C" int wesep_b200_relu_bwd(const WesepReluBwdArgs* a, void* stream) {
  if ((a->ld & 3) || !aligned16(a->y) || !aligned16(a->gy) || !aligned16(a->gx)) return fail(-1, "relu_bwd: alignment");
  relu_bwd_kernel<<<dim3(a->n * a->C, cdiv(a->T, 1024)), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("relu_bwd");
  return 0;
}
