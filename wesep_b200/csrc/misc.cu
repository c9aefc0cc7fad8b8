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

// ------------------------------------------------------------------------------------ cLN, vectorised
// 1024 threads = 16 frame-quads (float4 along time) x 64 channel groups; one sub-tile = 64 frames x C channels.
// With one CTA per SM the 148 live sub-tiles (C = 768: 2 x 196 KB each) stay L2-resident between the statistics
// pass and the apply pass, so DRAM sees each operand once.  Needs ld % 4 == 0 and 16-byte aligned bases.
constexpr int CLV_FR = 64, CLV_SUBS = 4;

// sum of a float4 pair over the 64 channel groups of one frame-quad; result valid in every thread
__device__ __forceinline__ void clv_reduce2(float4& u, float4& v, float4 (*red)[32][16], float4 (*red2)[8][16], int tid) {
  const int tx = tid & 15, warp = tid >> 5, lane = tid & 31;
  u.x += __shfl_xor_sync(0xffffffffu, u.x, 16); u.y += __shfl_xor_sync(0xffffffffu, u.y, 16);
  u.z += __shfl_xor_sync(0xffffffffu, u.z, 16); u.w += __shfl_xor_sync(0xffffffffu, u.w, 16);
  v.x += __shfl_xor_sync(0xffffffffu, v.x, 16); v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
  v.z += __shfl_xor_sync(0xffffffffu, v.z, 16); v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
  __syncthreads();   // previous users of red / red2 are done
  if (lane < 16) { red[0][warp][tx] = u; red[1][warp][tx] = v; }
  __syncthreads();
  if (tid < 256) {   // (which, group of 4 warps, tx)
    const int k = tid >> 7, g = (tid >> 4) & 7;
    float4 acc = red[k][4 * g][tx];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const float4 r = red[k][4 * g + w][tx];
      acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    red2[k][g][tx] = acc;
  }
  __syncthreads();
  u = red2[0][0][tx]; v = red2[1][0][tx];
#pragma unroll
  for (int g = 1; g < 8; ++g) {
    const float4 r0 = red2[0][g][tx], r1 = red2[1][g][tx];
    u.x += r0.x; u.y += r0.y; u.z += r0.z; u.w += r0.w;
    v.x += r1.x; v.y += r1.y; v.z += r1.z; v.w += r1.w;
  }
}

__global__ void __launch_bounds__(1024, 1) cln_fwd_v4_kernel(WesepClnFwdArgs a) {
  __shared__ float4 red[2][32][16];
  __shared__ float4 red2[2][8][16];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, n = blockIdx.y;
  const float invC = 1.f / (float)a.C;
  for (int sub = 0; sub < CLV_SUBS; ++sub) {
    const int t0 = (blockIdx.x * CLV_SUBS + sub) * CLV_FR;
    if (t0 >= a.T) break;   // block-uniform
    const int t = t0 + 4 * tx;
    const bool act = t < a.T;
    const float* x = a.x + (int64_t)n * a.C * a.ldx + t;
    // shifted one-pass statistics: sums of (x - x[channel 0]) and its square
    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) sh = __ldg(reinterpret_cast<const float4*>(x));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (act) {
#pragma unroll 4
      for (int c = ty; c < a.C; c += 64) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx));
        const float d0 = v.x - sh.x, d1 = v.y - sh.y, d2 = v.z - sh.z, d3 = v.w - sh.w;
        s.x += d0; s.y += d1; s.z += d2; s.w += d3;
        q.x = fmaf(d0, d0, q.x); q.y = fmaf(d1, d1, q.y); q.z = fmaf(d2, d2, q.z); q.w = fmaf(d3, d3, q.w);
      }
    }
    clv_reduce2(s, q, red, red2, tid);
    float mean[4], rstd[4];
    {
      const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w}, hv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float m = sv[i] * invC;
        mean[i] = hv[i] + m;
        rstd[i] = rsqrtf(fmaxf(qv[i] * invC - m * m, 0.f) + a.eps);
      }
    }
    if (act) {
      float* y = a.y + (int64_t)n * a.C * a.ldy + t;
#pragma unroll 4
      for (int c = ty; c < a.C; c += 64) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx));
        const float gm = __ldg(a.gamma + c), bt = __ldg(a.beta + c);
        float4 o;
        o.x = fmaf((v.x - mean[0]) * rstd[0], gm, bt); o.y = fmaf((v.y - mean[1]) * rstd[1], gm, bt);
        o.z = fmaf((v.z - mean[2]) * rstd[2], gm, bt); o.w = fmaf((v.w - mean[3]) * rstd[3], gm, bt);
        *reinterpret_cast<float4*>(y + (int64_t)c * a.ldy) = o;
      }
      if (ty == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (t + i < a.T) {
            a.mean[(int64_t)n * a.T + t + i] = mean[i];
            a.rstd[(int64_t)n * a.T + t + i] = rstd[i];
          }
      }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) cln_bwd_v4_kernel(WesepClnBwdArgs a) {
  extern __shared__ float cacc[];  // [2][C]
  __shared__ float4 red[2][32][16];
  __shared__ float4 red2[2][8][16];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, n = blockIdx.y;
  const float invC = 1.f / (float)a.C;
  for (int i = tid; i < 2 * a.C; i += 1024) cacc[i] = 0.f;
  __syncthreads();
  for (int sub = 0; sub < CLV_SUBS; ++sub) {
    const int t0 = (blockIdx.x * CLV_SUBS + sub) * CLV_FR;
    if (t0 >= a.T) break;   // block-uniform
    const int t = t0 + 4 * tx;
    const bool act = t < a.T;
    const float* x = a.x + (int64_t)n * a.C * a.ldx + t;
    const float* gy = a.gy + (int64_t)n * a.C * a.ldg + t;
    float mean[4], rstd[4];
    bool okf[4];   // frames >= T (pad columns hold arbitrary bits, possibly NaN): operands are REPLACED by 0, never scaled
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = t + i < a.T;
      mean[i] = ok ? __ldg(a.mean + (int64_t)n * a.T + t + i) : 0.f;
      rstd[i] = ok ? __ldg(a.rstd + (int64_t)n * a.T + t + i) : 0.f;
      okf[i] = ok;
    }
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (act) {
#pragma unroll 4
      for (int c = ty; c < a.C; c += 64) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gy + (int64_t)c * a.ldg));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx));
        const float gm = __ldg(a.gamma + c);
        const float g0 = okf[0] ? g.x * gm : 0.f, g1 = okf[1] ? g.y * gm : 0.f, g2 = okf[2] ? g.z * gm : 0.f,
                    g3 = okf[3] ? g.w * gm : 0.f;
        const float x0 = okf[0] ? v.x : 0.f, x1 = okf[1] ? v.y : 0.f, x2 = okf[2] ? v.z : 0.f, x3 = okf[3] ? v.w : 0.f;
        s1.x += g0; s1.y += g1; s1.z += g2; s1.w += g3;
        s2.x = fmaf(g0, (x0 - mean[0]) * rstd[0], s2.x); s2.y = fmaf(g1, (x1 - mean[1]) * rstd[1], s2.y);
        s2.z = fmaf(g2, (x2 - mean[2]) * rstd[2], s2.z); s2.w = fmaf(g3, (x3 - mean[3]) * rstd[3], s2.w);
      }
    }
    clv_reduce2(s1, s2, red, red2, tid);
    const float m1[4] = {s1.x * invC, s1.y * invC, s1.z * invC, s1.w * invC};
    const float m2[4] = {s2.x * invC, s2.y * invC, s2.z * invC, s2.w * invC};
    float* dx = a.dx + (int64_t)n * a.C * a.lddx + t;
#pragma unroll 2
    for (int c = ty; c < a.C + ty; c += 64) {   // same trip count for every thread (shuffles below); c >= C is idle
      const bool cok = c < a.C;
      float dg = 0.f, db = 0.f;
      if (act && cok) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gy + (int64_t)c * a.ldg));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx));
        const float gm = __ldg(a.gamma + c);
        const float gv[4] = {okf[0] ? g.x : 0.f, okf[1] ? g.y : 0.f, okf[2] ? g.z : 0.f, okf[3] ? g.w : 0.f};
        const float xv[4] = {okf[0] ? v.x : 0.f, okf[1] ? v.y : 0.f, okf[2] ? v.z : 0.f, okf[3] ? v.w : 0.f};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xh = (xv[i] - mean[i]) * rstd[i];
          o[i] = rstd[i] * (gv[i] * gm - m1[i] - xh * m2[i]);
          dg = fmaf(gv[i], xh, dg);
          db += gv[i];
        }
        *reinterpret_cast<float4*>(dx + (int64_t)c * a.lddx) = make_float4(o[0], o[1], o[2], o[3]);
      }
#pragma unroll
      for (int o_ = 8; o_ > 0; o_ >>= 1) {   // over the 16 frame-quads of this channel group (half a warp)
        dg += __shfl_xor_sync(0xffffffffu, dg, o_);
        db += __shfl_xor_sync(0xffffffffu, db, o_);
      }
      if (tx == 0 && cok) {   // channel c is owned by channel group ty = c % 64 only
        cacc[c] += dg;
        cacc[a.C + c] += db;
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < a.C; c += 1024) {
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
  const bool v4 = !((a->ldx | a->ldy) & 3) && aligned16(a->x) && aligned16(a->y);
  if (v4)
    cln_fwd_v4_kernel<<<dim3(cdiv(a->T, CLV_FR * CLV_SUBS), a->n), 1024, 0, (cudaStream_t)stream>>>(*a);
  else
    cln_fwd_kernel<<<dim3(cdiv(a->T, 32), a->n), dim3(32, 8), 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("cln_fwd");
  return 0;
}
extern "C" int wesep_b200_cln_bwd(const WesepClnBwdArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "cln: empty shape");
  const bool v4 = !((a->ldx | a->ldg | a->lddx) & 3) && aligned16(a->x) && aligned16(a->gy) && aligned16(a->dx);
  if (v4)
    cln_bwd_v4_kernel<<<dim3(cdiv(a->T, CLV_FR * CLV_SUBS), a->n), 1024, 2 * a->C * sizeof(float), (cudaStream_t)stream>>>(*a);
  else
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
