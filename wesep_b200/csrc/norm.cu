// GroupNorm(1, C) over (C, T) of every row of an act tensor [n][C][ld] — the norm in front of every ResRNN, every band
// split and every mask head of pBSRNN (wesep/models/bsrnn.py:23,39,204,274).  One CTA per row (the rows are independent):
// the forward makes two passes over the row (statistics, then apply — the second pass hits L2), the backward likewise
// (the two row sums of the gLN backward + per-channel dgamma / dbeta, then dx).  Replaces the one-CTA-per-(row, channel)
// gLN kernels of fuse.cu on this path: with T = 32 (band_comm) those launched a million 256-thread CTAs per call.
#include "common.cuh"

namespace wb {

constexpr int GN_THREADS = 512;
constexpr int GN_MAXC = 2048;

__device__ __forceinline__ void gn_block_sum2(double& a, double& b, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  if (warp == 0) {
    double x = lane < GN_THREADS / 32 ? red[lane] : 0.0, y = lane < GN_THREADS / 32 ? red[32 + lane] : 0.0;
    x = warp_sum(x);
    y = warp_sum(y);
    if (lane == 0) { red[0] = x; red[1] = y; }
  }
  __syncthreads();
  a = red[0];
  b = red[1];
  __syncthreads();
}

__global__ void __launch_bounds__(GN_THREADS) gn1_fwd_kernel(WesepGroupNorm1Args a) {
  __shared__ double red[64];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float* x = a.x + (int64_t)n * a.bsx;
  float* y = a.y + (int64_t)n * a.bsy;
  const int ld4 = (a.T + 3) >> 2, total = a.C * ld4;
  float s0 = 0.f, s1 = 0.f;
  for (int idx = tid; idx < total; idx += GN_THREADS) {
    const int c = idx / ld4, t = (idx - c * ld4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx + t));
    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.T) { s0 += xv[i]; s1 = fmaf(xv[i], xv[i], s1); }
  }
  double d0 = s0, d1 = s1;
  gn_block_sum2(d0, d1, red);
  if (tid == 0) { a.stats[2 * n] = d0; a.stats[2 * n + 1] = d1; }
  const double cnt = (double)a.C * a.T;
  const double m = d0 / cnt;
  double var = d1 / cnt - m * m;
  if (var < 0.0) var = 0.0;
  const float mu = (float)m, r = (float)(1.0 / sqrt(var + (double)a.eps));
  for (int idx = tid; idx < total; idx += GN_THREADS) {
    const int c = idx / ld4, t = (idx - c * ld4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx + t));
    const float gm = __ldg(a.gamma + c) * r, bt = __ldg(a.beta + c);
    const float sh = fmaf(-gm, mu, bt);
    float4 o;
    o.x = fmaf(gm, v.x, sh); o.y = t + 1 < a.T ? fmaf(gm, v.y, sh) : 0.f;
    o.z = t + 2 < a.T ? fmaf(gm, v.z, sh) : 0.f; o.w = t + 3 < a.T ? fmaf(gm, v.w, sh) : 0.f;
    *reinterpret_cast<float4*>(y + (int64_t)c * a.ldy + t) = o;
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn1_bwd_kernel(WesepGroupNorm1Args a) {
  __shared__ double red[64];
  __shared__ float chs[2 * GN_MAXC];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float* x = a.x + (int64_t)n * a.bsx;
  const float* gy = a.gy + (int64_t)n * a.bsg;
  float* dx = a.dx + (int64_t)n * a.bsdx;
  const int ld4 = (a.T + 3) >> 2, total = a.C * ld4;
  const double cnt = (double)a.C * a.T;
  const double m = a.stats[2 * n] / cnt;
  double var = a.stats[2 * n + 1] / cnt - m * m;
  if (var < 0.0) var = 0.0;
  const float mu = (float)m, r = (float)(1.0 / sqrt(var + (double)a.eps));
  for (int c = tid; c < 2 * a.C; c += GN_THREADS) chs[c] = 0.f;
  __syncthreads();
  float s0 = 0.f, s1 = 0.f;
  for (int idx = tid; idx < total; idx += GN_THREADS) {
    const int c = idx / ld4, t = (idx - c * ld4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx + t));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gy + (int64_t)c * a.ldg + t));
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float dg = 0.f, db = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.T) {
        const float yh = (xv[i] - mu) * r;
        dg = fmaf(gv[i], yh, dg);
        db += gv[i];
      }
    const float gm = __ldg(a.gamma + c);
    s0 = fmaf(gm, db, s0);
    s1 = fmaf(gm, dg, s1);
    atomicAdd(chs + c, dg);
    atomicAdd(chs + a.C + c, db);
  }
  double d0 = s0, d1 = s1;
  gn_block_sum2(d0, d1, red);            // (also orders the shared-memory atomics before the flush below)
  const float m1 = (float)(d0 / cnt), m2 = (float)(d1 / cnt);
  for (int idx = tid; idx < total; idx += GN_THREADS) {
    const int c = idx / ld4, t = (idx - c * ld4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * a.ldx + t));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gy + (int64_t)c * a.ldg + t));
    const float gm = __ldg(a.gamma + c);
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float yh = (xv[i] - mu) * r;
      o[i] = t + i < a.T ? r * (fmaf(gm, gv[i], -m1) - yh * m2) : 0.f;
    }
    *reinterpret_cast<float4*>(dx + (int64_t)c * a.lddx + t) = make_float4(o[0], o[1], o[2], o[3]);
  }
  for (int c = tid; c < a.C; c += GN_THREADS) {
    atomicAdd(a.dgamma + c, chs[c]);
    atomicAdd(a.dbeta + c, chs[a.C + c]);
  }
}

// ------------------------------------------------------------------------------------------------ wide rows
// Rows of millions of elements (TF-GridNet's GroupNorm(1, emb_dim) over (C, T, F) of an utterance, tfgridnet.py:174-176): one CTA
// per row would leave all but n SMs idle, so a (chunk, channel, row) grid accumulates the row sums with fp64 atomics.
constexpr int GNW_CHUNK = 8192;
constexpr int64_t GNW_MIN_ROW = 1 << 19;      // rows of at least this many elements take the wide path

__global__ void __launch_bounds__(256) gn1w_stats_kernel(WesepGroupNorm1Args a) {
  __shared__ double red[2][8];
  const int c = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  const float* x = a.x + (int64_t)n * a.bsx + (int64_t)c * a.ldx;
  const int c0 = blockIdx.x * GNW_CHUNK, end = min(c0 + GNW_CHUNK, a.T);
  double s0 = 0.0, s1 = 0.0;
  for (int t = c0 + tid; t < end; t += 256) {
    const double v = (double)__ldg(x + t);
    s0 += v; s1 = fma(v, v, s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s0; red[1][tid >> 5] = s1; }
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    atomicAdd(a.stats + (int64_t)n * 2 + tid, s);
  }
}
__device__ __forceinline__ void gnw_mean_rstd(const WesepGroupNorm1Args& a, int n, float& mu, float& rs) {
  const double cnt = (double)a.C * (double)a.T;
  const double mean = a.stats[(int64_t)n * 2] / cnt;
  const double var = fmax(a.stats[(int64_t)n * 2 + 1] / cnt - mean * mean, 0.0);
  mu = (float)mean;
  rs = (float)(1.0 / sqrt(var + (double)a.eps));
}
__global__ void __launch_bounds__(256) gn1w_apply_kernel(WesepGroupNorm1Args a) {
  const int c = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  float mu, rs;
  gnw_mean_rstd(a, n, mu, rs);
  const float g = __ldg(a.gamma + c) * rs, b = __ldg(a.beta + c);
  const float* x = a.x + (int64_t)n * a.bsx + (int64_t)c * a.ldx;
  float* y = a.y + (int64_t)n * a.bsy + (int64_t)c * a.ldy;
  const int c0 = blockIdx.x * GNW_CHUNK, end = min(c0 + GNW_CHUNK, a.T);
  for (int t = c0 + tid; t < end; t += 256) y[t] = fmaf(__ldg(x + t) - mu, g, b);
}
__global__ void __launch_bounds__(256) gn1w_bwd_reduce_kernel(WesepGroupNorm1Args a) {
  __shared__ double red[2][8];
  const int c = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  float mu, rs;
  gnw_mean_rstd(a, n, mu, rs);
  const float* x = a.x + (int64_t)n * a.bsx + (int64_t)c * a.ldx;
  const float* gy = a.gy + (int64_t)n * a.bsg + (int64_t)c * a.ldg;
  const int c0 = blockIdx.x * GNW_CHUNK, end = min(c0 + GNW_CHUNK, a.T);
  double s0 = 0.0, s1 = 0.0;
  for (int t = c0 + tid; t < end; t += 256) {
    const float g = __ldg(gy + t);
    s0 += (double)g;
    s1 = fma((double)g, (double)((__ldg(x + t) - mu) * rs), s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s0; red[1][tid >> 5] = s1; }
  __syncthreads();
  if (tid == 0) {
    double v0 = 0.0, v1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { v0 += red[0][w]; v1 += red[1][w]; }
    atomicAdd(a.dbeta + c, (float)v0);
    atomicAdd(a.dgamma + c, (float)v1);
    const double gm = (double)__ldg(a.gamma + c);
    atomicAdd(a.bsum + (int64_t)n * 2, gm * v0);
    atomicAdd(a.bsum + (int64_t)n * 2 + 1, gm * v1);
  }
}
__global__ void __launch_bounds__(256) gn1w_bwd_apply_kernel(WesepGroupNorm1Args a) {
  const int c = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  float mu, rs;
  gnw_mean_rstd(a, n, mu, rs);
  const double cnt = (double)a.C * (double)a.T;
  const float m0 = (float)(a.bsum[(int64_t)n * 2] / cnt), m1 = (float)(a.bsum[(int64_t)n * 2 + 1] / cnt);
  const float gm = __ldg(a.gamma + c);
  const float* x = a.x + (int64_t)n * a.bsx + (int64_t)c * a.ldx;
  const float* gy = a.gy + (int64_t)n * a.bsg + (int64_t)c * a.ldg;
  float* dx = a.dx + (int64_t)n * a.bsdx + (int64_t)c * a.lddx;
  const int c0 = blockIdx.x * GNW_CHUNK, end = min(c0 + GNW_CHUNK, a.T);
  for (int t = c0 + tid; t < end; t += 256) {
    const float xh = (__ldg(x + t) - mu) * rs;
    dx[t] = rs * (__ldg(gy + t) * gm - m0 - xh * m1);
  }
}

}  // namespace wb

using namespace wb;

static int check_gn(const WesepGroupNorm1Args* a, bool bwd) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "groupnorm1: empty shape");
  if (a->C > GN_MAXC) return fail(-2, "groupnorm1: more than 2048 channels");
  if ((a->ldx & 3) || (a->bsx & 3) || !aligned16(a->x) || !a->gamma || !a->beta || !a->stats) return fail(-1, "groupnorm1: x / parameters");
  if (a->ldx < ((a->T + 3) & ~3)) return fail(-1, "groupnorm1: row stride shorter than the padded length");
  if (!bwd && ((a->ldy & 3) || (a->bsy & 3) || !aligned16(a->y))) return fail(-1, "groupnorm1: y");
  if (bwd && ((a->ldg & 3) || (a->bsg & 3) || (a->lddx & 3) || (a->bsdx & 3) || !aligned16(a->gy) || !aligned16(a->dx) || !a->dgamma ||
              !a->dbeta))
    return fail(-1, "groupnorm1_bwd: gradient buffers");
  return 0;
}
extern "C" int wesep_b200_groupnorm1_fwd(const WesepGroupNorm1Args* a, void* stream) {
  if (int rc = check_gn(a, false)) return rc;
  if ((int64_t)a->C * a->T >= GNW_MIN_ROW && a->n <= 65535) {
    cudaStream_t st = (cudaStream_t)stream;
    WB_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)a->n * 2 * sizeof(double), st));
    const dim3 grid(cdiv(a->T, GNW_CHUNK), a->C, a->n);
    gn1w_stats_kernel<<<grid, 256, 0, st>>>(*a);
    WB_LAUNCH_CHECK("groupnorm1_wide_stats");
    gn1w_apply_kernel<<<grid, 256, 0, st>>>(*a);
    WB_LAUNCH_CHECK("groupnorm1_wide_apply");
    return 0;
  }
  gn1_fwd_kernel<<<a->n, GN_THREADS, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("groupnorm1_fwd");
  return 0;
}
extern "C" int wesep_b200_groupnorm1_bwd(const WesepGroupNorm1Args* a, void* stream) {
  if (int rc = check_gn(a, true)) return rc;
  if ((int64_t)a->C * a->T >= GNW_MIN_ROW && a->n <= 65535 && a->bsum) {
    cudaStream_t st = (cudaStream_t)stream;
    WB_CUDA(cudaMemsetAsync(a->bsum, 0, (size_t)a->n * 2 * sizeof(double), st));
    const dim3 grid(cdiv(a->T, GNW_CHUNK), a->C, a->n);
    gn1w_bwd_reduce_kernel<<<grid, 256, 0, st>>>(*a);
    WB_LAUNCH_CHECK("groupnorm1_wide_bwd_reduce");
    gn1w_bwd_apply_kernel<<<grid, 256, 0, st>>>(*a);
    WB_LAUNCH_CHECK("groupnorm1_wide_bwd_apply");
    return 0;
  }
  gn1_bwd_kernel<<<a->n, GN_THREADS, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("groupnorm1_bwd");
  return 0;
}
