// Building blocks of the wespeaker ResNet speaker encoder that pBSRNN trains jointly (wesep/models/bsrnn.py:217,352-356;
// examples/librimix/tse/v2/confs/bsrnn.yaml:56-64: ResNet34, feat_dim 80, embed_dim 256, TSTP):
//   * conv3x3 (pad 1, stride 1 / 2) = im2col (this file) + the tcgen05 pointwise GEMM over the 9 C gathered channels;
//     the adjoint of im2col is a gather too (no atomics);
//   * BatchNorm2d with batch statistics, fused with the residual add and the ReLU of a BasicBlock;
//   * TSTP pooling (mean and unbiased std over time).
// Feature maps are act tensors [n][C][H * W] (time = W contiguous), so BatchNorm2d over (n, H, W) is a per-channel reduction
// over (n, row) and every kernel here is a coalesced HBM-bound sweep.
#include "common.cuh"

namespace wb {

// ------------------------------------------------------------------------------------------------ im2col 3x3, pad 1
// col[n][(c * 9 + kh * 3 + kw)][ho * Wo + wo] = x[n][c][(ho * s + kh - 1) * W + (wo * s + kw - 1)]  (0 outside)
__global__ void __launch_bounds__(256) im2col3x3_kernel(WesepIm2colArgs a) {
  const int r = blockIdx.y;                       // row of col: c * 9 + kh * 3 + kw
  const int n = blockIdx.z;
  const int c = r / 9, kh = (r % 9) / 3, kw = r % 3;
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ldx;
  float* col = a.col + (int64_t)n * a.bsc + (int64_t)r * a.ldc;
  const int HWo = a.Ho * a.Wo;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HWo; i += gridDim.x * 256) {
    const int ho = i / a.Wo, wo = i - ho * a.Wo;
    const int h = ho * a.stride + kh - 1, w = wo * (a.stride_w ? a.stride_w : a.stride) + kw - 1;
    col[i] = (h >= 0 && h < a.H && w >= 0 && w < a.W) ? __ldg(x + (int64_t)h * a.W + w) : 0.f;
  }
}
// adjoint: gx[n][c][h * W + w] = sum over (kh, kw) with (h + 1 - kh) % s == 0, (w + 1 - kw) % s == 0 of
//          gcol[n][(c, kh, kw)][((h + 1 - kh) / s) * Wo + (w + 1 - kw) / s]
__global__ void __launch_bounds__(256) col2im3x3_kernel(WesepIm2colArgs a) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float* gcol = a.gcol + (int64_t)n * a.bsc + (int64_t)(9 * c) * a.ldc;
  float* gx = a.gx + ((int64_t)n * a.C + c) * a.ldx;
  const int HW = a.H * a.W;
  const int sw = a.stride_w ? a.stride_w : a.stride;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
    const int h = i / a.W, w = i - h * a.W;
    float s = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hn = h + 1 - kh;
      if (hn < 0 || hn % a.stride) continue;
      const int ho = hn / a.stride;
      if (ho >= a.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wn = w + 1 - kw;
        if (wn < 0 || wn % sw) continue;
        const int wo = wn / sw;
        if (wo >= a.Wo) continue;
        s += __ldg(gcol + (int64_t)(kh * 3 + kw) * a.ldc + (int64_t)ho * a.Wo + wo);
      }
    }
    gx[i] = s;
  }
}
// strided subsampling for the 1x1 stride-2 shortcut: y[n][c][ho * Wo + wo] = x[n][c][(ho * s) * W + wo * s]; adjoint scatters
template <bool BWD>
__global__ void __launch_bounds__(256) subsample_kernel(WesepIm2colArgs a) {
  const int c = blockIdx.y, n = blockIdx.z;
  const int HWo = a.Ho * a.Wo;
  if constexpr (!BWD) {
    const float* x = a.x + ((int64_t)n * a.C + c) * a.ldx;
    float* y = a.col + (int64_t)n * a.bsc + (int64_t)c * a.ldc;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HWo; i += gridDim.x * 256) {
      const int ho = i / a.Wo, wo = i - ho * a.Wo;
      y[i] = __ldg(x + (int64_t)ho * a.stride * a.W + wo * a.stride);
    }
  } else {
    const float* gy = a.gcol + (int64_t)n * a.bsc + (int64_t)c * a.ldc;
    float* gx = a.gx + ((int64_t)n * a.C + c) * a.ldx;
    const int HW = a.H * a.W;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
      const int h = i / a.W, w = i - h * a.W;
      const bool hit = (h % a.stride == 0) && (w % a.stride == 0) && (h / a.stride < a.Ho) && (w / a.stride < a.Wo);
      gx[i] = hit ? __ldg(gy + (int64_t)(h / a.stride) * a.Wo + w / a.stride) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm (+ add + ReLU)
// stats[c] = (sum, sum of squares) over (n, t) in fp64 (zeroed by the caller)
__global__ void __launch_bounds__(256) bn2_stats_kernel(WesepBn2Args a) {
  __shared__ float red[2 * 32];
  const int c = blockIdx.x, n = blockIdx.y;
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ld;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 4 * threadIdx.x; t < a.T; t += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + t));
    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.T) { s0 += xv[i]; s1 = fmaf(xv[i], xv[i], s1); }
  }
  float v[2] = {s0, s1};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.stats + 2 * c, (double)v[0]);
    atomicAdd(a.stats + 2 * c + 1, (double)v[1]);
  }
}
// y = act(scale[c] * x + shift[c] (+ res)), act = ReLU if a.relu
__global__ void __launch_bounds__(256) bn2_apply_kernel(WesepBn2Args a) {
  const int c = blockIdx.x, n = blockIdx.y;
  const int64_t row = ((int64_t)n * a.C + c) * a.ld;
  const float sc = __ldg(a.scale + c), sh = __ldg(a.shift + c);
  for (int t = 4 * threadIdx.x; t < a.T; t += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x + row + t));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.res) r = __ldg(reinterpret_cast<const float4*>(a.res + row + t));
    float o[4] = {fmaf(sc, v.x, sh) + r.x, fmaf(sc, v.y, sh) + r.y, fmaf(sc, v.z, sh) + r.z, fmaf(sc, v.w, sh) + r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (a.relu) o[i] = fmaxf(o[i], 0.f);
      if (t + i >= a.T) o[i] = 0.f;
    }
    *reinterpret_cast<float4*>(a.y + row + t) = make_float4(o[0], o[1], o[2], o[3]);
  }
}
// backward pass 1: g' = gy * (y > 0) (ReLU) ; bsum[c] += (sum g', sum g' * xhat), xhat = (x - mean[c]) * rstd[c]
__global__ void __launch_bounds__(256) bn2_bwd_reduce_kernel(WesepBn2Args a) {
  __shared__ float red[2 * 32];
  const int c = blockIdx.x, n = blockIdx.y;
  const int64_t row = ((int64_t)n * a.C + c) * a.ld;
  const float mu = __ldg(a.mean + c), r = __ldg(a.rstd + c);
  float s0 = 0.f, s1 = 0.f;
  for (int t = 4 * threadIdx.x; t < a.T; t += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x + row + t));
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gy + row + t));
    float4 y = make_float4(1.f, 1.f, 1.f, 1.f);
    if (a.relu) y = __ldg(reinterpret_cast<const float4*>(a.y + row + t));
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g.x, g.y, g.z, g.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < a.T && yv[i] > 0.f) { s0 += gv[i]; s1 = fmaf(gv[i], (xv[i] - mu) * r, s1); }
  }
  float v[2] = {s0, s1};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.bsum + 2 * c, (double)v[0]);
    atomicAdd(a.bsum + 2 * c + 1, (double)v[1]);
  }
}
// backward pass 2: gx = gamma[c] rstd[c] (g' - m0[c] - xhat m1[c]), m = bsum / count ; gres = g' (if requested)
__global__ void __launch_bounds__(256) bn2_bwd_apply_kernel(WesepBn2Args a) {
  const int c = blockIdx.x, n = blockIdx.y;
  const int64_t row = ((int64_t)n * a.C + c) * a.ld;
  const float mu = __ldg(a.mean + c), r = __ldg(a.rstd + c), gr = __ldg(a.gamma + c) * r;
  const float m0 = (float)(a.bsum[2 * c] / a.count), m1 = (float)(a.bsum[2 * c + 1] / a.count);
  for (int t = 4 * threadIdx.x; t < a.T; t += 1024) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x + row + t));
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gy + row + t));
    float4 y = make_float4(1.f, 1.f, 1.f, 1.f);
    if (a.relu) y = __ldg(reinterpret_cast<const float4*>(a.y + row + t));
    const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g.x, g.y, g.z, g.w}, yv[4] = {y.x, y.y, y.z, y.w};
    float o[4], gp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = t + i < a.T;
      gp[i] = (ok && yv[i] > 0.f) ? gv[i] : 0.f;
      o[i] = ok ? gr * (gp[i] - m0 - (xv[i] - mu) * r * m1) : 0.f;
    }
    *reinterpret_cast<float4*>(a.gx + row + t) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.gres) *reinterpret_cast<float4*>(a.gres + row + t) = make_float4(gp[0], gp[1], gp[2], gp[3]);
  }
}

// ------------------------------------------------------------------------------------------------ TSTP pooling
// x [n][R][ld] (T valid) -> out [n][2R] = (mean over time | sqrt(unbiased var + 1e-7)); one warp per row
template <bool BWD>
__global__ void __launch_bounds__(256) tstp_kernel(WesepTstpArgs a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.n * a.R) return;
  const int n = row / a.R, r = row - n * a.R;
  const float* x = a.x + (int64_t)row * a.ld;
  float s0 = 0.f;
  for (int t = lane; t < a.T; t += 32) s0 += __ldg(x + t);
  s0 = warp_sum(s0);
  const float mean = s0 / (float)a.T;
  float s1 = 0.f;
  for (int t = lane; t < a.T; t += 32) { const float d = __ldg(x + t) - mean; s1 = fmaf(d, d, s1); }
  s1 = warp_sum(s1);
  const float var = s1 / (float)(a.T - 1);
  const float sd = sqrtf(var + (a.eps > 0.f ? a.eps : 1e-7f));
  if constexpr (!BWD) {
    if (lane == 0) {
      a.out[(int64_t)n * 2 * a.R + r] = mean;
      a.out[(int64_t)n * 2 * a.R + a.R + r] = sd;
    }
  } else {
    const float gm = __ldg(a.gout + (int64_t)n * 2 * a.R + r) / (float)a.T;
    const float gs = __ldg(a.gout + (int64_t)n * 2 * a.R + a.R + r) / ((float)(a.T - 1) * sd);
    float* gx = a.gx + (int64_t)row * a.ld;
    for (int t = lane; t < a.T; t += 32) gx[t] = fmaf(gs, __ldg(x + t) - mean, gm);
  }
}

}  // namespace wb

using namespace wb;

static int check_i2c(const WesepIm2colArgs* a) {
  if (a->n <= 0 || a->C <= 0 || a->H <= 0 || a->W <= 0 || (a->stride != 1 && a->stride != 2)) return fail(-1, "im2col: shape / stride");
  if (a->stride_w != 0 && a->stride_w != 1 && a->stride_w != 2) return fail(-1, "im2col: stride_w");
  const int sw_ = a->stride_w ? a->stride_w : a->stride;
  if (a->Ho != (a->H - 1) / a->stride + 1 || a->Wo != (a->W - 1) / sw_ + 1) return fail(-1, "im2col: output size (pad 1, k 3)");
  if (a->ldx < (int64_t)a->H * a->W || a->ldc < (int64_t)a->Ho * a->Wo || a->bsc < a->ldc) return fail(-1, "im2col: row / batch strides");
  if ((int64_t)a->n > 65535 || 9 * (int64_t)a->C > 65535) return fail(-2, "im2col: too many rows / channels for one launch");
  return 0;
}
extern "C" int wesep_b200_im2col3x3_fwd(const WesepIm2colArgs* a, void* stream) {
  if (int rc = check_i2c(a)) return rc;
  if (!a->x || !a->col) return fail(-1, "im2col: pointers");
  im2col3x3_kernel<<<dim3(cdiv((int64_t)a->Ho * a->Wo, 1024), 9 * a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("im2col3x3_fwd");
  return 0;
}
extern "C" int wesep_b200_im2col3x3_bwd(const WesepIm2colArgs* a, void* stream) {
  if (int rc = check_i2c(a)) return rc;
  if (!a->gcol || !a->gx) return fail(-1, "col2im: pointers");
  col2im3x3_kernel<<<dim3(cdiv((int64_t)a->H * a->W, 1024), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("im2col3x3_bwd");
  return 0;
}
extern "C" int wesep_b200_subsample2d_fwd(const WesepIm2colArgs* a, void* stream) {
  if (int rc = check_i2c(a)) return rc;
  if (!a->x || !a->col) return fail(-1, "subsample: pointers");
  subsample_kernel<false><<<dim3(cdiv((int64_t)a->Ho * a->Wo, 1024), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("subsample2d_fwd");
  return 0;
}
extern "C" int wesep_b200_subsample2d_bwd(const WesepIm2colArgs* a, void* stream) {
  if (int rc = check_i2c(a)) return rc;
  if (!a->gcol || !a->gx) return fail(-1, "subsample_bwd: pointers");
  subsample_kernel<true><<<dim3(cdiv((int64_t)a->H * a->W, 1024), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("subsample2d_bwd");
  return 0;
}

static int check_bn2(const WesepBn2Args* a) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "bn2: empty shape");
  if ((a->ld & 3) || a->ld < a->T || !aligned16(a->x)) return fail(-1, "bn2: row stride / alignment");
  if (a->n > 65535) return fail(-2, "bn2: more than 65535 rows per launch");
  return 0;
}
extern "C" int wesep_b200_bn2_stats(const WesepBn2Args* a, void* stream) {
  if (int rc = check_bn2(a)) return rc;
  if (!a->stats) return fail(-1, "bn2_stats: stats buffer");
  bn2_stats_kernel<<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn2_stats");
  return 0;
}
extern "C" int wesep_b200_bn2_apply(const WesepBn2Args* a, void* stream) {
  if (int rc = check_bn2(a)) return rc;
  if (!a->scale || !a->shift || !a->y || !aligned16(a->y) || (a->res && !aligned16(a->res))) return fail(-1, "bn2_apply: pointers");
  bn2_apply_kernel<<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn2_apply");
  return 0;
}
extern "C" int wesep_b200_bn2_bwd(const WesepBn2Args* a, void* stream) {
  if (int rc = check_bn2(a)) return rc;
  if (!a->gy || !a->gx || !a->mean || !a->rstd || !a->gamma || !a->bsum || !aligned16(a->gy) || !aligned16(a->gx) ||
      (a->relu && (!a->y || !aligned16(a->y))) || (a->gres && !aligned16(a->gres)) || a->count <= 0.0)
    return fail(-1, "bn2_bwd: pointers");
  bn2_bwd_reduce_kernel<<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn2_bwd_reduce");
  bn2_bwd_apply_kernel<<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn2_bwd_apply");
  return 0;
}

extern "C" int wesep_b200_tstp_fwd(const WesepTstpArgs* a, void* stream) {
  if (a->n <= 0 || a->R <= 0 || a->T <= 1 || a->ld < a->T || !a->x || !a->out) return fail(-1, "tstp: shape / pointers (T >= 2)");
  tstp_kernel<false><<<cdiv((int64_t)a->n * a->R, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("tstp_fwd");
  return 0;
}
extern "C" int wesep_b200_tstp_bwd(const WesepTstpArgs* a, void* stream) {
  if (a->n <= 0 || a->R <= 0 || a->T <= 1 || a->ld < a->T || !a->x || !a->gout || !a->gx) return fail(-1, "tstp_bwd: shape / pointers");
  tstp_kernel<true><<<cdiv((int64_t)a->n * a->R, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("tstp_bwd");
  return 0;
}
