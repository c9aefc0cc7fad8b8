// Speaker-encoder pieces of Spex+ (ResNet4SpExplus / ResBlock, wesep/modules/tasnet/speaker.py:7-64) that are
// not GEMMs: BatchNorm1d (batch statistics) finalize, BN-apply + residual + PReLU + MaxPool1d(3) forward and the
// two-pass backward; plus pred_linear (nn.Linear 256 -> spksInTrain, convtasnet.py:115,194) and the
// cross-entropy loss (nn.CrossEntropyLoss, wesep/utils/losses.py:11).
#include "common.cuh"

namespace wb {

// scale[c] = w*rstd, shift[c] = b - w*mean*rstd from (sum, sumsq) over `count` elements (training) or from the
// running statistics (eval).  Training also updates running_mean / running_var (unbiased) like nn.BatchNorm1d.
__global__ void bn_finalize_kernel(WesepBnFinalizeArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  float mean, var;
  if (a.training) {
    const double m = a.ch_stats[2 * c] / a.count;
    double v = a.ch_stats[2 * c + 1] / a.count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    const double unbiased = a.count > 1.0 ? v * a.count / (a.count - 1.0) : v;
    a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
    a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
  } else {
    mean = a.running_mean[c];
    var = a.running_var[c];
  }
  const float rstd = rsqrtf(var + a.eps);
  const float w = a.weight[c], b = a.bias[c];
  a.scale[c] = w * rstd;
  a.shift[c] = b - w * mean * rstd;
  a.mean[c] = mean;
  a.rstd[c] = rstd;
}

// y[n][c][t'] = max_{j<pool} prelu(scale[c]*x[3t'+j] + shift[c] (+ res[3t'+j]); alpha);  pool in {1, 3}
__global__ void __launch_bounds__(256) bn_act_pool_fwd_kernel(WesepBnActPoolFwdArgs a) {
  const int Tp = a.pool == 3 ? a.T / 3 : a.T;
  const int tp = blockIdx.x * 256 + threadIdx.x;
  const int64_t row = blockIdx.y;  // n*C + c
  if (tp >= Tp) return;
  const int c = (int)(row % a.C);
  const float sc = __ldg(a.scale + c), sh = __ldg(a.shift + c), al = __ldg(a.alpha);
  const float* x = a.x + row * a.ldx;
  const float* r = a.res ? a.res + row * a.ldr : nullptr;
  float best = -INFINITY;
  for (int j = 0; j < a.pool; ++j) {
    const int t = a.pool * tp + j;
    float v = fmaf(sc, __ldg(x + t), sh);
    if (r) v += __ldg(r + t);
    const float s = prelu_f(v, al);
    best = (s > best || s != s) ? s : best;
  }
  a.y[row * a.ldy + tp] = best;
}

// Backward pass A: route gy through the max-pool (first maximum wins, like ATen) and the PReLU:
//   gv[t] = dL/d v[t],  v = scale*x + shift (+ res);  ch_sums[c] += (sum gv, sum gv*xhat),  dalpha += sum gs*v*[v<=0]
__global__ void __launch_bounds__(256) bn_act_pool_bwd_kernel(WesepBnActPoolBwdArgs a) {
  __shared__ float red[3 * 32];
  const int Tp = a.pool == 3 ? a.T / 3 : a.T;
  const int64_t row = blockIdx.y;
  const int c = (int)(row % a.C);
  const float sc = __ldg(a.scale + c), sh = __ldg(a.shift + c), al = __ldg(a.alpha);
  const float mean = __ldg(a.mean + c), rstd = __ldg(a.rstd + c);
  const float* x = a.x + row * a.ldx;
  const float* r = a.res ? a.res + row * a.ldr : nullptr;
  float* gv = a.gv + row * a.ldgv;
  float s_g = 0.f, s_gx = 0.f, s_al = 0.f;
  // one CTA sweeps a whole slice of the row (gridDim.x slices): the block reduction + fp64 atomics are paid once per
  // slice instead of once per 256 pooled frames
  const int per = (Tp + gridDim.x - 1) / gridDim.x;
  const int tp_end = min(Tp, (int)(blockIdx.x + 1) * per);
  for (int tp = blockIdx.x * per + threadIdx.x; tp < tp_end; tp += 256) {
    const float g = __ldg(a.gy + row * a.ldgy + tp);
    float v[3], xs[3], best = -INFINITY;
    int arg = 0;
    for (int j = 0; j < a.pool; ++j) {
      const int t = a.pool * tp + j;
      xs[j] = __ldg(x + t);
      v[j] = fmaf(sc, xs[j], sh);
      if (r) v[j] += __ldg(r + t);
      const float s = prelu_f(v[j], al);
      if (s > best || s != s) { best = s; arg = j; }
    }
    for (int j = 0; j < a.pool; ++j) {
      const int t = a.pool * tp + j;
      float out = 0.f;
      if (j == arg) {
        out = g * (v[j] > 0.f ? 1.f : al);
        s_al += v[j] > 0.f ? 0.f : g * v[j];
        s_g += out;
        s_gx = fmaf(out, (xs[j] - mean) * rstd, s_gx);
      }
      gv[t] = out;
    }
  }
  // frames beyond pool*Tp (T % 3 leftovers) receive no gradient
  if (a.pool == 3 && blockIdx.x == 0 && threadIdx.x < a.T - 3 * Tp) gv[3 * Tp + threadIdx.x] = 0.f;
  float vv[3] = {s_g, s_gx, s_al};
  block_sum<3>(vv, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.ch_sums + 2 * c, (double)vv[0]);
    atomicAdd(a.ch_sums + 2 * c + 1, (double)vv[1]);
    if (vv[2] != 0.f) atomicAdd(a.dalpha, vv[2]);
  }
}

// Backward pass B: BatchNorm backward.  training: dx = scale*(gv - m1 - xhat*m2), m1 = sum(gv)/count,
// m2 = sum(gv*xhat)/count;  eval: dx = scale*gv.  (dweight = sum gv*xhat, dbias = sum gv are read from ch_sums by the host.)
__global__ void __launch_bounds__(256) bn_bwd_kernel(WesepBnBwdArgs a) {
  const int t = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int64_t row = blockIdx.y;
  if (t >= a.T) return;
  const int c = (int)(row % a.C);
  const float sc = __ldg(a.scale + c), mean = __ldg(a.mean + c), rstd = __ldg(a.rstd + c);
  float m1 = 0.f, m2 = 0.f;
  if (a.training) {
    m1 = (float)(a.ch_sums[2 * c] / a.count);
    m2 = (float)(a.ch_sums[2 * c + 1] / a.count);
  }
  const float4 g = *reinterpret_cast<const float4*>(a.gv + row * a.ldgv + t);
  const float4 x = *reinterpret_cast<const float4*>(a.x + row * a.ldx + t);
  float4 o;
  o.x = sc * (g.x - m1 - (x.x - mean) * rstd * m2);
  o.y = sc * (g.y - m1 - (x.y - mean) * rstd * m2);
  o.z = sc * (g.z - m1 - (x.z - mean) * rstd * m2);
  o.w = sc * (g.w - m1 - (x.w - mean) * rstd * m2);
  *reinterpret_cast<float4*>(a.dx + row * a.lddx + t) = o;
}

// ------------------------------------------------------------------------------------ pred_linear + CE
// y[n][j] = sum_k W[j][k] x[n][k] + b[j]; one warp per output
__global__ void linear_fwd_kernel(WesepLinearArgs a) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= a.n * a.J) return;
  const int n = w / a.J, j = w % a.J;
  float s = 0.f;
  for (int k = lane; k < a.K; k += 32) s = fmaf(__ldg(a.W + (int64_t)j * a.K + k), __ldg(a.x + (int64_t)n * a.K + k), s);
  s = warp_sum(s);
  if (lane == 0) a.y[w] = s + (a.b ? __ldg(a.b + j) : 0.f);
}
// dW[j][k] = sum_n gy[n][j] x[n][k]; db[j] = sum_n gy[n][j]; dx[n][k] = sum_j gy[n][j] W[j][k]
__global__ void linear_bwd_kernel(WesepLinearArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.J * a.K) {
    const int j = i / a.K, k = i % a.K;
    float s = 0.f;
    for (int n = 0; n < a.n; ++n) s = fmaf(__ldg(a.gy + (int64_t)n * a.J + j), __ldg(a.x + (int64_t)n * a.K + k), s);
    a.dW[i] = s;
  }
  if (i < a.J && a.db) {
    float s = 0.f;
    for (int n = 0; n < a.n; ++n) s += __ldg(a.gy + (int64_t)n * a.J + i);
    a.db[i] = s;
  }
  if (i < a.n * a.K) {
    const int n = i / a.K, k = i % a.K;
    float s = 0.f;
    for (int j = 0; j < a.J; ++j) s = fmaf(__ldg(a.gy + (int64_t)n * a.J + j), __ldg(a.W + (int64_t)j * a.K + k), s);
    a.dx[i] = s;
  }
}
// loss = mean_n (logsumexp(z_n) - z_n[label_n]); dz[n][j] = (softmax_j - [j == label]) / n.  One warp per row.
__global__ void ce_kernel(WesepCeArgs a) {
  const int n = blockIdx.x, lane = threadIdx.x;
  const float* z = a.logits + (int64_t)n * a.J;
  float mx = -INFINITY;
  for (int j = lane; j < a.J; j += 32) mx = fmaxf(mx, __ldg(z + j));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int j = lane; j < a.J; j += 32) se += expf(__ldg(z + j) - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int lab = (int)a.labels[n];
  for (int j = lane; j < a.J; j += 32) {
    const float p = expf(__ldg(z + j) - lse);
    a.dlogits[(int64_t)n * a.J + j] = (p - (j == lab ? 1.f : 0.f)) / (float)a.n;
  }
  if (lane == 0) atomicAdd(a.loss, (lse - __ldg(z + lab)) / (float)a.n);
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_bn_finalize(const WesepBnFinalizeArgs* a, void* stream) {
  if (a->C <= 0) return fail(-1, "bn_finalize: empty");
  bn_finalize_kernel<<<cdiv(a->C, 128), 128, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn_finalize");
  return 0;
}
extern "C" int wesep_b200_bn_act_pool_fwd(const WesepBnActPoolFwdArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0 || !(a->pool == 1 || a->pool == 3)) return fail(-1, "bn_act_pool_fwd: bad shape");
  const int Tp = a->pool == 3 ? a->T / 3 : a->T;
  if (Tp <= 0) return fail(-1, "bn_act_pool_fwd: sequence shorter than the pooling window");
  if ((int64_t)a->n * a->C > 65535) return fail(-2, "bn_act_pool_fwd: n * C exceeds 65535 (gridDim.y): split the batch");
  bn_act_pool_fwd_kernel<<<dim3(cdiv(Tp, 256), a->n * a->C), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn_act_pool_fwd");
  return 0;
}
extern "C" int wesep_b200_bn_act_pool_bwd(const WesepBnActPoolBwdArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0 || !(a->pool == 1 || a->pool == 3)) return fail(-1, "bn_act_pool_bwd: bad shape");
  const int Tp = a->pool == 3 ? a->T / 3 : a->T;
  const int slices = cdiv(Tp, 2048);   // <= 8 passes of 256 pooled frames per CTA
  if ((int64_t)a->n * a->C > 65535) return fail(-2, "bn_act_pool_bwd: n * C exceeds 65535 (gridDim.y): split the batch");
  bn_act_pool_bwd_kernel<<<dim3(slices, a->n * a->C), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn_act_pool_bwd");
  return 0;
}
extern "C" int wesep_b200_bn_bwd(const WesepBnBwdArgs* a, void* stream) {
  if ((a->ldgv & 3) || (a->ldx & 3) || (a->lddx & 3) || !aligned16(a->gv) || !aligned16(a->x) || !aligned16(a->dx))
    return fail(-1, "bn_bwd: alignment");
  if ((int64_t)a->n * a->C > 65535) return fail(-2, "bn_bwd: n * C exceeds 65535 (gridDim.y): split the batch");
  bn_bwd_kernel<<<dim3(cdiv(a->T, 1024), a->n * a->C), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("bn_bwd");
  return 0;
}
extern "C" int wesep_b200_linear_fwd(const WesepLinearArgs* a, void* stream) {
  if (a->n <= 0 || a->J <= 0 || a->K <= 0) return fail(-1, "linear: empty");
  linear_fwd_kernel<<<cdiv((int64_t)a->n * a->J * 32, 256), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("linear_fwd");
  return 0;
}
extern "C" int wesep_b200_linear_bwd(const WesepLinearArgs* a, void* stream) {
  int work = a->J * a->K;
  if (a->n * a->K > work) work = a->n * a->K;
  linear_bwd_kernel<<<cdiv(work, 256), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("linear_bwd");
  return 0;
}
extern "C" int wesep_b200_cross_entropy(const WesepCeArgs* a, void* stream) {
  if (a->n <= 0 || a->J <= 0) return fail(-1, "cross_entropy: empty");
  WB_CUDA(cudaMemsetAsync(a->loss, 0, sizeof(float), (cudaStream_t)stream));
  ce_kernel<<<a->n, 32, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("cross_entropy");
  return 0;
}
