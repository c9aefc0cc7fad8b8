// pDPCCN building blocks (wesep/models/dpccn.py:206-290, wesep/modules/dpccn/convs.py:28-152).
// The 3x3 (transposed) convolutions run as im2col / col2im (resnet.cu) around the tcgen05 pointwise GEMM; this file
// holds the memory-bound rest, all on act tensors [n][C][T * F] (F contiguous, as the reference's NCHW maps):
//   * ELU + InstanceNorm (Conv2dBlock / ConvTrans2dBlock: norm(elu(x)); TCNBlock: elu(norm(x)))
//   * depthwise dilated Conv1d k = 3 of the TCN blocks
//   * AvgPool2d / bilinear Upsample of the pyramidal pooling tail
//   * the 4-D "multiply" speaker fusion (common/speaker.py:117-121): a per-(row, frequency) gain
#include "common.cuh"

namespace wb {

constexpr int EN_CHUNK = 8192;

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float elu_d(float x) { return x > 0.f ? 1.f : __expf(x); }

// ------------------------------------------------------------------------------------------------ ELU + InstanceNorm
// stats[row] = (sum v, sum v^2), v = elu(x) (mode 0) or x (mode 1); fp64, zeroed by the entry point
__global__ void __launch_bounds__(256) eluin_stats_kernel(WesepEluInArgs a) {
  __shared__ double red[2][8];
  const int64_t row = blockIdx.y;
  const int c0 = blockIdx.x * EN_CHUNK, tid = threadIdx.x;
  const float* x = a.x + row * a.ld;
  const int end = min(c0 + EN_CHUNK, a.L);
  double s0 = 0.0, s1 = 0.0;
  for (int i = c0 + tid; i < end; i += 256) {
    float v = __ldg(x + i);
    if (a.mode == 0) v = elu_f(v);
    s0 += (double)v;
    s1 = fma((double)v, (double)v, s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s0; red[1][tid >> 5] = s1; }
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    atomicAdd(a.stats + row * 2 + tid, s);
  }
}
__global__ void __launch_bounds__(256) eluin_apply_kernel(WesepEluInArgs a) {
  const int64_t row = blockIdx.y;
  const int c0 = blockIdx.x * EN_CHUNK, tid = threadIdx.x;
  const double mean = a.stats[row * 2] / (double)a.L;
  const double var = fmax(a.stats[row * 2 + 1] / (double)a.L - mean * mean, 0.0);   // biased, nn.InstanceNorm*
  const float mu = (float)mean, rs = (float)(1.0 / sqrt(var + (double)a.eps));
  if (blockIdx.x == 0 && tid == 0) { a.mr[row * 2] = mu; a.mr[row * 2 + 1] = rs; }
  const float* x = a.x + row * a.ld;
  float* y = a.y + row * a.ld;
  const int end = min(c0 + EN_CHUNK, a.L);
  for (int i = c0 + tid; i < end; i += 256) {
    const float v = __ldg(x + i);
    y[i] = a.mode == 0 ? (elu_f(v) - mu) * rs : elu_f((v - mu) * rs);
  }
}
// backward: u = gy (mode 0) | gy * elu'(xhat) (mode 1);  bsum[row] = (sum u, sum u * xhat), xhat = the normalised value
__global__ void __launch_bounds__(256) eluin_bwd_reduce_kernel(WesepEluInArgs a) {
  __shared__ double red[2][8];
  const int64_t row = blockIdx.y;
  const int c0 = blockIdx.x * EN_CHUNK, tid = threadIdx.x;
  const float mu = __ldg(a.mr + row * 2), rs = __ldg(a.mr + row * 2 + 1);
  const float* x = a.x + row * a.ld;
  const float* gy = a.gy + row * a.ld;
  const int end = min(c0 + EN_CHUNK, a.L);
  double s0 = 0.0, s1 = 0.0;
  for (int i = c0 + tid; i < end; i += 256) {
    const float v = __ldg(x + i);
    float u = __ldg(gy + i), xh;
    if (a.mode == 0) xh = (elu_f(v) - mu) * rs;
    else { xh = (v - mu) * rs; u *= elu_d(xh); }
    s0 += (double)u;
    s1 = fma((double)u, (double)xh, s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s0; red[1][tid >> 5] = s1; }
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    atomicAdd(a.stats + row * 2 + tid, s);
  }
}
__global__ void __launch_bounds__(256) eluin_bwd_apply_kernel(WesepEluInArgs a) {
  const int64_t row = blockIdx.y;
  const int c0 = blockIdx.x * EN_CHUNK, tid = threadIdx.x;
  const float mu = __ldg(a.mr + row * 2), rs = __ldg(a.mr + row * 2 + 1);
  const float m0 = (float)(a.stats[row * 2] / (double)a.L), m1 = (float)(a.stats[row * 2 + 1] / (double)a.L);
  const float* x = a.x + row * a.ld;
  const float* gy = a.gy + row * a.ld;
  float* gx = a.gx + row * a.ld;
  const int end = min(c0 + EN_CHUNK, a.L);
  for (int i = c0 + tid; i < end; i += 256) {
    const float v = __ldg(x + i);
    float u = __ldg(gy + i), xh;
    if (a.mode == 0) xh = (elu_f(v) - mu) * rs;
    else { xh = (v - mu) * rs; u *= elu_d(xh); }
    float g = rs * (u - m0 - xh * m1);
    if (a.mode == 0) g *= elu_d(v);
    gx[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ depthwise Conv1d k=3
// y[n][c][t] = b[c] + sum_k w[c][k] x[n][c][t + (k - 1) d]   (zero padding d each side, convs.py:122-131)
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(WesepDwConv1dArgs a) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ld;
  float* y = a.y + ((int64_t)n * a.C + c) * a.ld;
  const float w0 = __ldg(a.w + c * 3), w1 = __ldg(a.w + c * 3 + 1), w2 = __ldg(a.w + c * 3 + 2), b = a.b ? __ldg(a.b + c) : 0.f;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < a.L; t += gridDim.x * 256) {
    const float xm = t - a.dil >= 0 ? __ldg(x + t - a.dil) : 0.f;
    const float xp = t + a.dil < a.L ? __ldg(x + t + a.dil) : 0.f;
    y[t] = fmaf(w0, xm, fmaf(w1, __ldg(x + t), fmaf(w2, xp, b)));
  }
}
// gx[t] = sum_k w[k] gy[t - (k - 1) d];  gw[c][k] += sum_t gy[t] x[t + (k - 1) d];  gb[c] += sum_t gy[t]  (fp32 atomics per CTA)
__global__ void __launch_bounds__(256) dwconv_bwd_kernel(WesepDwConv1dArgs a) {
  __shared__ float red[4 * 32];
  const int c = blockIdx.y, n = blockIdx.z;
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ld;
  const float* gy = a.gy + ((int64_t)n * a.C + c) * a.ld;
  float* gx = a.gx + ((int64_t)n * a.C + c) * a.ld;
  const float w0 = __ldg(a.w + c * 3), w1 = __ldg(a.w + c * 3 + 1), w2 = __ldg(a.w + c * 3 + 2);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = blockIdx.x * 256 + threadIdx.x; t < a.L; t += gridDim.x * 256) {
    const float g = __ldg(gy + t);
    const float gm = t - a.dil >= 0 ? __ldg(gy + t - a.dil) : 0.f;
    const float gp = t + a.dil < a.L ? __ldg(gy + t + a.dil) : 0.f;
    gx[t] = fmaf(w0, gp, fmaf(w1, g, w2 * gm));
    const float xm = t - a.dil >= 0 ? __ldg(x + t - a.dil) : 0.f;
    const float xp = t + a.dil < a.L ? __ldg(x + t + a.dil) : 0.f;
    v[0] = fmaf(g, xm, v[0]); v[1] = fmaf(g, __ldg(x + t), v[1]); v[2] = fmaf(g, xp, v[2]); v[3] += g;
  }
  block_sum<4>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.gw + c * 3, v[0]); atomicAdd(a.gw + c * 3 + 1, v[1]); atomicAdd(a.gw + c * 3 + 2, v[2]);
    if (a.gb) atomicAdd(a.gb + c, v[3]);
  }
}

// ------------------------------------------------------------------------------------------------ AvgPool2d(k) / bilinear
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(WesepPool2dArgs a) {
  const int64_t row = blockIdx.y;
  const float* x = a.x + row * a.ldx;
  float* y = a.y + row * a.ldy;
  const float inv = 1.f / (float)(a.k * a.k);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.Ho * a.Wo; i += gridDim.x * 256) {
    const int ho = i / a.Wo, wo = i - ho * a.Wo;
    float s = 0.f;
    for (int dh = 0; dh < a.k; ++dh)
      for (int dw = 0; dw < a.k; ++dw) s += __ldg(x + (int64_t)(ho * a.k + dh) * a.W + wo * a.k + dw);
    y[i] = s * inv;
  }
}
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(WesepPool2dArgs a) {
  const int64_t row = blockIdx.y;
  const float* gy = a.gy + row * a.ldy;
  float* gx = a.gx + row * a.ldx;
  const float inv = 1.f / (float)(a.k * a.k);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.H * a.W; i += gridDim.x * 256) {
    const int h = i / a.W, w = i - h * a.W;
    const int ho = h / a.k, wo = w / a.k;
    gx[i] = (ho < a.Ho && wo < a.Wo) ? __ldg(gy + (int64_t)ho * a.Wo + wo) * inv : 0.f;
  }
}
// nn.Upsample(size, mode="bilinear") (align_corners False): src = max(scale (dst + 0.5) - 0.5, 0), scale = in / out
__device__ __forceinline__ void bl_coord(int d, float scale, int n_in, int& i0, int& i1, float& l1) {
  float s = scale * ((float)d + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = min((int)s, n_in - 1);
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
}
__global__ void __launch_bounds__(256) upsample_fwd_kernel(WesepUpsample2dArgs a) {
  const int64_t row = blockIdx.y;
  const float* x = a.x + row * a.ldi;
  float* y = a.y + row * a.ldo;
  const float sh = (float)a.Hi / (float)a.Ho, sw = (float)a.Wi / (float)a.Wo;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.Ho * a.Wo; i += gridDim.x * 256) {
    const int ho = i / a.Wo, wo = i - ho * a.Wo;
    int h0, h1, w0, w1; float lh, lw;
    bl_coord(ho, sh, a.Hi, h0, h1, lh);
    bl_coord(wo, sw, a.Wi, w0, w1, lw);
    const float v00 = __ldg(x + (int64_t)h0 * a.Wi + w0), v01 = __ldg(x + (int64_t)h0 * a.Wi + w1);
    const float v10 = __ldg(x + (int64_t)h1 * a.Wi + w0), v11 = __ldg(x + (int64_t)h1 * a.Wi + w1);
    y[i] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
  }
}
// gx zeroed by the entry point.  The low-resolution map of one plane is small (<= UP_SMEM_MAX floats): every CTA accumulates its
// chunk of output pixels into a shared-memory copy of the map and flushes it with one global atomic per touched element
// (a 32 x 32 pooling level concentrates 128 k scatter-adds on 120 addresses: global atomics alone took 440 us per launch).
constexpr int UP_SMEM_MAX = 10240;
constexpr int UP_CHUNK = 8192;
__global__ void __launch_bounds__(256) upsample_bwd_kernel(WesepUpsample2dArgs a) {
  extern __shared__ float up_acc[];
  const int64_t row = blockIdx.y;
  const float* gy = a.gy + row * a.ldo;
  float* gx = a.gx + row * a.ldi;
  const int nin = a.Hi * a.Wi;
  const bool use_smem = nin <= UP_SMEM_MAX;
  if (use_smem) {
    for (int i = threadIdx.x; i < nin; i += 256) up_acc[i] = 0.f;
    __syncthreads();
  }
  float* acc = use_smem ? up_acc : gx;
  const float sh = (float)a.Hi / (float)a.Ho, sw = (float)a.Wi / (float)a.Wo;
  const int c0 = blockIdx.x * UP_CHUNK, end = min(c0 + UP_CHUNK, a.Ho * a.Wo);
  for (int i = c0 + threadIdx.x; i < end; i += 256) {
    const int ho = i / a.Wo, wo = i - ho * a.Wo;
    int h0, h1, w0, w1; float lh, lw;
    bl_coord(ho, sh, a.Hi, h0, h1, lh);
    bl_coord(wo, sw, a.Wi, w0, w1, lw);
    const float g = __ldg(gy + i);
    atomicAdd(acc + h0 * a.Wi + w0, (1.f - lh) * (1.f - lw) * g);
    atomicAdd(acc + h0 * a.Wi + w1, (1.f - lh) * lw * g);
    atomicAdd(acc + h1 * a.Wi + w0, lh * (1.f - lw) * g);
    atomicAdd(acc + h1 * a.Wi + w1, lh * lw * g);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nin; i += 256) {
      const float v = up_acc[i];
      if (v != 0.f) atomicAdd(gx + i, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ per-frequency gain
// y[n][c][t * F + f] = x * s[n][f];  gs[n][f] += sum_{c,t} gy x  (zeroed by the entry point)
constexpr int CS_MAXF = 1024;
template <bool BWD>
__global__ void __launch_bounds__(256) colscale_kernel(WesepColScaleArgs a) {
  __shared__ float sacc[BWD ? CS_MAXF : 1];
  const int c = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
  const int64_t row = ((int64_t)n * a.C + c) * a.ld;
  const float* s = a.s + (int64_t)n * a.F;
  if (BWD) {
    for (int f = tid; f < a.F; f += 256) sacc[f] = 0.f;
    __syncthreads();
  }
  const int L = a.T * a.F;
  for (int i = blockIdx.x * 256 + tid; i < L; i += gridDim.x * 256) {
    const int f = i % a.F;
    if (!BWD) a.y[row + i] = __ldg(a.x + row + i) * __ldg(s + f);
    else {
      const float g = __ldg(a.gy + row + i);
      a.gx[row + i] = g * __ldg(s + f);
      atomicAdd(&sacc[f], g * __ldg(a.x + row + i));
    }
  }
  if (BWD) {
    __syncthreads();
    for (int f = tid; f < a.F; f += 256) atomicAdd(a.gs + (int64_t)n * a.F + f, sacc[f]);
  }
}

}  // namespace wb

using namespace wb;

static int check_eluin(const WesepEluInArgs* a) {
  if (!a || a->rows <= 0 || a->rows > 65535 || a->L <= 0 || a->ld < a->L) return fail(-1, "elu_in: bad shape (rows <= 65535)");
  if (a->mode != 0 && a->mode != 1) return fail(-1, "elu_in: mode");
  if (!a->x || !a->stats || !a->mr) return fail(-1, "elu_in: null buffer");
  return 0;
}
extern "C" int wesep_b200_elu_in_fwd(const WesepEluInArgs* a, void* stream) {
  if (int rc = check_eluin(a)) return rc;
  if (!a->y) return fail(-1, "elu_in: null output");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)a->rows * 2 * sizeof(double), st));
  const dim3 grid(cdiv(a->L, EN_CHUNK), (unsigned)a->rows);
  eluin_stats_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("elu_in_stats");
  eluin_apply_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("elu_in_apply");
  return 0;
}
extern "C" int wesep_b200_elu_in_bwd(const WesepEluInArgs* a, void* stream) {
  if (int rc = check_eluin(a)) return rc;
  if (!a->gy || !a->gx) return fail(-1, "elu_in: null gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->stats, 0, (size_t)a->rows * 2 * sizeof(double), st));
  const dim3 grid(cdiv(a->L, EN_CHUNK), (unsigned)a->rows);
  eluin_bwd_reduce_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("elu_in_bwd_reduce");
  eluin_bwd_apply_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("elu_in_bwd_apply");
  return 0;
}

static int check_dw(const WesepDwConv1dArgs* a) {
  if (!a || a->n <= 0 || a->n > 65535 || a->C <= 0 || a->C > 65535 || a->L <= 0 || a->ld < a->L || a->dil <= 0)
    return fail(-1, "dwconv1d: bad shape");
  if (!a->x || !a->w) return fail(-1, "dwconv1d: null buffer");
  return 0;
}
extern "C" int wesep_b200_dwconv1d_fwd(const WesepDwConv1dArgs* a, void* stream) {
  if (int rc = check_dw(a)) return rc;
  if (!a->y) return fail(-1, "dwconv1d: null output");
  dwconv_fwd_kernel<<<dim3(cdiv(a->L, 2048), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("dwconv1d_fwd");
  return 0;
}
extern "C" int wesep_b200_dwconv1d_bwd(const WesepDwConv1dArgs* a, void* stream) {
  if (int rc = check_dw(a)) return rc;
  if (!a->gy || !a->gx || !a->gw) return fail(-1, "dwconv1d: null gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->gw, 0, (size_t)a->C * 3 * sizeof(float), st));
  if (a->gb) WB_CUDA(cudaMemsetAsync(a->gb, 0, (size_t)a->C * sizeof(float), st));
  dwconv_bwd_kernel<<<dim3(cdiv(a->L, 2048), a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("dwconv1d_bwd");
  return 0;
}

static int check_pool(const WesepPool2dArgs* a) {
  if (!a || a->rows <= 0 || a->rows > 65535 || a->H <= 0 || a->W <= 0 || a->k <= 0) return fail(-1, "avgpool2d: bad shape");
  if (a->Ho != a->H / a->k || a->Wo != a->W / a->k || a->Ho <= 0 || a->Wo <= 0) return fail(-1, "avgpool2d: output size");
  if (a->ldx < (int64_t)a->H * a->W || a->ldy < (int64_t)a->Ho * a->Wo) return fail(-1, "avgpool2d: row strides");
  return 0;
}
extern "C" int wesep_b200_avgpool2d_fwd(const WesepPool2dArgs* a, void* stream) {
  if (int rc = check_pool(a)) return rc;
  if (!a->x || !a->y) return fail(-1, "avgpool2d: null buffer");
  avgpool_fwd_kernel<<<dim3(cdiv((int64_t)a->Ho * a->Wo, 256), (unsigned)a->rows), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("avgpool2d_fwd");
  return 0;
}
extern "C" int wesep_b200_avgpool2d_bwd(const WesepPool2dArgs* a, void* stream) {
  if (int rc = check_pool(a)) return rc;
  if (!a->gy || !a->gx) return fail(-1, "avgpool2d: null buffer");
  avgpool_bwd_kernel<<<dim3(cdiv((int64_t)a->H * a->W, 1024), (unsigned)a->rows), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("avgpool2d_bwd");
  return 0;
}

static int check_up(const WesepUpsample2dArgs* a) {
  if (!a || a->rows <= 0 || a->rows > 65535 || a->Hi <= 0 || a->Wi <= 0 || a->Ho <= 0 || a->Wo <= 0) return fail(-1, "upsample2d: bad shape");
  if (a->ldi < (int64_t)a->Hi * a->Wi || a->ldo < (int64_t)a->Ho * a->Wo) return fail(-1, "upsample2d: row strides");
  return 0;
}
extern "C" int wesep_b200_upsample2d_fwd(const WesepUpsample2dArgs* a, void* stream) {
  if (int rc = check_up(a)) return rc;
  if (!a->x || !a->y) return fail(-1, "upsample2d: null buffer");
  upsample_fwd_kernel<<<dim3(cdiv((int64_t)a->Ho * a->Wo, 1024), (unsigned)a->rows), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("upsample2d_fwd");
  return 0;
}
extern "C" int wesep_b200_upsample2d_bwd(const WesepUpsample2dArgs* a, void* stream) {
  if (int rc = check_up(a)) return rc;
  if (!a->gy || !a->gx) return fail(-1, "upsample2d: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->gx, 0, (size_t)a->rows * a->ldi * sizeof(float), st));
  const size_t smem = (size_t)a->Hi * a->Wi <= (size_t)UP_SMEM_MAX ? (size_t)a->Hi * a->Wi * sizeof(float) : 0;
  upsample_bwd_kernel<<<dim3(cdiv((int64_t)a->Ho * a->Wo, UP_CHUNK), (unsigned)a->rows), 256, smem, st>>>(*a);
  WB_LAUNCH_CHECK("upsample2d_bwd");
  return 0;
}

static int check_cs(const WesepColScaleArgs* a) {
  if (!a || a->n <= 0 || a->n > 65535 || a->C <= 0 || a->C > 65535 || a->T <= 0 || a->F <= 0 || a->F > CS_MAXF)
    return fail(-1, "colscale: bad shape (F <= 1024)");
  if (a->ld < (int64_t)a->T * a->F || !a->x || !a->s) return fail(-1, "colscale: row stride / null buffer");
  return 0;
}
extern "C" int wesep_b200_colscale_fwd(const WesepColScaleArgs* a, void* stream) {
  if (int rc = check_cs(a)) return rc;
  if (!a->y) return fail(-1, "colscale: null output");
  colscale_kernel<false><<<dim3(cdiv((int64_t)a->T * a->F, 2048), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("colscale_fwd");
  return 0;
}
extern "C" int wesep_b200_colscale_bwd(const WesepColScaleArgs* a, void* stream) {
  if (int rc = check_cs(a)) return rc;
  if (!a->gy || !a->gx || !a->gs) return fail(-1, "colscale: null gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->gs, 0, (size_t)a->n * a->F * sizeof(float), st));
  colscale_kernel<true><<<dim3(cdiv((int64_t)a->T * a->F, 8192), a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("colscale_bwd");
  return 0;
}
