// wespeaker ECAPA-TDNN building blocks: the dilated Conv1d of the TDNN / Res2 layers as a gather (im2col1d)
// in front of the tcgen05 pointwise GEMM, the squeeze-excitation gate's ReLU / sigmoid, and the attentive statistics of ASTP.
// BatchNorm1d, the 1x1 convolutions, tanh, the softmax over time and the global-context statistics reuse bn2 / conv1x1 / tanh /
// softmax / tstp.  All HBM-bound streaming kernels on act tensors [n][C][T].
#include "common.cuh"

namespace wb {

// col[n][c*K + k][l] = x[n][c][l * stride + k * dil - pad]   (0 outside [0, T)), l < Tout.
// 'same' Conv1d: stride 1, pad = dil (K-1)/2, Tout = T;  F.unfold(kernel (K, 1), stride (hs, 1)): dil 1, pad 0, Tout = (T-K)/hs + 1
__device__ __forceinline__ void i1_geom(const WesepIm2col1dArgs& a, int& stride, int& pad, int& Tout) {
  stride = a.stride > 0 ? a.stride : 1;
  pad = a.unfold ? 0 : a.dil * (a.K - 1) / 2;
  Tout = a.unfold ? (a.T - a.K) / stride + 1 : a.T;
}
__global__ void __launch_bounds__(256) im2col1d_kernel(WesepIm2col1dArgs a) {
  const int r = blockIdx.y, n = blockIdx.z;
  const int c = r / a.K, k = r - c * a.K;
  int stride, pad, Tout;
  i1_geom(a, stride, pad, Tout);
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ldx;
  float* col = a.col + (int64_t)n * a.bsc + (int64_t)r * a.ldc;
  for (int l = blockIdx.x * 256 + threadIdx.x; l < Tout; l += gridDim.x * 256) {
    const int s = l * stride + k * a.dil - pad;
    col[l] = (s >= 0 && s < a.T) ? __ldg(x + s) : 0.f;
  }
}
// adjoint: gx[n][c][s] = sum over (k, l) with l * stride + k * dil - pad == s of gcol[n][c*K + k][l]
__global__ void __launch_bounds__(256) col2im1d_kernel(WesepIm2col1dArgs a) {
  const int c = blockIdx.y, n = blockIdx.z;
  int stride, pad, Tout;
  i1_geom(a, stride, pad, Tout);
  const float* g = a.gcol + (int64_t)n * a.bsc + (int64_t)c * a.K * a.ldc;
  float* gx = a.gx + ((int64_t)n * a.C + c) * a.ldx;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < a.T; s += gridDim.x * 256) {
    float acc = 0.f;
    for (int k = 0; k < a.K; ++k) {
      const int num = s + pad - k * a.dil;
      if (num < 0 || num % stride) continue;
      const int l = num / stride;
      if (l < Tout) acc += __ldg(g + (int64_t)k * a.ldc + l);
    }
    gx[s] = acc;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256) unary_kernel(WesepUnaryArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.count; i += (int64_t)gridDim.x * 256) {
    if (!BWD) {
      const float v = __ldg(a.x + i);
      a.y[i] = a.mode == 0 ? fmaxf(v, 0.f) : 1.f / (1.f + __expf(-v));
    } else {
      const float y = __ldg(a.y + i), g = __ldg(a.gy + i);
      a.gx[i] = a.mode == 0 ? (y > 0.f ? g : 0.f) : g * y * (1.f - y);
    }
  }
}

// out = a + b over `count` floats (float4 body)
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t count) {
  const int64_t n4 = count >> 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) out[i] = __ldg(a + i) + __ldg(b + i);
}
// y[r][t] = x[r][t] * v[t]
__global__ void __launch_bounds__(256) colvec_mul_kernel(WesepColVecArgs a) {
  const int64_t r = blockIdx.y;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < a.L; t += gridDim.x * 256) a.y[r * a.ldy + t] = __ldg(a.x + r * a.ldx + t) * __ldg(a.v + t);
}

// one warp per (n, c) row
template <bool BWD>
__global__ void __launch_bounds__(256) astp_kernel(WesepAstpArgs a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.n * a.C) return;
  const int n = row / a.C, c = row - n * a.C;
  const float* x = a.x + (int64_t)row * a.ld;
  const float* al = a.alpha + (int64_t)row * a.ld;
  float s0 = 0.f, s1 = 0.f;
  for (int t = lane; t < a.T; t += 32) {
    const float v = __ldg(x + t), w = __ldg(al + t);
    s0 = fmaf(w, v, s0);
    s1 = fmaf(w * v, v, s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  const float mean = s0, var = s1 - mean * mean;
  const bool live = var > 1e-10f;
  const float sd = sqrtf(live ? var : 1e-10f);
  if constexpr (!BWD) {
    if (lane == 0) {
      a.out[(int64_t)n * 2 * a.C + c] = mean;
      a.out[(int64_t)n * 2 * a.C + a.C + c] = sd;
    }
  } else {
    const float gm = __ldg(a.gout + (int64_t)n * 2 * a.C + c), gs = __ldg(a.gout + (int64_t)n * 2 * a.C + a.C + c);
    const float dvar = live ? gs / (2.f * sd) : 0.f;           // clamp(min = 1e-10): no gradient below the floor
    const float dmean = gm - 2.f * mean * dvar;
    float* gx = a.gx + (int64_t)row * a.ld;
    float* ga = a.galpha + (int64_t)row * a.ld;
    for (int t = lane; t < a.T; t += 32) {
      const float v = __ldg(x + t), w = __ldg(al + t);
      gx[t] = w * fmaf(2.f * v, dvar, dmean);
      ga[t] = v * fmaf(v, dvar, dmean);
    }
  }
}

}  // namespace wb

using namespace wb;

static int tout_of(const WesepIm2col1dArgs* a) { return a->unfold ? (a->T - a->K) / (a->stride > 0 ? a->stride : 1) + 1 : a->T; }
static int check_i1(const WesepIm2col1dArgs* a) {
  if (!a || a->n <= 0 || a->n > 65535 || a->C <= 0 || a->T <= 0 || a->K <= 0 || a->dil <= 0 || a->stride < 0) return fail(-1, "im2col1d: shape");
  if (!a->unfold && (!(a->K & 1) || a->stride > 1)) return fail(-1, "im2col1d: the 'same' convolution needs an odd kernel and stride 1");
  if (a->unfold && (a->dil != 1 || a->T < a->K)) return fail(-1, "im2col1d: unfold needs dilation 1 and T >= K");
  if ((int64_t)a->C * a->K > 65535) return fail(-2, "im2col1d: too many gathered rows for one launch");
  const int tout_ = a->unfold ? (a->T - a->K) / (a->stride > 0 ? a->stride : 1) + 1 : a->T;
  if (a->ldx < a->T || a->ldc < tout_ || a->bsc < (int64_t)a->C * a->K * a->ldc) return fail(-1, "im2col1d: strides");
  return 0;
}
extern "C" int wesep_b200_im2col1d_fwd(const WesepIm2col1dArgs* a, void* stream) {
  if (int rc = check_i1(a)) return rc;
  if (!a->x || !a->col) return fail(-1, "im2col1d: pointers");
  im2col1d_kernel<<<dim3(cdiv(tout_of(a), 1024), a->C * a->K, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("im2col1d_fwd");
  return 0;
}
extern "C" int wesep_b200_im2col1d_bwd(const WesepIm2col1dArgs* a, void* stream) {
  if (int rc = check_i1(a)) return rc;
  if (!a->gcol || !a->gx) return fail(-1, "col2im1d: pointers");
  col2im1d_kernel<<<dim3(cdiv(a->T, 1024), a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("im2col1d_bwd");
  return 0;
}
extern "C" int wesep_b200_unary_fwd(const WesepUnaryArgs* a, void* stream) {
  if (!a || a->count <= 0 || (a->mode != 0 && a->mode != 1) || !a->x || !a->y) return fail(-1, "unary: bad arguments");
  unary_kernel<false><<<(unsigned)((a->count + 1023) / 1024 < 4096 ? (a->count + 1023) / 1024 : 4096), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("unary_fwd");
  return 0;
}
extern "C" int wesep_b200_unary_bwd(const WesepUnaryArgs* a, void* stream) {
  if (!a || a->count <= 0 || (a->mode != 0 && a->mode != 1) || !a->y || !a->gy || !a->gx) return fail(-1, "unary_bwd: bad arguments");
  unary_kernel<true><<<(unsigned)((a->count + 1023) / 1024 < 4096 ? (a->count + 1023) / 1024 : 4096), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("unary_bwd");
  return 0;
}
extern "C" int wesep_b200_astp_fwd(const WesepAstpArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->C <= 0 || a->T <= 0 || a->ld < a->T || !a->x || !a->alpha || !a->out) return fail(-1, "astp: bad arguments");
  astp_kernel<false><<<cdiv((int64_t)a->n * a->C, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("astp_fwd");
  return 0;
}
extern "C" int wesep_b200_astp_bwd(const WesepAstpArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->C <= 0 || a->T <= 0 || a->ld < a->T || !a->x || !a->alpha || !a->gout || !a->gx || !a->galpha)
    return fail(-1, "astp_bwd: bad arguments");
  astp_kernel<true><<<cdiv((int64_t)a->n * a->C, 8), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("astp_bwd");
  return 0;
}

extern "C" int wesep_b200_add(const WesepAddArgs* p, void* stream) {
  if (!p || !p->a || !p->b || !p->out || p->count <= 0) return fail(-1, "add: bad arguments");
  if (!aligned16(p->a) || !aligned16(p->b) || !aligned16(p->out)) return fail(-1, "add: 16-byte alignment");
  const int64_t blocks = (p->count / 4 + 255) / 256;
  add_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks > 65535 ? 65535 : blocks)), 256, 0, (cudaStream_t)stream>>>(p->a, p->b, p->out, p->count);
  WB_LAUNCH_CHECK("add");
  return 0;
}
extern "C" int wesep_b200_colvec_mul(const WesepColVecArgs* a, void* stream) {
  if (!a || a->rows <= 0 || a->rows > 65535 || a->L <= 0 || a->ldx < a->L || a->ldy < a->L || !a->x || !a->v || !a->y) return fail(-1, "colvec_mul: bad arguments");
  colvec_mul_kernel<<<dim3(cdiv(a->L, 2048), (unsigned)a->rows), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("colvec_mul");
  return 0;
}
