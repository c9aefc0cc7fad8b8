// SI-SDR loss (auraloss.time.SISDRLoss restated) and the fused per-tensor
// clip + Adam step over a flat parameter arena (wesep/utils/funcs.py:79-88 + torch.optim.Adam).
#include "common.cuh"

namespace wb {

constexpr int SD_CHUNK = 4096;  // samples per CTA
constexpr double SD_EPS = 1e-8;

__global__ void __launch_bounds__(256) sisdr_sums_kernel(WesepSisdrFwdArgs a) {
  __shared__ double red[5][8];
  const int i = blockIdx.z, n = blockIdx.y, c0 = blockIdx.x * SD_CHUNK, tid = threadIdx.x;
  const float* x = a.est[i] + (int64_t)n * a.ld_est[i];
  const float* t = a.tgt + (int64_t)n * a.ld_tgt;
  // fp64 accumulation: Q = <x~,x~> - 2a<x~,t~> + a^2<t~,t~> cancels to ~1e-6 of its terms at 60 dB
  double sx = 0.0, st = 0.0, sxt = 0.0, sxx = 0.0, stt = 0.0;
  const int end = min(c0 + SD_CHUNK, a.L);
  for (int s = c0 + tid; s < end; s += 256) {
    const double xv = (double)__ldg(x + s), tv = (double)__ldg(t + s);
    sx += xv; st += tv;
    sxt = fma(xv, tv, sxt); sxx = fma(xv, xv, sxx); stt = fma(tv, tv, stt);
  }
  double v[5] = {sx, st, sxt, sxx, stt};
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = warp_sum(v[k]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) red[k][tid >> 5] = v[k];
  }
  __syncthreads();
  if (tid < 5) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    atomicAdd(a.sums + ((int64_t)i * a.n + n) * 5 + tid, s);
  }
}

struct SdRow {
  double mx, mt, al, b, P, Q, R;
};
__device__ __forceinline__ SdRow sd_row(const double* s, int L) {
  SdRow r;
  const double Ld = (double)L;
  r.mx = s[0] / Ld; r.mt = s[1] / Ld;
  const double a = s[2] - s[0] * s[1] / Ld;   // <x~,t~>
  r.b = s[4] - s[1] * s[1] / Ld;              // <t~,t~>
  const double xx = s[3] - s[0] * s[0] / Ld;  // <x~,x~>
  r.al = a / (r.b + SD_EPS);
  r.P = r.al * r.al * r.b;
  r.Q = xx - 2.0 * r.al * a + r.al * r.al * r.b;
  if (r.Q < 0.0) r.Q = 0.0;
  r.R = r.P / (r.Q + SD_EPS) + SD_EPS;
  return r;
}

__global__ void sisdr_finalize_kernel(WesepSisdrFwdArgs a) {
  // one warp per estimate
  const int i = blockIdx.x, lane = threadIdx.x;
  double acc = 0.0;
  for (int n = lane; n < a.n; n += 32) {
    SdRow r = sd_row(a.sums + ((int64_t)i * a.n + n) * 5, a.L);
    double v = 10.0 * log10(r.R);
    a.sisdr_rows[(int64_t)i * a.n + n] = (float)v;
    acc += v;
  }
  acc = warp_sum(acc);
  if (lane == 0) a.loss[i] = (float)(-acc / (double)a.n);
}

__global__ void __launch_bounds__(256) sisdr_bwd_kernel(WesepSisdrBwdArgs a) {
  const int i = blockIdx.z, n = blockIdx.y, c0 = blockIdx.x * SD_CHUNK, tid = threadIdx.x;
  SdRow r = sd_row(a.sums + ((int64_t)i * a.n + n) * 5, a.L);
  // l = -10 log10(R): dl/dx~ = c1 x~ + c2 t~ (closed form, checked against autograd)
  const double k = 10.0 / log(10.0);
  const double q = r.Q + SD_EPS;
  const double be = r.b + SD_EPS;
  double c1 = k / r.R * 2.0 * r.P / (q * q);
  double c2 = -k / r.R * (2.0 * r.al * r.b / (be * q) + r.P / (q * q) * (4.0 * r.al - 2.0 * r.al * r.b / be));
  const double gl = (double)__ldg(a.gloss + i) / (double)a.n;
  const float f1 = (float)(c1 * gl), f2 = (float)(c2 * gl);
  const float mx = (float)r.mx, mt = (float)r.mt;
  const float* x = a.est[i] + (int64_t)n * a.ld_est[i];
  const float* t = a.tgt + (int64_t)n * a.ld_tgt;
  float* g = a.gest[i] + (int64_t)n * a.ld_gest[i];
  const int end = min(c0 + SD_CHUNK, a.L);
  for (int s = c0 + tid; s < end; s += 256) g[s] = f1 * (__ldg(x + s) - mx) + f2 * (__ldg(t + s) - mt);
}

// ------------------------------------------------------------------------------------ clip + Adam
constexpr int OPT_CHUNK = 4096;

__global__ void __launch_bounds__(256) opt_sumsq_kernel(WesepClipAdamArgs a) {
  __shared__ float red[32];
  const int ch = blockIdx.x, tid = threadIdx.x;
  const int seg = a.chunk_seg[ch];
  const int64_t off = a.chunk_off[ch];
  const int64_t end = min(off + (int64_t)OPT_CHUNK, a.seg_off[seg + 1]);
  float s = 0.f;
  for (int64_t j = off + tid * 4; j < end; j += 1024) {
    float4 g = *reinterpret_cast<const float4*>(a.grad + j);
    g.x *= a.grad_scale; g.y *= a.grad_scale; g.z *= a.grad_scale; g.w *= a.grad_scale;
    s += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
  }
  float v[1] = {s};
  block_sum<1>(v, red);
  if (tid == 0) atomicAdd(a.sumsq + seg, (double)v[0]);
}

__global__ void __launch_bounds__(256) opt_adam_kernel(WesepClipAdamArgs a, float bc1, float bc2_sqrt) {
  const int ch = blockIdx.x, tid = threadIdx.x;
  const int seg = a.chunk_seg[ch];
  const int64_t off = a.chunk_off[ch];
  const int64_t end = min(off + (int64_t)OPT_CHUNK, a.seg_off[seg + 1]);
  const float norm = (float)sqrt(a.sumsq[seg]);
  if (tid == 0 && off == a.seg_off[seg]) a.norms[seg] = norm;
  float gs = a.grad_scale;
  if (a.clip > 0.f) {
    float coef = a.clip / (norm + 1e-6f);
    if (coef < 1.f) gs *= coef;
  }
  if (a.dyn) {   // graph replay: schedule-dependent scalars live in device memory
    bc1 = __ldg(a.dyn + 1);
    bc2_sqrt = __ldg(a.dyn + 2);
  }
  const float step_size = (a.dyn ? __ldg(a.dyn) : a.lr) / bc1;
  for (int64_t j = off + tid * 4; j < end; j += 1024) {
    float4 g = *reinterpret_cast<const float4*>(a.grad + j);
    float4 p = *reinterpret_cast<const float4*>(a.param + j);
    float4 m = *reinterpret_cast<const float4*>(a.exp_avg + j);
    float4 v = *reinterpret_cast<const float4*>(a.exp_avg_sq + j);
    float* gp = &g.x; float* pp = &p.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gc = gp[q] * gs;
      gp[q] = gc;
      const float g2 = fmaf(a.weight_decay, pp[q], gc);
      mp[q] = a.beta1 * mp[q] + (1.f - a.beta1) * g2;
      vp[q] = a.beta2 * vp[q] + (1.f - a.beta2) * g2 * g2;
      const float denom = sqrtf(vp[q]) / bc2_sqrt + a.eps;
      pp[q] -= step_size * (mp[q] / denom);
    }
    *reinterpret_cast<float4*>(a.grad + j) = g;
    *reinterpret_cast<float4*>(a.param + j) = p;
    *reinterpret_cast<float4*>(a.exp_avg + j) = m;
    *reinterpret_cast<float4*>(a.exp_avg_sq + j) = v;
  }
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_sisdr_fwd(const WesepSisdrFwdArgs* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a->n_est < 1 || a->n_est > 4 || a->n <= 0 || a->L <= 0) return fail(-1, "sisdr: bad shape");
  WB_CUDA(cudaMemsetAsync(a->sums, 0, sizeof(double) * 5 * a->n * a->n_est, st));
  sisdr_sums_kernel<<<dim3(cdiv(a->L, SD_CHUNK), a->n, a->n_est), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("sisdr_sums");
  sisdr_finalize_kernel<<<a->n_est, 32, 0, st>>>(*a);
  WB_LAUNCH_CHECK("sisdr_finalize");
  return 0;
}

extern "C" int wesep_b200_sisdr_bwd(const WesepSisdrBwdArgs* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a->n_est < 1 || a->n_est > 4 || a->n <= 0 || a->L <= 0) return fail(-1, "sisdr: bad shape");
  sisdr_bwd_kernel<<<dim3(cdiv(a->L, SD_CHUNK), a->n, a->n_est), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("sisdr_bwd");
  return 0;
}

extern "C" int wesep_b200_clip_adam(const WesepClipAdamArgs* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a->n_seg <= 0 || a->n_chunk <= 0 || a->step < 1) return fail(-1, "clip_adam: bad shape");
  if (!aligned16(a->param) || !aligned16(a->grad) || !aligned16(a->exp_avg) || !aligned16(a->exp_avg_sq))
    return fail(-1, "clip_adam: arenas must be 16-byte aligned");
  WB_CUDA(cudaMemsetAsync(a->sumsq, 0, sizeof(double) * a->n_seg, st));
  opt_sumsq_kernel<<<a->n_chunk, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("opt_sumsq");
  const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
  const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
  opt_adam_kernel<<<a->n_chunk, 256, 0, st>>>(*a, (float)bc1, (float)sqrt(bc2));
  WB_LAUNCH_CHECK("opt_adam");
  return 0;
}
