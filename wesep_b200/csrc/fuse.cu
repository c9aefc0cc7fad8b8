// Alternative speaker-fusion front of each TCN repeat (FuseSeparation with spk_fuse_type in
// {concat, additive, multiply, FiLM}: wesep/modules/tasnet/separation.py:116-135,172-181):
//     z = gLN( PReLU( a[n][c] * x + b[n][c] ) )
// where the per-(row, channel) scale/shift come from the speaker embedding (SpeakerFuseLayer
// wesep/modules/common/speaker.py:81-125, FiLM wesep/modules/common/norm.py:118-139) — the reference expands the
// embedding over all frames and runs a Linear per frame; here the Linear runs once per row.
// One CTA per (n, c) row; two passes forward (stats, apply) and two backward (sums, apply).
#include "common.cuh"

namespace wb {

constexpr float FUSE_EPS = 1e-5f;

__device__ __forceinline__ void fuse_ab(const WesepFuseArgs& a, int n, int c, float& sa, float& sb) {
  sa = a.ra ? __ldg(a.ra + (int64_t)n * a.C + c) : 1.f;
  sb = a.rb ? __ldg(a.rb + (int64_t)n * a.C + c) : 0.f;
}

// MODE 0: forward stats | 1: forward apply | 2: backward sums | 3: backward apply
template <int MODE>
__global__ void __launch_bounds__(256) fuse_kernel(WesepFuseArgs a) {
  __shared__ float red[4 * 32];
  const int c = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  float sa, sb;
  fuse_ab(a, n, c, sa, sb);
  const float al = __ldg(a.alpha);
  const float* x = a.x + ((int64_t)n * a.C + c) * a.ldx;
  float mu = 0.f, r = 1.f;
  if constexpr (MODE >= 1) gln_mean_rstd(a.stats + 2 * n, (double)a.C * a.T, a.eps > 0.f ? a.eps : FUSE_EPS, mu, r);
  const float gm = __ldg(a.gamma + c), bt = __ldg(a.beta + c);
  float m1 = 0.f, m2 = 0.f;
  if constexpr (MODE == 3) {
    const double cnt = (double)a.C * a.T;
    m1 = (float)(a.rowsums[2 * n] / cnt);
    m2 = (float)(a.rowsums[2 * n + 1] / cnt);
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int t = 4 * tid; t < a.T; t += 1024) {
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + t));
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
    float g4v[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (MODE >= 2) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gz + ((int64_t)n * a.C + c) * a.ldg + t));
      g4v[0] = g4.x; g4v[1] = g4.y; g4v[2] = g4.z; g4v[3] = g4.w;
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = t + i < a.T;
      const float v = fmaf(sa, xv[i], sb);
      const float y = prelu_f(v, al);
      if constexpr (MODE == 0) {
        if (ok) { s0 += y; s1 = fmaf(y, y, s1); }
      } else if constexpr (MODE == 1) {
        o[i] = fmaf(gm, (y - mu) * r, bt);
      } else if constexpr (MODE == 2) {
        if (ok) {
          const float yh = (y - mu) * r, g = g4v[i];
          s0 = fmaf(gm, g, s0);            // sum h
          s1 = fmaf(gm * g, yh, s1);       // sum h*yhat
          s2 = fmaf(g, yh, s2);            // dgamma
          s3 += g;                         // dbeta
        }
      } else {
        const float yh = (y - mu) * r;
        const float dy = r * (gm * g4v[i] - m1 - yh * m2);
        const float gp = ok ? dy * (v > 0.f ? 1.f : al) : 0.f;
        o[i] = gp * sa;
        if (ok) {
          s0 = fmaf(gp, xv[i], s0);        // d ra
          s1 += gp;                        // d rb
          s2 += v > 0.f ? 0.f : dy * v;    // d alpha
        }
      }
    }
    if constexpr (MODE == 1) *reinterpret_cast<float4*>(a.y + ((int64_t)n * a.C + c) * a.ldy + t) = make_float4(o[0], o[1], o[2], o[3]);
    if constexpr (MODE == 3) *reinterpret_cast<float4*>(a.dx + ((int64_t)n * a.C + c) * a.lddx + t) = make_float4(o[0], o[1], o[2], o[3]);
  }
  if constexpr (MODE == 1) return;
  float v[4] = {s0, s1, s2, s3};
  block_sum<4>(v, red);
  if (tid == 0) {
    if constexpr (MODE == 0) {
      atomicAdd(a.stats + 2 * n, (double)v[0]);
      atomicAdd(a.stats + 2 * n + 1, (double)v[1]);
    } else if constexpr (MODE == 2) {
      atomicAdd(a.rowsums + 2 * n, (double)v[0]);
      atomicAdd(a.rowsums + 2 * n + 1, (double)v[1]);
      atomicAdd(a.dgamma + c, v[2]);
      atomicAdd(a.dbeta + c, v[3]);
    } else {
      if (a.dra) a.dra[(int64_t)n * a.C + c] = v[0];
      if (a.drb) a.drb[(int64_t)n * a.C + c] = v[1];
      if (v[2] != 0.f) atomicAdd(a.dalpha, v[2]);
    }
  }
}

static int check_fuse(const WesepFuseArgs& a) {
  if (a.n <= 0 || a.C <= 0 || a.T <= 0) return fail(-1, "fuse: empty shape");
  if ((a.ldx & 3) || !aligned16(a.x)) return fail(-1, "fuse: alignment");
  return 0;
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_fuse_prelu_gln_fwd(const WesepFuseArgs* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_fuse(*a)) return rc;
  if ((a->ldy & 3) || !aligned16(a->y)) return fail(-1, "fuse: output alignment");
  WB_CUDA(cudaMemsetAsync(a->stats, 0, sizeof(double) * 2 * a->n, st));
  fuse_kernel<0><<<dim3(a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("fuse_stats");
  fuse_kernel<1><<<dim3(a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("fuse_apply");
  return 0;
}

extern "C" int wesep_b200_fuse_prelu_gln_bwd(const WesepFuseArgs* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_fuse(*a)) return rc;
  if ((a->ldg & 3) || (a->lddx & 3) || !aligned16(a->gz) || !aligned16(a->dx)) return fail(-1, "fuse: gradient alignment");
  WB_CUDA(cudaMemsetAsync(a->rowsums, 0, sizeof(double) * 2 * a->n, st));
  fuse_kernel<2><<<dim3(a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("fuse_bwd_sums");
  fuse_kernel<3><<<dim3(a->C, a->n), 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("fuse_bwd_apply");
  return 0;
}
