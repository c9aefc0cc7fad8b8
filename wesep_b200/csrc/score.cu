// Evaluation path: peak normalisation of the separated waves and batched
// SI-SNR / SI-SNRi on the device, replacing the per-utterance numpy scoring of
// wesep/bin/infer.py:124-129 (peak rule) and wesep/utils/score.py:7-36 (cal_SISNR / cal_SISNRi).
//
// Three launches per batch, all HBM-bound streaming passes over [n][L] fp32 rows:
//   1. score_peak_kernel   : per row  max(x) > 0 ?  and  max|x|          (reads est once)
//   2. score_sums_kernel   : optional in-place  est = est / max|est| * 0.9  (fp32, the reference's
//                            operation order) and the 8 fp64 moments of (est, ref, mix) over the
//                            first len[r] samples                          (reads est/ref/mix, writes est)
//   3. score_final_kernel  : closed form of cal_SISNR from the moments, per row.
#include "common.cuh"

namespace wb {

constexpr int SC_CHUNK = 8192;  // samples per CTA
constexpr double SC_EPS = 1e-8;

// ws layout per row: [0..7] fp64 moments Se,Sr,Sm,See,Srr,Smm,Ser,Smr; [8] two packed u32: bits of
// max(x,0) and bits of max|x| (both non-negative floats -> unsigned order == float order).
constexpr int SC_WS = 9;

__global__ void __launch_bounds__(256) score_peak_kernel(WesepScoreArgs a) {
  __shared__ float red[2][8];
  const int r = blockIdx.y, c0 = blockIdx.x * SC_CHUNK, tid = threadIdx.x;
  const float* e = a.est + (int64_t)r * a.ld_est;
  const int end = min(c0 + SC_CHUNK, a.L);
  float pmax = 0.f, amax = 0.f;
  for (int s = c0 + tid; s < end; s += 256) {
    const float v = __ldg(e + s);
    pmax = fmaxf(pmax, v);
    amax = fmaxf(amax, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = pmax; red[1][tid >> 5] = amax; }
  __syncthreads();
  if (tid < 2) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v = fmaxf(v, red[tid][w]);
    unsigned* pk = reinterpret_cast<unsigned*>(a.ws + (int64_t)r * SC_WS + 8);
    atomicMax(pk + tid, __float_as_uint(v));
  }
}

__global__ void __launch_bounds__(256) score_sums_kernel(WesepScoreArgs a) {
  __shared__ double red[8][8];
  __shared__ int s_norm;
  const int r = blockIdx.y, c0 = blockIdx.x * SC_CHUNK, tid = threadIdx.x;
  float* e = a.est + (int64_t)r * a.ld_est;
  const float* rf = a.ref + (int64_t)r * a.ld_ref;
  const float* mx = a.mix + (int64_t)r * a.ld_mix;
  // infer.py:124 — the whole batch is scaled iff every row has a positive sample
  if (tid == 0) {
    int ok = a.peak_norm != 0;
    for (int i = 0; ok && i < a.n; ++i) {
      const unsigned* pk = reinterpret_cast<const unsigned*>(a.ws + (int64_t)i * SC_WS + 8);
      ok = __uint_as_float(pk[0]) > 0.f;
    }
    s_norm = ok;
    if (blockIdx.x == 0 && r == 0 && a.normed) *a.normed = ok;
  }
  __syncthreads();
  const bool norm = s_norm != 0;
  const float amax = __uint_as_float(reinterpret_cast<const unsigned*>(a.ws + (int64_t)r * SC_WS + 8)[1]);
  const int len = a.len ? min(max(__ldg(a.len + r), 0), a.L) : a.L;
  const int end = min(c0 + SC_CHUNK, a.L);
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = c0 + tid; s < end; s += 256) {
    float ev = e[s];
    if (norm) {
      ev = __fmul_rn(__fdiv_rn(ev, amax), 0.9f);  // outputs / abs(outputs).max() * 0.9 in fp32
      e[s] = ev;
    }
    if (s < len) {
      const double ed = (double)ev, rd = (double)__ldg(rf + s), md = (double)__ldg(mx + s);
      v[0] += ed; v[1] += rd; v[2] += md;
      v[3] = fma(ed, ed, v[3]); v[4] = fma(rd, rd, v[4]); v[5] = fma(md, md, v[5]);
      v[6] = fma(ed, rd, v[6]); v[7] = fma(md, rd, v[7]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = warp_sum(v[k]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k][tid >> 5] = v[k];
  }
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[tid][w];
    if (s != 0.0) atomicAdd(a.ws + (int64_t)r * SC_WS + tid, s);
  }
}

// cal_SISNR (score.py:7-21): t = <x~,r~> r~ / (|r~|^2 + eps);  20 log10(eps + |t| / (|x~ - t| + eps))
__device__ __forceinline__ double sisnr_closed(double Sx, double Sr, double Sxx, double Srr, double Sxr, double Ld) {
  const double xr = Sxr - Sx * Sr / Ld;
  const double rr = fmax(Srr - Sr * Sr / Ld, 0.0);
  const double xx = fmax(Sxx - Sx * Sx / Ld, 0.0);
  const double al = xr / (rr + SC_EPS);
  const double tt = al * al * rr;
  const double nn = fmax(xx - 2.0 * al * xr + tt, 0.0);
  return 20.0 * log10(SC_EPS + sqrt(tt) / (sqrt(nn) + SC_EPS));
}

__global__ void score_final_kernel(WesepScoreArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  const double* s = a.ws + (int64_t)r * SC_WS;
  const int len = a.len ? min(max(a.len[r], 0), a.L) : a.L;
  const double Ld = (double)max(len, 1);
  const double est = sisnr_closed(s[0], s[1], s[3], s[4], s[6], Ld);
  const double mix = sisnr_closed(s[2], s[1], s[5], s[4], s[7], Ld);
  a.sisnr[r] = (float)est;
  a.sisnri[r] = (float)(est - mix);
}

}  // namespace wb

using namespace wb;

extern "C" int64_t wesep_b200_score_ws_bytes(int n) { return (int64_t)(n > 0 ? n : 0) * SC_WS * sizeof(double); }

extern "C" int wesep_b200_score(const WesepScoreArgs* a, void* stream) {
  if (!a) return fail(-1, "score: null args");
  if (a->n <= 0 || a->L <= 0) return fail(-1, "score: bad shape");
  if (a->n > 65535) return fail(-1, "score: more than 65535 rows per call");
  if (!a->est || !a->ref || !a->mix || !a->ws || !a->sisnr || !a->sisnri) return fail(-1, "score: null buffer");
  if (a->ld_est < a->L || a->ld_ref < a->L || a->ld_mix < a->L) return fail(-1, "score: row stride < L");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->ws, 0, (size_t)wesep_b200_score_ws_bytes(a->n), st));
  const dim3 grid(cdiv(a->L, SC_CHUNK), a->n);
  if (a->peak_norm) {
    score_peak_kernel<<<grid, 256, 0, st>>>(*a);
    WB_LAUNCH_CHECK("score_peak");
  }
  score_sums_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("score_sums");
  score_final_kernel<<<cdiv(a->n, 128), 128, 0, st>>>(*a);
  WB_LAUNCH_CHECK("score_final");
  return 0;
}
