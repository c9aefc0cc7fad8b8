// Device-side data front end: the per-sample CPU work of the reference DataLoader workers
// moved to streaming kernels over whole batches.
//   * chunk + SNR mixing   — random_chunk / get_random_chunk (wesep/dataset/processor.py:536-573, the offsets are
//                            drawn by the host) and snr_mixer (processor.py:276-320)
//   * Kaldi fbank + CMN    — compute_fbank (processor.py:480-512, torchaudio.compliance.kaldi.fbank: dither,
//                            DC removal, pre-emphasis, Hamming window, 512-point power spectrum, 80 mel bins, log)
//                            and apply_cmvn (processor.py:515-535, mean only)
// All of it is HBM-bound byte-streaming work: no tensor cores here.
#include "common.cuh"

namespace wb {

// ------------------------------------------------------------------------------------------------ mixing
constexpr int MX_CHUNK = 4096;   // samples per CTA
constexpr int MX_MAXS = 4;
// ws per mixture (doubles): [0..S) chunk energies, [4] two u32 = bits of max|mix| and of max_s max|scaled spk_s|
constexpr int MX_WS = 5;

// sample j of the chunk of (mixture m, speaker s): utterances shorter than T tile (processor.py:562-570)
__device__ __forceinline__ float mx_src(const WesepMixArgs& a, int m, int s, int j) {
  const int q = m * a.S + s;
  const int ul = __ldg(a.ulen + q);
  int64_t i = (ul >= a.T) ? (int64_t)__ldg(a.chunk0 + q) + j : (int64_t)(j % ul);
  return __ldg(a.pool + __ldg(a.start + q) + i);
}

__global__ void __launch_bounds__(256) mix_energy_kernel(WesepMixArgs a) {
  __shared__ double red[MX_MAXS][8];
  const int m = blockIdx.y, c0 = blockIdx.x * MX_CHUNK, tid = threadIdx.x;
  const int end = min(c0 + MX_CHUNK, a.T);
  double e[MX_MAXS] = {0, 0, 0, 0};
  for (int j = c0 + tid; j < end; j += 256) {
#pragma unroll
    for (int s = 0; s < MX_MAXS; ++s)
      if (s < a.S) { const double v = (double)mx_src(a, m, s, j); e[s] = fma(v, v, e[s]); }
  }
#pragma unroll
  for (int s = 0; s < MX_MAXS; ++s) e[s] = warp_sum(e[s]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int s = 0; s < MX_MAXS; ++s) red[s][tid >> 5] = e[s];
  }
  __syncthreads();
  if (tid < a.S) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[tid][w];
    atomicAdd(a.ws + (int64_t)m * MX_WS + tid, v);
  }
}

// interference *= sqrt(target_energy / energy) * 10**(snr/20)   (processor.py:299-300, all fp32 tensors)
__device__ __forceinline__ void mx_gains(const WesepMixArgs& a, int m, float (&g)[MX_MAXS]) {
  const double* w = a.ws + (int64_t)m * MX_WS;
  const float te = (float)w[0];
  g[0] = 1.f;
#pragma unroll
  for (int s = 1; s < MX_MAXS; ++s) {
    g[s] = 0.f;
    if (s < a.S) {
      const float snr = a.snr_db ? __ldg(a.snr_db + m * a.S + s) : 0.f;
      const float k = (float)pow(10.0, (double)snr / 20.0);   // the reference's Python float, cast when it meets the tensor
      g[s] = __fmul_rn(__fsqrt_rn(__fdiv_rn(te, (float)w[s])), k);
    }
  }
}

template <bool WRITE>
__global__ void __launch_bounds__(256) mix_apply_kernel(WesepMixArgs a) {
  __shared__ float red[2][8];
  const int m = blockIdx.y, c0 = blockIdx.x * MX_CHUNK, tid = threadIdx.x;
  const int end = min(c0 + MX_CHUNK, a.T);
  float g[MX_MAXS];
  mx_gains(a, m, g);
  float scal = 1.f;
  if (WRITE) {
    const unsigned* pk = reinterpret_cast<const unsigned*>(a.ws + (int64_t)m * MX_WS + 4);
    const float amp = fmaxf(__uint_as_float(pk[0]), __uint_as_float(pk[1]));
    scal = amp != 0.f ? (float)(1.0 / (double)amp) : 1.f;     // mix_scaling = 1 / max_amp (Python float), processor.py:309-312
  }
  float mmax = 0.f, smax = 0.f;
  for (int j = c0 + tid; j < end; j += 256) {
    float v[MX_MAXS], mix = 0.f;
#pragma unroll
    for (int s = 0; s < MX_MAXS; ++s) {
      v[s] = 0.f;
      if (s < a.S) {
        v[s] = __fmul_rn(mx_src(a, m, s, j), g[s]);
        mix = s == 0 ? v[0] : __fadd_rn(mix, v[s]);              // torch.sum over the stacked speakers, in order
      }
    }
    if (WRITE) {
      a.mix[(int64_t)m * a.ld + j] = __fmul_rn(mix, scal);
#pragma unroll
      for (int s = 0; s < MX_MAXS; ++s)
        if (s < a.S) a.spk[((int64_t)s * a.M + m) * a.ld + j] = __fmul_rn(v[s], scal);
    } else {
      mmax = fmaxf(mmax, fabsf(mix));
#pragma unroll
      for (int s = 0; s < MX_MAXS; ++s) smax = fmaxf(smax, fabsf(v[s]));
    }
  }
  if (!WRITE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mmax = fmaxf(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = mmax; red[1][tid >> 5] = smax; }
    __syncthreads();
    if (tid < 2) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v = fmaxf(v, red[tid][w]);
      atomicMax(reinterpret_cast<unsigned*>(a.ws + (int64_t)m * MX_WS + 4) + tid, __float_as_uint(v));
    }
  }
}

// ------------------------------------------------------------------------------------------------ fbank
constexpr int FB_WARPS = 8;       // frames per CTA (one warp each)
constexpr int FB_NFFT = 512;
constexpr int FB_LOG2 = 9;
constexpr int FB_MAXMEL = 96;

// 64-bit mix (splitmix64) -> two uniforms -> one standard normal (Box-Muller); the dither only needs to be white
__device__ __forceinline__ float fb_randn(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u1 = ((float)(uint32_t)(z >> 40) + 1.f) * (1.f / 16777216.f);   // (0, 1]
  const float u2 = (float)(uint32_t)((z >> 8) & 0xFFFFFFu) * (1.f / 16777216.f);
  float sn, cs;
  sincospif(2.f * u2, &sn, &cs);
  return sqrtf(-2.f * logf(u1)) * cs;
}

__global__ void __launch_bounds__(FB_WARPS * 32) fbank_kernel(WesepFbankArgs a) {
  __shared__ float s_re[FB_WARPS][FB_NFFT];
  __shared__ float s_im[FB_WARPS][FB_NFFT];
  __shared__ float s_tw[2][FB_NFFT / 2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < FB_NFFT / 2; j += FB_WARPS * 32) {
    float sn, cs;
    sincospif((float)j * (2.f / FB_NFFT), &sn, &cs);
    s_tw[0][j] = cs;
    s_tw[1][j] = -sn;
  }
  __syncthreads();
  const int r = blockIdx.y;
  const int len = a.len ? min(max(__ldg(a.len + r), 0), a.T) : a.T;
  const int nfr = len >= a.frame_len ? 1 + (len - a.frame_len) / a.frame_shift : 0;
  const int f = blockIdx.x * FB_WARPS + warp;
  if (f >= a.max_frames) return;
  float* out = a.out + (int64_t)r * a.bs_out + (int64_t)f * a.num_mel;
  if (f >= nfr) {                                   // padding frames of a shorter utterance
    for (int m = lane; m < a.num_mel; m += 32) out[m] = 0.f;
    return;
  }
  float* re = s_re[warp];
  float* im = s_im[warp];
  const float* x = a.wav + (int64_t)r * a.ld_wav + (int64_t)f * a.frame_shift;
  // 1. scale to int16 range, dither, DC removal (kaldi.py _get_window)
  constexpr int PER = (FB_NFFT + 31) / 32;
  float v[PER];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = lane + 32 * i;
    v[i] = 0.f;
    if (j < a.frame_len) {
      v[i] = __ldg(x + j) * a.scale;
      if (a.dither != 0.f) v[i] += a.dither * fb_randn(a.seed, ((uint64_t)r * a.max_frames + f) * FB_NFFT + j);
      sum += v[i];
    }
  }
  if (a.remove_dc) {
    const float mean = warp_sum(sum) / (float)a.frame_len;
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] -= mean;
  }
  // 2. pre-emphasis y[j] = x[j] - c x[max(j-1,0)], window, bit-reversed store, zero padding
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = lane + 32 * i;
    const float up = __shfl_up_sync(0xffffffffu, v[i], 1);
    const float wrap = __shfl_sync(0xffffffffu, i > 0 ? v[i > 0 ? i - 1 : 0] : 0.f, 31);
    const float prev = lane ? up : (i ? wrap : v[0]);
    float y = 0.f;
    if (j < a.frame_len) y = (v[i] - a.preemph * prev) * __ldg(a.window + j);
    const int br = __brev((unsigned)j) >> (32 - FB_LOG2);
    re[br] = y;
    im[br] = 0.f;
  }
  __syncwarp();
  // 3. radix-2 DIT FFT, 256 butterflies per stage over the warp
#pragma unroll
  for (int s = 0; s < FB_LOG2; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int q = 0; q < FB_NFFT / 64; ++q) {
      const int b = lane + 32 * q;
      const int j = b & (half - 1);
      const int i0 = ((b >> s) << (s + 1)) + j, i1 = i0 + half;
      const int t = j << (FB_LOG2 - 1 - s);
      const float wr = s_tw[0][t], wi = s_tw[1][t];
      const float br_ = re[i1], bi_ = im[i1];
      const float tr = br_ * wr - bi_ * wi, ti = br_ * wi + bi_ * wr;
      const float ar = re[i0], ai = im[i0];
      re[i0] = ar + tr; im[i0] = ai + ti;
      re[i1] = ar - tr; im[i1] = ai - ti;
    }
    __syncwarp();
  }
  // 4. power spectrum (bins 0..256) in place
#pragma unroll
  for (int i = 0; i < (FB_NFFT / 2 + 32) / 32; ++i) {
    const int k = lane + 32 * i;
    if (k <= FB_NFFT / 2) re[k] = re[k] * re[k] + im[k] * im[k];
  }
  __syncwarp();
  // 5. triangular mel filters (rows of a.mel are zero outside [lo, hi)), log floor
  for (int m = lane; m < a.num_mel; m += 32) {
    const float* w = a.mel + (int64_t)m * (FB_NFFT / 2 + 1);
    const int lo = __ldg(a.mel_lo + m), hi = __ldg(a.mel_hi + m);
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(re[k], __ldg(w + k), acc);
    out[m] = logf(fmaxf(acc, a.log_floor));
  }
}

// CMN: mat - mean(mat, dim=0) over the valid frames of each utterance (apply_cmvn, norm_mean only)
__global__ void __launch_bounds__(512) fbank_cmn_kernel(WesepFbankArgs a) {
  __shared__ double red[512];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int len = a.len ? min(max(__ldg(a.len + r), 0), a.T) : a.T;
  const int nfr = min(len >= a.frame_len ? 1 + (len - a.frame_len) / a.frame_shift : 0, a.max_frames);
  if (nfr == 0) return;
  float* out = a.out + (int64_t)r * a.bs_out;
  const int slices = 512 / a.num_mel;              // frame slices per mel bin
  const int m = tid % a.num_mel, sl = tid / a.num_mel;
  double acc = 0.0;
  if (sl < slices)
    for (int f = sl; f < nfr; f += slices) acc += (double)out[(int64_t)f * a.num_mel + m];
  red[tid] = acc;
  __syncthreads();
  if (tid < a.num_mel) {
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += red[k * a.num_mel + tid];
    red[tid] = s / (double)nfr;
  }
  __syncthreads();
  if (sl < slices) {
    const float mean = (float)red[m];
    for (int f = sl; f < nfr; f += slices) out[(int64_t)f * a.num_mel + m] -= mean;
  }
}

// ------------------------------------------------------------------------------------------------ "consistent" features
// The in-model enrollment features of the recipes with `spk_feat: False`, `feat_type: consistent`
// (wesep/models/bsrnn.py:231-241,343-351 and the same block in convtasnet / dpccn / tfgridnet): PreEmphasis
// (wesep/modules/common/speaker.py:10-23) -> torchaudio MelSpectrogram (hamming, power 2, HTK mel) + 1e-8 -> log -> minus the
// mean over frames -> [B, frames, n_mels].  The STFT and the mel projection are the DFT / pointwise GEMMs; these three
// kernels are the elementwise rest.
__global__ void __launch_bounds__(256) preemph_kernel(WesepPreEmphArgs a) {
  const int r = blockIdx.y;
  const float* x = a.x + (int64_t)r * a.ldx;
  float* y = a.y + (int64_t)r * a.ldy;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < a.L; t += gridDim.x * 256) {
    const float prev = t > 0 ? __ldg(x + t - 1) : __ldg(x + (a.L > 1 ? 1 : 0));      // F.pad(.., (1, 0), "reflect")
    y[t] = __fmaf_rn(-a.coef, prev, __ldg(x + t));
  }
}
// pw[n][f][t] = re^2 + im^2 from spec rows [0, F) (real) and [F, 2F) (imaginary)
__global__ void __launch_bounds__(256) power_spec_kernel(WesepPowerSpecArgs a) {
  const int f = blockIdx.y, n = blockIdx.z;
  const float* re = a.spec + (int64_t)n * a.bs + (int64_t)f * a.ld;
  const float* im = re + (int64_t)a.F * a.ld;
  float* pw = a.pw + (int64_t)n * a.bsp + (int64_t)f * a.ldp;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < a.T; t += gridDim.x * 256) {
    const float r = __ldg(re + t), i = __ldg(im + t);
    pw[t] = fmaf(r, r, i * i);
  }
}
// out[n][t][m] = log(mel[n][m][t] + eps) - mean_t log(mel[n][m][t] + eps)     (one CTA per (m, n))
__global__ void __launch_bounds__(256) log_cmn_kernel(WesepLogCmnArgs a) {
  __shared__ double red[8];
  __shared__ float s_mean;
  const int m = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const float* x = a.mel + ((int64_t)n * a.M + m) * a.ld;
  double s = 0.0;
  for (int t = tid; t < a.T; t += 256) s += (double)logf(__ldg(x + t) + a.eps);
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w];
    s_mean = (float)(v / (double)a.T);
  }
  __syncthreads();
  const float mean = s_mean;
  float* out = a.out + (int64_t)n * a.T * a.M + m;
  for (int t = tid; t < a.T; t += 256) out[(int64_t)t * a.M] = logf(__ldg(x + t) + a.eps) - mean;
}

}  // namespace wb

using namespace wb;

extern "C" int64_t wesep_b200_mix_ws_bytes(int M) { return (int64_t)(M > 0 ? M : 0) * MX_WS * sizeof(double); }

extern "C" int wesep_b200_mix(const WesepMixArgs* a, void* stream) {
  if (!a) return fail(-1, "mix: null args");
  if (a->M <= 0 || a->M > 65535 || a->T <= 0) return fail(-1, "mix: bad shape");
  if (a->S < 1 || a->S > MX_MAXS) return fail(-1, "mix: 1..4 speakers per mixture");
  if (!a->pool || !a->start || !a->ulen || !a->chunk0 || !a->mix || !a->spk || !a->ws) return fail(-1, "mix: null buffer");
  if (a->ld < a->T) return fail(-1, "mix: row stride < T");
  cudaStream_t st = (cudaStream_t)stream;
  WB_CUDA(cudaMemsetAsync(a->ws, 0, (size_t)wesep_b200_mix_ws_bytes(a->M), st));
  const dim3 grid(cdiv(a->T, MX_CHUNK), a->M);
  mix_energy_kernel<<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("mix_energy");
  mix_apply_kernel<false><<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("mix_peak");
  mix_apply_kernel<true><<<grid, 256, 0, st>>>(*a);
  WB_LAUNCH_CHECK("mix_write");
  return 0;
}

extern "C" int wesep_b200_fbank(const WesepFbankArgs* a, void* stream) {
  if (!a) return fail(-1, "fbank: null args");
  if (a->n <= 0 || a->n > 65535 || a->T <= 0 || a->max_frames <= 0) return fail(-1, "fbank: bad shape");
  if (a->n_fft != FB_NFFT) return fail(-1, "fbank: only the 512-point transform (25 ms @ 16 kHz rounded up) is built");
  if (a->frame_len <= 0 || a->frame_len > FB_NFFT || a->frame_shift <= 0) return fail(-1, "fbank: bad frame geometry");
  if (a->num_mel < 4 || a->num_mel > FB_MAXMEL) return fail(-1, "fbank: 4..96 mel bins");
  if (!a->wav || !a->window || !a->mel || !a->mel_lo || !a->mel_hi || !a->out) return fail(-1, "fbank: null buffer");
  if (a->bs_out < (int64_t)a->max_frames * a->num_mel) return fail(-1, "fbank: batch stride of out too small");
  cudaStream_t st = (cudaStream_t)stream;
  fbank_kernel<<<dim3(cdiv(a->max_frames, FB_WARPS), a->n), FB_WARPS * 32, 0, st>>>(*a);
  WB_LAUNCH_CHECK("fbank");
  if (a->cmn) {
    fbank_cmn_kernel<<<a->n, 512, 0, st>>>(*a);
    WB_LAUNCH_CHECK("fbank_cmn");
  }
  return 0;
}

extern "C" int wesep_b200_preemphasis(const WesepPreEmphArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->n > 65535 || a->L <= 0 || a->ldx < a->L || a->ldy < a->L || !a->x || !a->y) return fail(-1, "preemphasis: bad arguments");
  preemph_kernel<<<dim3(cdiv(a->L, 2048), a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("preemphasis");
  return 0;
}
extern "C" int wesep_b200_power_spec(const WesepPowerSpecArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->n > 65535 || a->F <= 0 || a->F > 65535 || a->T <= 0 || a->ld < a->T || a->ldp < a->T || !a->spec || !a->pw)
    return fail(-1, "power_spec: bad arguments");
  power_spec_kernel<<<dim3(cdiv(a->T, 1024), a->F, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("power_spec");
  return 0;
}
extern "C" int wesep_b200_log_cmn(const WesepLogCmnArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->n > 65535 || a->M <= 0 || a->T <= 0 || a->ld < a->T || !a->mel || !a->out) return fail(-1, "log_cmn: bad arguments");
  log_cmn_kernel<<<dim3(a->M, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("log_cmn");
  return 0;
}
