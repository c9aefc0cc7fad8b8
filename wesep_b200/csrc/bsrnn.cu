// pBSRNN building blocks (wesep/models/bsrnn.py): layout changes, the LSTM cell and the mask-head elementwise ops.
// The GEMM-shaped parts of the model (input / recurrent / output projections of ResRNN, band split, mask MLPs,
// STFT / iSTFT as windowed-DFT GEMMs) run on the conv1x1 kernels of gemm_tc.cu; these kernels are the HBM-bound glue.
#include "gemm_mma.cuh"

namespace wb {

// ------------------------------------------------------------------------------------ swap-outer-inner transpose
// in [nb][Q][C][ld_in] (S valid columns) -> out [nb][S][C][ld_out] (Q valid columns): out[b][s][c][q] = in[b][q][c][s]
// (+ res[b][s][c][q]).  Used for: ResRNN reference layout <-> time-major layout, and BSNet's permute(0,3,2,1).
// 32 x 32 tiles through shared memory: both the read (along s) and the write (along q) are coalesced.
__global__ void __launch_bounds__(256) swap_oi_kernel(WesepTransposeArgs a) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int s0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  const int c = blockIdx.z % a.C, b = blockIdx.z / a.C;
  const float* in = a.in + (int64_t)b * a.Q * a.C * a.ld_in + (int64_t)c * a.ld_in;
  const int64_t in_q = (int64_t)a.C * a.ld_in;
  float* out = a.out + (int64_t)b * a.S * a.C * a.ld_out + (int64_t)c * a.ld_out;
  const float* res = a.res ? a.res + (int64_t)b * a.S * a.C * a.ld_out + (int64_t)c * a.ld_out : nullptr;
  const int64_t out_s = (int64_t)a.C * a.ld_out;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = q0 + ty + 8 * j, s = s0 + tx;
    tile[ty + 8 * j][tx] = (q < a.Q && s < a.S) ? __ldg(in + q * in_q + s) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = s0 + ty + 8 * j, q = q0 + tx;
    if (s < a.S && q < a.Q) {
      float v = tile[tx][ty + 8 * j];
      if (res) v += __ldg(res + s * out_s + q);
      out[s * out_s + q] = v;
    }
  }
}

// ------------------------------------------------------------------------------------ LSTM cell (time-major step)
// One time step of one direction for all Q sequences: G [4Hd][ld] holds the gate pre-activations (rows i | f | g | o in
// blocks of Hd, nn.LSTM order) and is overwritten by the activations (saved for the backward); c_prev may be NULL (zeros).
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(WesepLstmCellArgs a) {
  const int q = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (q >= a.Q) return;
  float* gi = a.G + (int64_t)j * a.ld + q;
  float* gf = gi + (int64_t)a.Hd * a.ld;
  float* gg = gf + (int64_t)a.Hd * a.ld;
  float* go = gg + (int64_t)a.Hd * a.ld;
  const float4 vi = *reinterpret_cast<const float4*>(gi), vf = *reinterpret_cast<const float4*>(gf);
  const float4 vg = *reinterpret_cast<const float4*>(gg), vo = *reinterpret_cast<const float4*>(go);
  float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.c_prev) cp = *reinterpret_cast<const float4*>(a.c_prev + (int64_t)j * a.ld + q);
  const float xi[4] = {vi.x, vi.y, vi.z, vi.w}, xf[4] = {vf.x, vf.y, vf.z, vf.w}, xg[4] = {vg.x, vg.y, vg.z, vg.w},
              xo[4] = {vo.x, vo.y, vo.z, vo.w}, xc[4] = {cp.x, cp.y, cp.z, cp.w};
  float oi[4], of[4], og[4], oo[4], oc[4], oh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool ok = q + k < a.Q;             // pad columns hold arbitrary bits: write zeros there
    const float i_ = sigmoid_f(xi[k]), f_ = sigmoid_f(xf[k]), g_ = tanhf(xg[k]), o_ = sigmoid_f(xo[k]);
    const float c_ = fmaf(f_, xc[k], i_ * g_);
    oi[k] = ok ? i_ : 0.f; of[k] = ok ? f_ : 0.f; og[k] = ok ? g_ : 0.f; oo[k] = ok ? o_ : 0.f;
    oc[k] = ok ? c_ : 0.f;
    oh[k] = ok ? o_ * tanhf(c_) : 0.f;
  }
  *reinterpret_cast<float4*>(gi) = make_float4(oi[0], oi[1], oi[2], oi[3]);
  *reinterpret_cast<float4*>(gf) = make_float4(of[0], of[1], of[2], of[3]);
  *reinterpret_cast<float4*>(gg) = make_float4(og[0], og[1], og[2], og[3]);
  *reinterpret_cast<float4*>(go) = make_float4(oo[0], oo[1], oo[2], oo[3]);
  *reinterpret_cast<float4*>(a.c + (int64_t)j * a.ld + q) = make_float4(oc[0], oc[1], oc[2], oc[3]);
  *reinterpret_cast<float4*>(a.h + (int64_t)j * a.ld + q) = make_float4(oh[0], oh[1], oh[2], oh[3]);
}

// Backward of one step: G holds the saved activations (i, f, g, o) and is overwritten by d(pre-activations);
// dh = dL/dh_s (already including the recurrent contribution), dc_in = dL/dc_s from step s+1 (NULL = 0).
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(WesepLstmCellArgs a) {
  const int q = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (q >= a.Q) return;
  float* gi = a.G + (int64_t)j * a.ld + q;
  float* gf = gi + (int64_t)a.Hd * a.ld;
  float* gg = gf + (int64_t)a.Hd * a.ld;
  float* go = gg + (int64_t)a.Hd * a.ld;
  const int64_t o = (int64_t)j * a.ld + q;
  const float4 vi = *reinterpret_cast<const float4*>(gi), vf = *reinterpret_cast<const float4*>(gf);
  const float4 vg = *reinterpret_cast<const float4*>(gg), vo = *reinterpret_cast<const float4*>(go);
  const float4 vc = *reinterpret_cast<const float4*>(a.c + o), vdh = *reinterpret_cast<const float4*>(a.dh + o);
  float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), dci = cp;
  if (a.c_prev) cp = *reinterpret_cast<const float4*>(a.c_prev + o);
  if (a.dc_in) dci = *reinterpret_cast<const float4*>(a.dc_in + o);
  const float xi[4] = {vi.x, vi.y, vi.z, vi.w}, xf[4] = {vf.x, vf.y, vf.z, vf.w}, xg[4] = {vg.x, vg.y, vg.z, vg.w},
              xo[4] = {vo.x, vo.y, vo.z, vo.w}, xc[4] = {vc.x, vc.y, vc.z, vc.w}, xp[4] = {cp.x, cp.y, cp.z, cp.w},
              xdh[4] = {vdh.x, vdh.y, vdh.z, vdh.w}, xdc[4] = {dci.x, dci.y, dci.z, dci.w};
  float di[4], df[4], dg[4], dO[4], dcp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool ok = q + k < a.Q;
    const float tc = tanhf(xc[k]);
    const float dh = ok ? xdh[k] : 0.f;
    const float dc = fmaf(dh * xo[k], 1.f - tc * tc, ok ? xdc[k] : 0.f);
    di[k] = ok ? dc * xg[k] * xi[k] * (1.f - xi[k]) : 0.f;
    df[k] = ok ? dc * xp[k] * xf[k] * (1.f - xf[k]) : 0.f;
    dg[k] = ok ? dc * xi[k] * (1.f - xg[k] * xg[k]) : 0.f;
    dO[k] = ok ? dh * tc * xo[k] * (1.f - xo[k]) : 0.f;
    dcp[k] = ok ? dc * xf[k] : 0.f;
  }
  *reinterpret_cast<float4*>(gi) = make_float4(di[0], di[1], di[2], di[3]);
  *reinterpret_cast<float4*>(gf) = make_float4(df[0], df[1], df[2], df[3]);
  *reinterpret_cast<float4*>(gg) = make_float4(dg[0], dg[1], dg[2], dg[3]);
  *reinterpret_cast<float4*>(go) = make_float4(dO[0], dO[1], dO[2], dO[3]);
  *reinterpret_cast<float4*>(a.dc_prev + o) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
}

// ------------------------------------------------------------------------------------ per-row channel affine
// y = ra[n][c] * x + rb[n][c]: SpeakerFuseLayer multiply / additive on the band-split feature (speaker.py:103-121),
// the per-row vector computed once by the Linear instead of at every (band, frame).  One CTA per (n, c) row.
template <bool BWD>
__global__ void __launch_bounds__(256) rowaffine_kernel(WesepRowAffineArgs a) {
  __shared__ float red[2 * 32];
  const int c = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const int64_t row = ((int64_t)n * a.C + c) * a.ld;
  const float sa = a.ra ? __ldg(a.ra + (int64_t)n * a.C + c) : 1.f;
  const float sb = a.rb ? __ldg(a.rb + (int64_t)n * a.C + c) : 0.f;
  float s0 = 0.f, s1 = 0.f;
  for (int t = 4 * tid; t < a.T; t += 1024) {
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(a.x + row + t));
    if constexpr (!BWD) {
      *reinterpret_cast<float4*>(a.y + row + t) = make_float4(fmaf(sa, x4.x, sb), fmaf(sa, x4.y, sb), fmaf(sa, x4.z, sb),
                                                              fmaf(sa, x4.w, sb));
    } else {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gy + row + t));
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k < a.T) { s0 = fmaf(gv[k], xv[k], s0); s1 += gv[k]; }
      *reinterpret_cast<float4*>(a.dx + row + t) = make_float4(sa * g4.x, sa * g4.y, sa * g4.z, sa * g4.w);
    }
  }
  if constexpr (BWD) {
    float v[2] = {s0, s1};
    block_sum<2>(v, red);
    if (tid == 0) {
      if (a.dra) a.dra[(int64_t)n * a.C + c] = v[0];
      if (a.drb) a.drb[(int64_t)n * a.C + c] = v[1];
    }
  }
}

// ------------------------------------------------------------------------------------ tanh
__global__ void __launch_bounds__(256) tanh_fwd_kernel(WesepTanhArgs a) {
  const int t = (blockIdx.y * 256 + threadIdx.x) * 4;
  if (t >= a.T) return;
  const int64_t o = (int64_t)blockIdx.x * a.ld + t;
  const float4 v = *reinterpret_cast<const float4*>(a.x + o);
  *reinterpret_cast<float4*>(a.y + o) = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
}
__global__ void __launch_bounds__(256) tanh_bwd_kernel(WesepTanhArgs a) {
  const int t = (blockIdx.y * 256 + threadIdx.x) * 4;
  if (t >= a.T) return;
  const int64_t o = (int64_t)blockIdx.x * a.ld + t;
  const float4 y = *reinterpret_cast<const float4*>(a.y + o), g = *reinterpret_cast<const float4*>(a.gy + o);
  *reinterpret_cast<float4*>(a.dx + o) = make_float4(g.x * (1.f - y.x * y.x), g.y * (1.f - y.y * y.y), g.z * (1.f - y.z * y.z),
                                                     g.w * (1.f - y.w * y.w));
}

// ------------------------------------------------------------------------------------ mask head tail (bsrnn.py:368-379)
// o [n][4*bw][ldo] = (value | gate) x (re | im) x bw;  m = value * sigmoid(gate);  est = mixture * m (complex product).
// s / e: band slices (bw re rows then bw im rows) of the mixture / estimate spectrograms.
template <bool BWD>
__global__ void __launch_bounds__(256) mask_apply_kernel(WesepMaskApplyArgs a) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int f = blockIdx.y, n = blockIdx.z;
  if (t >= a.T) return;
  const float* o = a.o + (int64_t)n * a.bso + t;
  const int64_t bw = a.bw;
  const float v_re = o[(0 * bw + f) * a.ldo], v_im = o[(1 * bw + f) * a.ldo];
  const float q_re = o[(2 * bw + f) * a.ldo], q_im = o[(3 * bw + f) * a.ldo];
  const float g_re = sigmoid_f(q_re), g_im = sigmoid_f(q_im);
  const float m_re = v_re * g_re, m_im = v_im * g_im;
  const float* s = a.s + (int64_t)n * a.bss + t;
  const float s_re = s[f * a.lds], s_im = s[(bw + f) * a.lds];
  if constexpr (!BWD) {
    float* e = a.e + (int64_t)n * a.bse + t;
    e[f * a.lde] = s_re * m_re - s_im * m_im;
    e[(bw + f) * a.lde] = s_re * m_im + s_im * m_re;
  } else {
    const float* ge = a.ge + (int64_t)n * a.bse + t;
    const float e_re = ge[f * a.lde], e_im = ge[(bw + f) * a.lde];
    const float dm_re = e_re * s_re + e_im * s_im, dm_im = e_im * s_re - e_re * s_im;
    float* go = a.go + (int64_t)n * a.bso + t;
    go[(0 * bw + f) * a.ldo] = dm_re * g_re;
    go[(1 * bw + f) * a.ldo] = dm_im * g_im;
    go[(2 * bw + f) * a.ldo] = dm_re * v_re * g_re * (1.f - g_re);
    go[(3 * bw + f) * a.ldo] = dm_im * v_im * g_im * (1.f - g_im);
  }
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_swap_outer_inner(const WesepTransposeArgs* a, void* stream) {
  if (a->nb <= 0 || a->Q <= 0 || a->C <= 0 || a->S <= 0) return fail(-1, "swap_outer_inner: empty shape");
  if (a->ld_in < a->S || a->ld_out < a->Q) return fail(-1, "swap_outer_inner: row strides");
  if ((int64_t)a->nb * a->C > 65535) return fail(-2, "swap_outer_inner: nb * C too large for one launch");
  swap_oi_kernel<<<dim3(cdiv(a->S, 32), cdiv(a->Q, 32), a->nb * a->C), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("swap_outer_inner");
  return 0;
}

static int check_cell(const WesepLstmCellArgs* a, bool bwd) {
  if (a->Hd <= 0 || a->Q <= 0) return fail(-1, "lstm_cell: empty shape");
  if ((a->ld & 3) || a->ld < a->Q || !aligned16(a->G) || !aligned16(a->c) || (a->c_prev && !aligned16(a->c_prev)))
    return fail(-1, "lstm_cell: ld must be a multiple of 4 and the rows 16-byte aligned");
  if (!bwd && !aligned16(a->h)) return fail(-1, "lstm_cell: h alignment");
  if (bwd && (!a->dh || !a->dc_prev || !aligned16(a->dh) || !aligned16(a->dc_prev) || (a->dc_in && !aligned16(a->dc_in))))
    return fail(-1, "lstm_cell_bwd: gradient buffers missing or misaligned");
  return 0;
}
extern "C" int wesep_b200_lstm_cell_fwd(const WesepLstmCellArgs* a, void* stream) {
  if (int rc = check_cell(a, false)) return rc;
  lstm_cell_fwd_kernel<<<dim3(cdiv(a->Q, 1024), a->Hd), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("lstm_cell_fwd");
  return 0;
}
extern "C" int wesep_b200_lstm_cell_bwd(const WesepLstmCellArgs* a, void* stream) {
  if (int rc = check_cell(a, true)) return rc;
  lstm_cell_bwd_kernel<<<dim3(cdiv(a->Q, 1024), a->Hd), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("lstm_cell_bwd");
  return 0;
}

extern "C" int wesep_b200_rowaffine_fwd(const WesepRowAffineArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "rowaffine: empty shape");
  if ((a->ld & 3) || !aligned16(a->x) || !aligned16(a->y)) return fail(-1, "rowaffine: alignment");
  rowaffine_kernel<false><<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("rowaffine_fwd");
  return 0;
}
extern "C" int wesep_b200_rowaffine_bwd(const WesepRowAffineArgs* a, void* stream) {
  if (a->n <= 0 || a->C <= 0 || a->T <= 0) return fail(-1, "rowaffine: empty shape");
  if ((a->ld & 3) || !aligned16(a->x) || !aligned16(a->gy) || !aligned16(a->dx)) return fail(-1, "rowaffine: alignment");
  rowaffine_kernel<true><<<dim3(a->C, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("rowaffine_bwd");
  return 0;
}
extern "C" int wesep_b200_tanh_fwd(const WesepTanhArgs* a, void* stream) {
  if (a->rows <= 0 || a->T <= 0 || (a->ld & 3) || !aligned16(a->x) || !aligned16(a->y)) return fail(-1, "tanh: shape / alignment");
  tanh_fwd_kernel<<<dim3((unsigned)a->rows, cdiv(a->T, 1024)), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("tanh_fwd");
  return 0;
}
extern "C" int wesep_b200_tanh_bwd(const WesepTanhArgs* a, void* stream) {
  if (a->rows <= 0 || a->T <= 0 || (a->ld & 3) || !aligned16(a->y) || !aligned16(a->gy) || !aligned16(a->dx))
    return fail(-1, "tanh_bwd: shape / alignment");
  tanh_bwd_kernel<<<dim3((unsigned)a->rows, cdiv(a->T, 1024)), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("tanh_bwd");
  return 0;
}
extern "C" int wesep_b200_mask_apply_fwd(const WesepMaskApplyArgs* a, void* stream) {
  if (a->n <= 0 || a->bw <= 0 || a->T <= 0 || !a->o || !a->s || !a->e) return fail(-1, "mask_apply: shape / pointers");
  mask_apply_kernel<false><<<dim3(cdiv(a->T, 256), a->bw, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("mask_apply_fwd");
  return 0;
}
extern "C" int wesep_b200_mask_apply_bwd(const WesepMaskApplyArgs* a, void* stream) {
  if (a->n <= 0 || a->bw <= 0 || a->T <= 0 || !a->o || !a->s || !a->ge || !a->go) return fail(-1, "mask_apply_bwd: shape / pointers");
  mask_apply_kernel<true><<<dim3(cdiv(a->T, 256), a->bw, a->n), 256, 0, (cudaStream_t)stream>>>(*a);
  WB_LAUNCH_CHECK("mask_apply_bwd");
  return 0;
}
