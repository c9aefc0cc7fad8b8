// GEMM kernels' host dispatch + the weight-gradient (time-contraction) GEMM.
#include "gemm_mma.cuh"

namespace wb {

int launch_gemm_wx(const GemmWxP& p, bool a_trans, int pro, int epi, cudaStream_t st) {
  if (p.n <= 0 || p.M <= 0 || p.Kd <= 0 || p.T <= 0) return fail(-1, "gemm_wx: empty shape");
  if (eff_gemm_backend(p.backend_sel) == 1 && eff_gemm_mode(p.mode_sel) == 0 && p.ws && (size_t)p.ws_bytes >= gemm_wx_tc_ws_bytes(p.M, p.Kd) &&
      gemm_wx_tc_eligible(p, pro, epi))
    return launch_gemm_wx_tc(p, a_trans, pro, epi, p.ws, st);
  if ((p.ldx & 3) || (p.ldw & 3) || !aligned16(p.X) || !aligned16(p.W)) return fail(-1, "gemm_wx: ldx/ldw/base alignment");
  if ((p.Kd & 3) || (a_trans && (p.M & 3))) return fail(-1, "gemm_wx: Kd (and M when transposed) must be multiples of 4");
  if (p.ep.ldy & 1) return fail(-1, "gemm_wx: ldy must be even");
  if (pro >= 2 && p.Kd > G_MAXK) return fail(-2, "gemm_wx: Kd too large for per-channel prologue");
#define WB_CASE(AT, PRO, EPI) \
  if (a_trans == AT && pro == PRO && epi == EPI) return launch_gemm_wx_t<AT, PRO, EPI>(p, st);
  WB_CASE(false, 0, 0)
  WB_CASE(false, 0, 1)
  WB_CASE(false, 0, 2)
  WB_CASE(false, 0, 3)
  WB_CASE(false, 1, 0)
  WB_CASE(false, 2, 0)
  WB_CASE(false, 2, 2)
  WB_CASE(false, 3, 0)
  WB_CASE(true, 0, 0)
  WB_CASE(true, 0, 1)
  WB_CASE(true, 0, 2)
  WB_CASE(true, 0, 10)
  WB_CASE(true, 2, 0)
#undef WB_CASE
  return fail(-2, "gemm_wx: unsupported (w_trans, pro, epi) combination");
}

// ------------------------------------------------------------------------------------ gemm_dw
// C[M][N] += sum_t A[m][t] * fb(B[j][t]) for one row n and one time chunk per CTA.
template <int PRO, bool X3>
__global__ void __launch_bounds__(G_THREADS, 2) gemm_dw_kernel(const GemmDwP p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                              // [stage][128][20]
  float* Bs = smem + G_STAGES * G_BM * G_A_LD;   // [stage][128][20]
  __shared__ float sc[G_BN], sh[G_BN];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int tiles_n = (p.N + G_BN - 1) / G_BN;
  const int m0 = (blockIdx.x / tiles_n) * G_BM, n0 = (blockIdx.x % tiles_n) * G_BN;
  const int row = blockIdx.z;
  const int tbeg = blockIdx.y * p.t_chunk;
  const int tend = min(tbeg + p.t_chunk, p.T);
  const float* An = p.A + (int64_t)row * p.bsa;
  const float* Bn = p.B + (int64_t)row * p.bsb;

  float alpha = 1.f;
  if constexpr (PRO >= 1) alpha = p.xb.alpha ? __ldg(p.xb.alpha) : 1.f;
  if constexpr (PRO >= 2) {
    float mu = 0.f, r = 1.f;
    if (p.xb.row_stats) gln_mean_rstd(p.xb.row_stats + 2 * row, p.xb.count, p.xb.eps, mu, r);
    if (tid < G_BN) {
      int c = n0 + tid;
      float gm = (p.xb.ch_scale && c < p.N) ? __ldg(p.xb.ch_scale + c) : 1.f;
      float bt = (p.xb.ch_shift && c < p.N) ? __ldg(p.xb.ch_shift + c) : 0.f;
      sc[tid] = gm * r;
      sh[tid] = bt - gm * mu * r;
    }
    __syncthreads();
  }
  float csc[4], csh[4];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    csc[ni] = 1.f; csh[ni] = 0.f;
    if constexpr (PRO >= 2) { csc[ni] = sc[wn * 32 + ni * 8 + g]; csh[ni] = sh[wn * 32 + ni * 8 + g]; }
  }

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  auto load_tiles = [&](int s, int tk) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int c = tid + i * G_THREADS;
      int r = c >> 2, kc = (c & 3) * 4;
      bool oka = (m0 + r < p.M) && (tk + kc < tend);
      const float* sa = oka ? An + (int64_t)(m0 + r) * p.lda + tk + kc : An;
      cp_async16(As + s * G_BM * G_A_LD + r * G_A_LD + kc, sa, oka);
      bool okb = (n0 + r < p.N) && (tk + kc < tend);
      const float* sb = okb ? Bn + (int64_t)(n0 + r) * p.ldb + tk + kc : Bn;
      cp_async16(Bs + s * G_BN * G_A_LD + r * G_A_LD + kc, sb, okb);
    }
  };

  const int KT = (tend - tbeg + G_BK - 1) / G_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < KT) load_tiles(s, tbeg + s * G_BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + G_STAGES - 1;
      if (nk < KT) load_tiles(nk % G_STAGES, tbeg + nk * G_BK);
      cp_async_commit();
    }
    const float* a_s = As + (kt % G_STAGES) * G_BM * G_A_LD;
    const float* b_s = Bs + (kt % G_STAGES) * G_BN * G_A_LD;
    const int tk = tbeg + kt * G_BK;
#pragma unroll
    for (int kk = 0; kk < G_BK; kk += 8) {
      // time indices of this thread's two k positions; elements at t >= tend are padding/garbage -> 0
      const bool k0ok = tk + kk + tig < tend, k1ok = tk + kk + tig + 4 < tend;
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        int col = wn * 32 + ni * 8 + g;
        float x0 = b_s[col * G_A_LD + kk + tig];
        float x1 = b_s[col * G_A_LD + kk + tig + 4];
        if constexpr (PRO == 1) { x0 = prelu_f(x0, alpha); x1 = prelu_f(x1, alpha); }
        if constexpr (PRO == 2) { x0 = fmaf(csc[ni], prelu_f(x0, alpha), csh[ni]); x1 = fmaf(csc[ni], prelu_f(x1, alpha), csh[ni]); }
        if constexpr (PRO == 3) { x0 = prelu_f(fmaf(csc[ni], x0, csh[ni]), alpha); x1 = prelu_f(fmaf(csc[ni], x1, csh[ni]), alpha); }
        x0 = k0ok ? x0 : 0.f;
        x1 = k1ok ? x1 : 0.f;
        if constexpr (X3) { split_tf32(x0, bh[ni][0], bl[ni][0]); split_tf32(x1, bh[ni][1], bl[ni][1]); }
        else { bh[ni][0] = f2tf32(x0); bh[ni][1] = f2tf32(x1); }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        int r = wm * 64 + mi * 16 + g;
        float a[4];
        a[0] = a_s[r * G_A_LD + kk + tig];
        a[1] = a_s[(r + 8) * G_A_LD + kk + tig];
        a[2] = a_s[r * G_A_LD + kk + tig + 4];
        a[3] = a_s[(r + 8) * G_A_LD + kk + tig + 4];
        a[0] = k0ok ? a[0] : 0.f; a[1] = k0ok ? a[1] : 0.f;
        a[2] = k1ok ? a[2] : 0.f; a[3] = k1ok ? a[3] : 0.f;
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (X3) split_tf32(a[q], ah[q], al[q]);
          else ah[q] = f2tf32(a[q]);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if constexpr (X3) { mma_tf32(acc[mi][ni], al, bh[ni]); mma_tf32(acc[mi][ni], ah, bl[ni]); }
          mma_tf32(acc[mi][ni], ah, bh[ni]);
        }
      }
    }
  }
  cp_async_wait<0>();

  float* C = p.C + (p.per_row ? (int64_t)row * p.M * p.ldc : 0);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int m = m0 + wm * 64 + mi * 16 + g + half * 8;
      if (m >= p.M) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        int c = n0 + wn * 32 + ni * 8 + 2 * tig;
        if (c < p.N) atomicAdd(C + (int64_t)m * p.ldc + c, acc[mi][ni][half * 2]);
        if (c + 1 < p.N) atomicAdd(C + (int64_t)m * p.ldc + c + 1, acc[mi][ni][half * 2 + 1]);
      }
    }
}

template <int PRO>
static int launch_gemm_dw_t(const GemmDwP& p, cudaStream_t st) {
  dim3 grid(cdiv(p.M, G_BM) * cdiv(p.N, G_BN), cdiv(p.T, p.t_chunk), p.n);
  size_t smem = (size_t)G_STAGES * (G_BM + G_BN) * G_A_LD * sizeof(float);
  if (eff_gemm_mode(p.mode_sel) == 0) {
    auto k = gemm_dw_kernel<PRO, true>;
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, G_THREADS, smem, st>>>(p);
  } else {
    auto k = gemm_dw_kernel<PRO, false>;
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, G_THREADS, smem, st>>>(p);
  }
  WB_LAUNCH_CHECK("gemm_dw");
  return 0;
}

bool gemm_dw_uses_tc(const GemmDwP& p, int pro_b) {
  return eff_gemm_backend(p.backend_sel) == 1 && eff_gemm_mode(p.mode_sel) == 0 && gemm_dw_tc_eligible(p, pro_b);
}

int launch_gemm_dw(const GemmDwP& pin, int pro_b, cudaStream_t st) {
  GemmDwP p = pin;
  if (p.n <= 0 || p.M <= 0 || p.N <= 0 || p.T <= 0) return fail(-1, "gemm_dw: empty shape");
  if ((p.lda & 3) || (p.ldb & 3) || !aligned16(p.A) || !aligned16(p.B)) return fail(-1, "gemm_dw: lda/ldb/base alignment");
  if (p.ldc < p.N) return fail(-1, "gemm_dw: ldc < N");
  if (gemm_dw_uses_tc(p, pro_b)) return launch_gemm_dw_tc(p, pro_b, st);
  if (p.a_rowsum) return fail(-2, "gemm_dw: a_rowsum is only produced by the tcgen05 path");
  if (p.t_chunk <= 0) {
    // aim for >= ~2 waves of CTAs (296 resident) without making chunks tiny
    int tiles = cdiv(p.M, G_BM) * cdiv(p.N, G_BN) * p.n;
    int want = (600 + tiles - 1) / tiles;
    int ch = cdiv(p.T, want < 1 ? 1 : want);
    ch = ((ch + G_BK - 1) / G_BK) * G_BK;
    if (ch < 256) ch = 256;
    p.t_chunk = ch;
  }
  if (p.t_chunk % G_BK) return fail(-1, "gemm_dw: t_chunk must be a multiple of 16");
  if (pro_b == 0) return launch_gemm_dw_t<0>(p, st);
  if (pro_b == 1) return launch_gemm_dw_t<1>(p, st);
  if (pro_b == 2) return launch_gemm_dw_t<2>(p, st);
  if (pro_b == 3) return launch_gemm_dw_t<3>(p, st);
  return fail(-2, "gemm_dw: unsupported prologue");
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_conv1x1(const WesepGemmArgs* a, void* stream) {
  GemmWxP p{};
  p.n = a->n; p.M = a->M; p.Kd = a->Kd; p.T = a->T;
  p.W = a->W; p.ldw = a->ldw;
  p.X = a->X; p.ldx = a->ldx; p.bsx = a->bsx ? a->bsx : (int64_t)a->Kd * a->ldx;
  p.xf = XformP{a->alpha, a->ch_scale, a->ch_shift, a->row_stats, a->stat_count, a->stat_eps};
  EpiP& e = p.ep;
  e.Y = a->Y; e.ldy = a->ldy; e.bsy = a->bsy ? a->bsy : (int64_t)a->M * a->ldy;
  e.bias = a->bias; e.row_bias = a->row_bias;
  e.R = a->R; e.ldr = a->ldr; e.bsr = a->bsr ? a->bsr : (int64_t)a->M * a->ldr;
  e.Y2 = a->Y2; e.ldy2 = a->ldy2; e.bsy2 = a->bsy2 ? a->bsy2 : (int64_t)a->M * a->ldy2;
  e.out_stats = a->out_stats; e.out_alpha = a->out_alpha; e.ch_stats = a->ch_stats;
  p.ws = a->ws; p.ws_bytes = a->ws_bytes;
  if (a->mode_sel < 0 || a->mode_sel > 2 || a->backend_sel < 0 || a->backend_sel > 2) return fail(-2, "conv1x1: mode_sel / backend_sel must be 0..2");
  p.mode_sel = a->mode_sel; p.backend_sel = a->backend_sel;
  if (a->epi < 0 || a->epi > 3) return fail(-2, "conv1x1: epi must be 0..3");
  if ((a->epi == 2 || a->epi == 3) && (!a->R || (a->ldr & 1))) return fail(-1, "conv1x1: residual/aux missing or odd ldr");
  if (a->epi == 3 && (!a->Y2 || (a->ldy2 & 1))) return fail(-1, "conv1x1: Y2 missing");
  return launch_gemm_wx(p, a->w_trans != 0, a->pro, a->epi, (cudaStream_t)stream);
}

extern "C" int wesep_b200_conv1x1_dw(const WesepGemmDwArgs* a, void* stream) {
  GemmDwP p{};
  p.n = a->n; p.M = a->M; p.N = a->N; p.T = a->T;
  p.A = a->A; p.lda = a->lda; p.bsa = a->bsa ? a->bsa : (int64_t)a->M * a->lda;
  p.B = a->B; p.ldb = a->ldb; p.bsb = a->bsb ? a->bsb : (int64_t)a->N * a->ldb;
  p.C = a->C; p.ldc = a->ldc ? a->ldc : a->N; p.per_row = a->per_row;
  p.xb = XformP{a->alpha_b, a->ch_scale_b, a->ch_shift_b, a->row_stats_b, a->stat_count, a->stat_eps};
  p.t_chunk = 0;
  if (a->mode_sel < 0 || a->mode_sel > 2 || a->backend_sel < 0 || a->backend_sel > 2) return fail(-2, "conv1x1_dw: mode_sel / backend_sel must be 0..2");
  p.mode_sel = a->mode_sel; p.backend_sel = a->backend_sel;
  return launch_gemm_dw(p, a->pro_b, (cudaStream_t)stream);
}
