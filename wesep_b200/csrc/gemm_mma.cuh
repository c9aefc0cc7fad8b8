// Pointwise-conv GEMMs on the legacy tensor path (mma.sync m16n8k8 TF32, 3xTF32 split for
// fp32-grade accuracy), cp.async 4-stage pipeline, operand prologue + fused epilogues.
//
//   gemm_wx :  Y[n][M][T]  = W[M][Kd] * f(X[n][Kd][T])     (forward 1x1 convs and dX GEMMs)
//   gemm_dw :  C[M][N]    += sum_t fa(A[n][M][t]) * fb(B[n][N][t])   (weight gradients)
//
// CTA tile 128x128, 8 warps (2 x 4), warp tile 64x32, BK = 16.
#pragma once
#include "common.cuh"

namespace wb {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_STAGES = 4, G_THREADS = 256;
constexpr int G_A_LD = G_BK + 4;    // As[m][k]   (row-major weights)
constexpr int G_AT_LD = G_BM + 8;   // AsT[k][m]  (transposed weights)
constexpr int G_B_LD = G_BN + 8;    // Bs[k][t]
constexpr int G_A_TILE = (G_BM * G_A_LD > G_BK * G_AT_LD) ? G_BM * G_A_LD : G_BK * G_AT_LD;
constexpr int G_B_TILE = G_BK * G_B_LD;
constexpr int G_MAXK = 1024;        // max contraction length with a per-channel prologue

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------------ parameters
struct XformP {             // operand prologue: v = sc[c] * prelu(x, alpha) + sh[c]
  const float* alpha;       // device scalar or nullptr (=1, identity)
  const float* ch_scale;    // [C] or nullptr (=1)
  const float* ch_shift;    // [C] or nullptr (=0)
  const double* row_stats;  // [n][2] gLN sums or nullptr
  double count;
  float eps;
};

struct EpiP {
  float* Y; int64_t ldy; int64_t bsy;          // [n][M][ldy]
  const float* bias; const float* row_bias;    // [M], [n][M]
  const float* R; int64_t ldr; int64_t bsr;    // residual / aux
  float* Y2; int64_t ldy2; int64_t bsy2;
  double* out_stats; const float* out_alpha;   // [n][2]
  double* ch_stats;                            // [M][2]
  // --- TCN backward "B2" epilogue (epi 10) ---
  const float* d; int64_t ldd; int64_t bsd;    // pre-activation d [n][H][ld]
  const float* a2; const float* g2;            // PReLU_2 slope, gLN2 weight [H]
  const double* stats2; double count2; float eps2;
  const double* rowsc;                         // [n][8]: [0]=mean(h) [1]=mean(h*yhat2) (inputs)
  double* rowacc;                              // [n][8]: [2]=P1 [3]=P2 [4]=P3 [5]=dalpha2 (outputs, +=)
  const float* g1; const float* be1; const float* bd; const float* wd;  // [H],[H],[H],[H][3]
  int dil;
};

struct GemmWxP {
  int n, M, Kd, T;
  const float* W; int64_t ldw;
  const float* X; int64_t ldx; int64_t bsx;
  XformP xf;
  EpiP ep;
  void* ws; int64_t ws_bytes;   // optional workspace for the tcgen05 backend
  int ws_presplit;              // tcgen05 backend: `ws` already holds this weight's hi / lo tiles (skip the split launch)
  int mode_sel, backend_sel;    // per-call precision mode / backend: 0 = the process default, otherwise value + 1
};

// tcgen05 backend (gemm_tc.cu)
bool gemm_wx_tc_eligible(const GemmWxP& p, int pro, int epi);
size_t gemm_wx_tc_ws_bytes(int M, int Kd);
int launch_gemm_wx_tc(const GemmWxP& p, bool a_trans, int pro, int epi, void* ws, cudaStream_t st);
extern int g_gemm_backend;
inline int eff_gemm_mode(int sel) { return sel > 0 ? sel - 1 : g_gemm_mode; }
inline int eff_gemm_backend(int sel) { return sel > 0 ? sel - 1 : g_gemm_backend; }

// ------------------------------------------------------------------------------------ epilogues
template <int EPI>
struct EpiState {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
};

template <int EPI>
__device__ __forceinline__ void epi_store2(float* p, float v0, float v1, bool ok1) {
  if (ok1) {
    *reinterpret_cast<float2*>(p) = make_float2(v0, v1);
  } else {
    p[0] = v0;
  }
}

// Called once per (row m, column pair t,t+1) accumulator pair. t < T guaranteed, ok1 = (t+1 < T).
template <int EPI>
__device__ __forceinline__ void epi_apply(const GemmWxP& p, EpiState<EPI>& st, int n, int m, int t, float v0, float v1,
                                          bool ok1, float bias_m, float& chs, float& chq) {
  const EpiP& e = p.ep;
  v0 += bias_m;
  v1 += bias_m;
  if constexpr (EPI == 0) {
    epi_store2<EPI>(e.Y + n * e.bsy + (int64_t)m * e.ldy + t, v0, v1, ok1);
    if (e.out_stats) {
      float a = e.out_alpha ? __ldg(e.out_alpha) : 1.f;
      float y0 = prelu_f(v0, a), y1 = ok1 ? prelu_f(v1, a) : 0.f;
      st.s0 += y0 + y1;
      st.s1 += y0 * y0 + y1 * y1;
    }
    if (e.ch_stats) {
      float w1 = ok1 ? v1 : 0.f;
      chs += v0 + w1;
      chq += v0 * v0 + w1 * w1;
    }
  } else if constexpr (EPI == 1) {
    epi_store2<EPI>(e.Y + n * e.bsy + (int64_t)m * e.ldy + t, fmaxf(v0, 0.f), fmaxf(v1, 0.f), ok1);
  } else if constexpr (EPI == 2) {
    const float* r = e.R + n * e.bsr + (int64_t)m * e.ldr + t;
    float r0 = r[0], r1 = ok1 ? r[1] : 0.f;
    epi_store2<EPI>(e.Y + n * e.bsy + (int64_t)m * e.ldy + t, v0 + r0, v1 + r1, ok1);
  } else if constexpr (EPI == 3) {
    const float* r = e.R + n * e.bsr + (int64_t)m * e.ldr + t;
    float r0 = r[0], r1 = ok1 ? r[1] : 0.f;
    float m0 = fmaxf(v0, 0.f), m1 = fmaxf(v1, 0.f);
    epi_store2<EPI>(e.Y2 + n * e.bsy2 + (int64_t)m * e.ldy2 + t, m0, m1, ok1);
    epi_store2<EPI>(e.Y + n * e.bsy + (int64_t)m * e.ldy + t, r0 * m0, r1 * m1, ok1);
  }
}

// ------------------------------------------------------------------------------------ gemm_wx
template <bool A_TRANS>
__device__ __forceinline__ void load_a_tile(float* As, const GemmWxP& p, int m0, int k0, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int c = tid + i * G_THREADS;  // 512 chunks of 16 B
    if constexpr (!A_TRANS) {
      int row = c >> 2, kc = (c & 3) * 4;
      bool ok = (m0 + row < p.M) && (k0 + kc < p.Kd);
      const float* src = ok ? p.W + (int64_t)(m0 + row) * p.ldw + k0 + kc : p.W;
      cp_async16(As + row * G_A_LD + kc, src, ok);
    } else {
      int krow = c >> 5, mc = (c & 31) * 4;
      bool ok = (k0 + krow < p.Kd) && (m0 + mc < p.M);
      const float* src = ok ? p.W + (int64_t)(k0 + krow) * p.ldw + m0 + mc : p.W;
      cp_async16(As + krow * G_AT_LD + mc, src, ok);
    }
  }
}
__device__ __forceinline__ void load_b_tile(float* Bs, const float* Xn, int64_t ldx, int Kd, int T, int k0, int t0,
                                            int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int c = tid + i * G_THREADS;
    int krow = c >> 5, tc = (c & 31) * 4;
    bool ok = (k0 + krow < Kd) && (t0 + tc < T);
    const float* src = ok ? Xn + (int64_t)(k0 + krow) * ldx + t0 + tc : Xn;
    cp_async16(Bs + krow * G_B_LD + tc, src, ok);
  }
}

// PRO: 0 identity, 1 prelu, 2 sc*prelu(x)+sh (PReLU then gLN apply), 3 prelu(sc*x+sh) (BatchNorm apply then PReLU)
template <bool A_TRANS, int PRO, int EPI, bool X3>
__global__ void __launch_bounds__(G_THREADS, 2) gemm_wx_kernel(const GemmWxP p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + G_STAGES * G_A_TILE;
  float* sc = Bs + G_STAGES * G_B_TILE;  // [Kd] (PRO 2)
  float* sh = sc + (PRO >= 2 ? p.Kd : 0);
  __shared__ float red[4 * 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int t0 = blockIdx.x * G_BN, m0 = blockIdx.y * G_BM, n = blockIdx.z;
  const float* Xn = p.X + (int64_t)n * p.bsx;

  float alpha = 1.f;
  if constexpr (PRO >= 1) alpha = p.xf.alpha ? __ldg(p.xf.alpha) : 1.f;
  if constexpr (PRO >= 2) {
    float mu = 0.f, r = 1.f;
    if (p.xf.row_stats) gln_mean_rstd(p.xf.row_stats + 2 * n, p.xf.count, p.xf.eps, mu, r);
    for (int k = tid; k < p.Kd; k += G_THREADS) {
      float gm = p.xf.ch_scale ? __ldg(p.xf.ch_scale + k) : 1.f;
      float bt = p.xf.ch_shift ? __ldg(p.xf.ch_shift + k) : 0.f;
      sc[k] = gm * r;
      sh[k] = bt - gm * mu * r;
    }
  }

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  const int KT = (p.Kd + G_BK - 1) / G_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < KT) {
      load_a_tile<A_TRANS>(As + s * G_A_TILE, p, m0, s * G_BK, tid);
      load_b_tile(Bs + s * G_B_TILE, Xn, p.ldx, p.Kd, p.T, s * G_BK, t0, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + G_STAGES - 1;
      if (nk < KT) {
        int s = nk % G_STAGES;
        load_a_tile<A_TRANS>(As + s * G_A_TILE, p, m0, nk * G_BK, tid);
        load_b_tile(Bs + s * G_B_TILE, Xn, p.ldx, p.Kd, p.T, nk * G_BK, t0, tid);
      }
      cp_async_commit();
    }
    const float* a_s = As + (kt % G_STAGES) * G_A_TILE;
    const float* b_s = Bs + (kt % G_STAGES) * G_B_TILE;
#pragma unroll
    for (int kk = 0; kk < G_BK; kk += 8) {
      // ---- B fragments (with prologue), all 4 n-frags
      uint32_t bh[4][2], bl[4][2];
      const int kg0 = kt * G_BK + kk + tig, kg1 = kg0 + 4;
      float c0 = 1.f, d0 = 0.f, c1 = 1.f, d1 = 0.f;
      if constexpr (PRO >= 2) {
        // rows beyond Kd hold zero-filled X; force their transformed value to 0 as well
        c0 = kg0 < p.Kd ? sc[kg0] : 0.f; d0 = kg0 < p.Kd ? sh[kg0] : 0.f;
        c1 = kg1 < p.Kd ? sc[kg1] : 0.f; d1 = kg1 < p.Kd ? sh[kg1] : 0.f;
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        int col = wn * 32 + ni * 8 + g;
        float x0 = b_s[(kk + tig) * G_B_LD + col];
        float x1 = b_s[(kk + tig + 4) * G_B_LD + col];
        if constexpr (PRO == 1) { x0 = prelu_f(x0, alpha); x1 = prelu_f(x1, alpha); }
        if constexpr (PRO == 2) { x0 = fmaf(c0, prelu_f(x0, alpha), d0); x1 = fmaf(c1, prelu_f(x1, alpha), d1); }
        if constexpr (PRO == 3) { x0 = prelu_f(fmaf(c0, x0, d0), alpha); x1 = prelu_f(fmaf(c1, x1, d1), alpha); }
        if constexpr (X3) {
          split_tf32(x0, bh[ni][0], bl[ni][0]);
          split_tf32(x1, bh[ni][1], bl[ni][1]);
        } else {
          bh[ni][0] = f2tf32(x0); bh[ni][1] = f2tf32(x1);
        }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        int row = wm * 64 + mi * 16 + g;
        float a[4];
        if constexpr (!A_TRANS) {
          a[0] = a_s[row * G_A_LD + kk + tig];
          a[1] = a_s[(row + 8) * G_A_LD + kk + tig];
          a[2] = a_s[row * G_A_LD + kk + tig + 4];
          a[3] = a_s[(row + 8) * G_A_LD + kk + tig + 4];
        } else {
          a[0] = a_s[(kk + tig) * G_AT_LD + row];
          a[1] = a_s[(kk + tig) * G_AT_LD + row + 8];
          a[2] = a_s[(kk + tig + 4) * G_AT_LD + row];
          a[3] = a_s[(kk + tig + 4) * G_AT_LD + row + 8];
        }
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (X3) split_tf32(a[q], ah[q], al[q]);
          else ah[q] = f2tf32(a[q]);
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if constexpr (X3) {
            mma_tf32(acc[mi][ni], al, bh[ni]);
            mma_tf32(acc[mi][ni], ah, bl[ni]);
          }
          mma_tf32(acc[mi][ni], ah, bh[ni]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---------------------------------------------------------------- epilogue
  EpiState<EPI> st;
  const EpiP& e = p.ep;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * 64 + mi * 16 + g + half * 8;
      const bool mok = m < p.M;
      float bias_m = 0.f;
      if (mok) {
        if (e.bias) bias_m = __ldg(e.bias + m);
        if (e.row_bias) bias_m += __ldg(e.row_bias + (int64_t)n * p.M + m);
      }
      float chs = 0.f, chq = 0.f;
      if constexpr (EPI == 10) {
        // TCN backward: acc = (W3^T g)[c=m][t]  ->  dd = dL/d(d), row sums for gLN1 backward
        if (mok) {
          float mu2, r2;
          gln_mean_rstd(e.stats2 + 2 * n, e.count2, e.eps2, mu2, r2);
          const float a2 = __ldg(e.a2), gam2 = __ldg(e.g2 + m);
          const float mh = (float)(e.rowsc[8 * n + 0] / e.count2), mhy = (float)(e.rowsc[8 * n + 1] / e.count2);
          const float gam1 = __ldg(e.g1 + m), bet1 = __ldg(e.be1 + m), bdm = __ldg(e.bd + m);
          const float w0 = __ldg(e.wd + 3 * m), w1 = __ldg(e.wd + 3 * m + 1), w2 = __ldg(e.wd + 3 * m + 2);
          const float* drow = e.d + n * e.bsd + (int64_t)m * e.ldd;
          float* orow = e.Y + n * e.bsy + (int64_t)m * e.ldy;
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int t = t0 + wn * 32 + ni * 8 + 2 * tig;
            if (t < p.T) {
              const bool ok1 = t + 1 < p.T;
              float out2[2];
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                float ddv = 0.f;
                if (q == 0 || ok1) {
                  const int tt = t + q;
                  const float dv = drow[tt];
                  const float h = acc[mi][ni][half * 2 + q] * gam2;
                  const float y2 = prelu_f(dv, a2);
                  const float yh = (y2 - mu2) * r2;
                  const float dy2 = r2 * (h - mh - yh * mhy);
                  ddv = dy2 * (dv > 0.f ? 1.f : a2);
                  const float kap = w1 + (tt >= e.dil ? w0 : 0.f) + (tt < p.T - e.dil ? w2 : 0.f);
                  st.s0 += ddv * gam1 * kap;           // P1
                  st.s1 += ddv * (dv - bdm);           // P2
                  st.s2 += ddv * bet1 * kap;           // P3
                  st.s3 += dv > 0.f ? 0.f : dy2 * dv;  // dalpha2
                }
                out2[q] = ddv;
              }
              epi_store2<EPI>(orow + t, out2[0], out2[1], ok1);
            }
          }
        }
      } else {
        if (mok) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int t = t0 + wn * 32 + ni * 8 + 2 * tig;
            if (t < p.T)
              epi_apply<EPI>(p, st, n, m, t, acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1], t + 1 < p.T, bias_m,
                             chs, chq);
          }
        }
        if constexpr (EPI == 0) {
          if (e.ch_stats) {  // per-channel (row m) sums: reduce over the 4 lanes of the quad
            chs += __shfl_xor_sync(0xffffffffu, chs, 1);
            chq += __shfl_xor_sync(0xffffffffu, chq, 1);
            chs += __shfl_xor_sync(0xffffffffu, chs, 2);
            chq += __shfl_xor_sync(0xffffffffu, chq, 2);
            if (tig == 0 && mok) {
              atomicAdd(e.ch_stats + 2 * m, (double)chs);
              atomicAdd(e.ch_stats + 2 * m + 1, (double)chq);
            }
          }
        }
      }
    }
  }
  if constexpr (EPI == 0) {
    if (e.out_stats) {
      float v[2] = {st.s0, st.s1};
      block_sum<2>(v, red);
      if (tid == 0) {
        atomicAdd(e.out_stats + 2 * n, (double)v[0]);
        atomicAdd(e.out_stats + 2 * n + 1, (double)v[1]);
      }
    }
  }
  if constexpr (EPI == 10) {
    float v[4] = {st.s0, st.s1, st.s2, st.s3};
    block_sum<4>(v, red);
    if (tid == 0) {
      atomicAdd(e.rowacc + 8 * n + 2, (double)v[0]);
      atomicAdd(e.rowacc + 8 * n + 3, (double)v[1]);
      atomicAdd(e.rowacc + 8 * n + 4, (double)v[2]);
      atomicAdd(e.rowacc + 8 * n + 5, (double)v[3]);
    }
  }
}

inline size_t gemm_wx_smem(int pro, int Kd) {
  return (size_t)(G_STAGES * (G_A_TILE + G_B_TILE) + (pro >= 2 ? 2 * Kd : 0)) * sizeof(float);
}

template <bool A_TRANS, int PRO, int EPI>
inline int launch_gemm_wx_t(const GemmWxP& p, cudaStream_t st) {
  dim3 grid(cdiv(p.T, G_BN), cdiv(p.M, G_BM), p.n);
  size_t smem = gemm_wx_smem(PRO, p.Kd);
  if (eff_gemm_mode(p.mode_sel) == 0) {
    auto k = gemm_wx_kernel<A_TRANS, PRO, EPI, true>;
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, G_THREADS, smem, st>>>(p);
  } else {
    auto k = gemm_wx_kernel<A_TRANS, PRO, EPI, false>;
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, G_THREADS, smem, st>>>(p);
  }
  WB_LAUNCH_CHECK("gemm_wx");
  return 0;
}

// Runtime dispatch over (w_trans, pro, epi). Only the combinations the path uses are instantiated.
int launch_gemm_wx(const GemmWxP& p, bool a_trans, int pro, int epi, cudaStream_t st);

// ------------------------------------------------------------------------------------ gemm_dw
struct GemmDwP {
  int n, M, N, T;
  const float* A; int64_t lda; int64_t bsa;
  const float* B; int64_t ldb; int64_t bsb;
  float* C; int64_t ldc; int per_row;
  XformP xb;   // prologue on B: v = sc[c]*prelu(b)+sh[c] with sc = ch_scale*r_n, sh = ch_shift - ch_scale*mu_n*r_n
  int t_chunk; // time steps per CTA (multiple of G_BK)
  float* a_rowsum;  // optional, tcgen05 path only: += [n][M] sum_t A[n][m][t] (a by-product of the operand transform)
  int mode_sel, backend_sel;   // as in GemmWxP
};
int launch_gemm_dw(const GemmDwP& p, int pro_b, cudaStream_t st);
bool gemm_dw_uses_tc(const GemmDwP& p, int pro_b);   // true when launch_gemm_dw will take the tcgen05 path
bool gemm_dw_tc_eligible(const GemmDwP& p, int pro_b);
int launch_gemm_dw_tc(const GemmDwP& p, int pro_b, cudaStream_t st);

}  // namespace wb
