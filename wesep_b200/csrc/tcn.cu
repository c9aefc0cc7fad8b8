// Fused TCN block (Conv1DBlock / Conv1DBlock4Fuse, wesep/modules/tasnet/convs.py:43-160):
// HBM-bound depthwise-dilated stencil kernels + gLN bookkeeping, and the block-level C entry
// points that chain them with the tensor-core 1x1-conv GEMMs (gemm_mma.cuh).
// Math: SURVEY.md Appendix E.1.
#include "gemm_mma.cuh"

namespace wb {

constexpr int ST_CH = 4;        // channel rows per CTA
constexpr int ST_TT = 2048;     // time steps per CTA (forward)
constexpr int ST_TTB = 1024;    // time steps per CTA (backward)
constexpr int ST_THREADS = 256; // 64 threads per channel row
constexpr float GLN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------ K3 forward
// d[c][t] = sum_j wd[c][j] * z1[c][t + (j-1) dil] + bd[c],  z1 = g1*(prelu(u,a1)-mu1)*r1 + be1 inside
// [0,T), zero outside (conv zero padding, convs.py:64-74).  Accumulates gLN2 stats of prelu(d,a2).
struct StFwdP {
  int n, H, T, dil; int64_t ld;
  const float* u; float* d;
  const float* a1; const float* g1; const float* be1;
  const float* wd; const float* bd; const float* a2;
  const double* stats1; double* stats2; double count;
  int tt;   // time steps per CTA of the vectorised kernels (multiple of 256; set by launch_dw_fwd)
};

__global__ void __launch_bounds__(ST_THREADS) tcn_dw_fwd_generic_kernel(const StFwdP p) {
  extern __shared__ float z[];  // [ST_CH][ST_TT + 2*dil]
  __shared__ float red[2 * 32];
  const int tid = threadIdx.x, sub = tid & 63, chl = tid >> 6;
  const int t0 = blockIdx.x * ST_TT, c = blockIdx.y * ST_CH + chl, n = blockIdx.z;
  const int dil = p.dil, W = ST_TT + 2 * dil;
  float s = 0.f, q = 0.f;
  if (c < p.H) {
    float mu, r;
    gln_mean_rstd(p.stats1 + 2 * n, p.count, GLN_EPS, mu, r);
    const float a1 = __ldg(p.a1), gm = __ldg(p.g1 + c), bt = __ldg(p.be1 + c);
    const float sc = gm * r, sh = bt - gm * mu * r;
    const float* urow = p.u + ((int64_t)n * p.H + c) * p.ld;
    float* zr = z + chl * W;
    for (int i = sub; i < W; i += 64) {
      int t = t0 - dil + i;
      float v = 0.f;
      if (t >= 0 && t < p.T) v = fmaf(sc, prelu_f(__ldg(urow + t), a1), sh);
      zr[i] = v;
    }
  }
  __syncthreads();
  if (c < p.H) {
    const float w0 = __ldg(p.wd + 3 * c), w1 = __ldg(p.wd + 3 * c + 1), w2 = __ldg(p.wd + 3 * c + 2);
    const float bd = __ldg(p.bd + c), a2 = __ldg(p.a2);
    const float* zr = z + chl * W + dil;
    float* drow = p.d + ((int64_t)n * p.H + c) * p.ld;
#pragma unroll 4
    for (int i = sub; i < ST_TT; i += 64) {
      int t = t0 + i;
      if (t < p.T) {
        float dv = fmaf(w0, zr[i - dil], fmaf(w1, zr[i], fmaf(w2, zr[i + dil], bd)));
        drow[t] = dv;
        float y = prelu_f(dv, a2);
        s += y;
        q += y * y;
      }
    }
  }
  float v[2] = {s, q};
  block_sum<2>(v, red);
  if (tid == 0) {
    atomicAdd(p.stats2 + 2 * n, (double)v[0]);
    atomicAdd(p.stats2 + 2 * n + 1, (double)v[1]);
  }
}

// ------------------------------------------------------------------------------------ B3 backward
// inputs dd = dL/d(d), u;  outputs du = dL/du and parameter-gradient partial sums.
struct StBwdP {
  int n, H, T, dil; int64_t ld;
  const float* u; const float* dd; float* du;
  const float* a1; const float* g1; const float* be1; const float* wd;
  const double* stats1; double count;
  const double* rowacc;          // [n][8]: [2]=P1 [3]=P2 [4]=P3
  float* dg1; float* dbe1; float* dwd; float* dbd; float* da1;  // += [H],[H],[H][3],[H],[1]
  float* sdu;                    // += [n][H] sum_t du
  int tt;                        // time steps per CTA of the vectorised kernels (multiple of 256; set by launch_dw_bwd)
};

__global__ void __launch_bounds__(ST_THREADS) tcn_dw_bwd_generic_kernel(const StBwdP p) {
  extern __shared__ float sm[];  // dds[CH][W], z1s[CH][W], us[CH][TT]
  __shared__ float red[32];
  __shared__ float chred[ST_CH][2][8];
  const int tid = threadIdx.x, sub = tid & 63, chl = tid >> 6, lane = tid & 31;
  const int t0 = blockIdx.x * ST_TTB, c = blockIdx.y * ST_CH + chl, n = blockIdx.z;
  const int dil = p.dil, W = ST_TTB + 2 * dil;
  float* dds = sm + chl * W;
  float* z1s = sm + ST_CH * W + chl * W;
  float* us = sm + 2 * ST_CH * W + chl * ST_TTB;
  float mu = 0.f, r = 1.f, a1 = 1.f, gm = 0.f;
  if (c < p.H) {
    gln_mean_rstd(p.stats1 + 2 * n, p.count, GLN_EPS, mu, r);
    a1 = __ldg(p.a1);
    gm = __ldg(p.g1 + c);
    const float bt = __ldg(p.be1 + c);
    const float sc = gm * r, sh = bt - gm * mu * r;
    const float* urow = p.u + ((int64_t)n * p.H + c) * p.ld;
    const float* drow = p.dd + ((int64_t)n * p.H + c) * p.ld;
    for (int i = sub; i < W; i += 64) {
      int t = t0 - dil + i;
      float dv = 0.f, zv = 0.f;
      if (t >= 0 && t < p.T) {
        dv = __ldg(drow + t);
        float uv = __ldg(urow + t);
        zv = fmaf(sc, prelu_f(uv, a1), sh);
        int j = i - dil;
        if (j >= 0 && j < ST_TTB) us[j] = uv;
      }
      dds[i] = dv;
      z1s[i] = zv;
    }
  }
  __syncthreads();
  float acc[8];  // 0 sdz(dbe1) 1 sdzy(dg1) 2..4 dw0..2 5 sdd(dbd) 6 sdu 7 unused
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float dal = 0.f;
  if (c < p.H) {
    const double iMc = __drcp_rn(p.count);
    const float m1 = (float)(p.rowacc[8 * n + 2] * iMc);
    const float m2 = (float)((p.rowacc[8 * n + 3] - p.rowacc[8 * n + 4]) * iMc);
    const float w0 = __ldg(p.wd + 3 * c), w1 = __ldg(p.wd + 3 * c + 1), w2 = __ldg(p.wd + 3 * c + 2);
    float* durow = p.du + ((int64_t)n * p.H + c) * p.ld;
    const float* dc = dds + dil;
    const float* zc = z1s + dil;
#pragma unroll 2
    for (int i = sub; i < ST_TTB; i += 64) {
      int t = t0 + i;
      if (t < p.T) {
        const float ddv = dc[i];
        // dz1[t] = sum_j w[j] * dd[t - (j-1) dil]
        const float dz = fmaf(w0, dc[i + dil], fmaf(w1, ddv, w2 * dc[i - dil]));
        const float uv = us[i];
        const float yh = (prelu_f(uv, a1) - mu) * r;
        const float dy = r * (gm * dz - m1 - yh * m2);
        const float duv = dy * (uv > 0.f ? 1.f : a1);
        durow[t] = duv;
        acc[0] += dz;
        acc[1] += dz * yh;
        acc[2] += ddv * zc[i - dil];
        acc[3] += ddv * zc[i];
        acc[4] += ddv * zc[i + dil];
        acc[5] += ddv;
        acc[6] += duv;
        dal += uv > 0.f ? 0.f : dy * uv;
      }
    }
  }
  // per-channel reduction: the 64 threads (2 warps) of a channel row
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 7; ++i) chred[chl][(tid >> 5) & 1][i] = acc[i];
  }
  float v[1] = {dal};
  block_sum<1>(v, red);  // contains __syncthreads (also publishes chred)
  if (tid == 0 && v[0] != 0.f) atomicAdd(p.da1, v[0]);
  if (sub == 0 && c < p.H) {
    float r7[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) r7[i] = chred[chl][0][i] + chred[chl][1][i];
    atomicAdd(p.dbe1 + c, r7[0]);
    atomicAdd(p.dg1 + c, r7[1]);
    atomicAdd(p.dwd + 3 * c + 0, r7[2]);
    atomicAdd(p.dwd + 3 * c + 1, r7[3]);
    atomicAdd(p.dwd + 3 * c + 2, r7[4]);
    atomicAdd(p.dbd + c, r7[5]);
    atomicAdd(p.sdu + (int64_t)n * p.H + c, r7[6]);
  }
}


// ------------------------------------------------------------------------------------ vectorised stencils
// Same math as the *_generic kernels above, 4 consecutive time steps per thread, taps read straight from global
// memory as 128-bit loads: the centre vector comes from DRAM once, the two tap vectors hit L1 / L2 (another
// thread's centre), so there is no shared-memory fill phase and no barrier.  DM = 0: dilation is a multiple of 4
// (taps are 16-byte aligned); DM = 1,2,3: dilation == DM (taps assembled from the previous / centre / next vectors).
// Pad columns [T, ld) of the inputs hold arbitrary bits (possibly NaN): every element >= T is replaced, never scaled.
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
// L / C / R = x[t + k - dil], x[t + k], x[t + k + dil] (k < 4) with `fill` outside [0, T)
template <int DM>
__device__ __forceinline__ void taps4g(const float* row, int t, int T, int dil, float fill, float (&L)[4], float (&C)[4],
                                       float (&R)[4]) {
  ld4(row + t, C);
  if constexpr (DM == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { L[k] = fill; R[k] = fill; }
    if (t - dil >= 0) ld4(row + t - dil, L);
    if (t + dil < T) ld4(row + t + dil, R);
    if (t + dil + 3 >= T) {   // right edge: at most the last vectors of a row
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t + k >= T) C[k] = fill;
        if (t + dil + k >= T) R[k] = fill;
      }
    }
  } else {
    float a[12];
    float P[4] = {fill, fill, fill, fill}, N[4] = {fill, fill, fill, fill};
    if (t >= 4) ld4(row + t - 4, P);
    if (t + 4 < T) ld4(row + t + 4, N);
    if (t + 7 >= T) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t + k >= T) C[k] = fill;
        if (t + 4 + k >= T) N[k] = fill;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = P[k]; a[4 + k] = C[k]; a[8 + k] = N[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      L[k] = a[4 + k - DM];
      R[k] = a[4 + k + DM];
    }
  }
}

// PReLU with slope a <= 1 is max(x, a*x) (2 instructions instead of compare + multiply + select); the kernels pick
// the FAST instantiation when the (kernel-uniform) learned slopes allow it.
template <bool FAST>
__device__ __forceinline__ float prelu_t(float x, float a) {
  if constexpr (FAST) return fmaxf(x, a * x);
  else return prelu_f(x, a);
}

template <int DM, bool FAST>
__device__ __forceinline__ void dw_fwd_body(const StFwdP& p, int sub, int t0, int c, int n, float& s, float& q) {
  const int dil = p.dil;
  float mu, r;
  gln_mean_rstd(p.stats1 + 2 * n, p.count, GLN_EPS, mu, r);
  const float a1 = __ldg(p.a1), gm = __ldg(p.g1 + c), bt = __ldg(p.be1 + c);
  const float sc = gm * r, sh = bt - gm * mu * r;
  const float w0 = __ldg(p.wd + 3 * c), w1 = __ldg(p.wd + 3 * c + 1), w2 = __ldg(p.wd + 3 * c + 2);
  const float bd = __ldg(p.bd + c), a2 = __ldg(p.a2);
  // interior: d = sum_j w_j (sc*y_j + sh) + bd = sum_j (w_j sc) y_j + (bd + sh sum_j w_j),  y = prelu(u)
  const float v0 = w0 * sc, v1 = w1 * sc, v2 = w2 * sc, bdp = fmaf(sh, w0 + w1 + w2, bd);
  const float* urow = p.u + ((int64_t)n * p.H + c) * p.ld;
  float* drow = p.d + ((int64_t)n * p.H + c) * p.ld;
  // two vectors per pass, all six tap loads issued before the first use: the kernel is DRAM-latency bound, so the bytes
  // in flight per thread (not the instruction count) set its speed
  auto emit = [&](int t, const float (&L)[4], const float (&C)[4], const float (&R)[4]) {
    float o[4];
    const bool interior = (t - dil >= 0) && (t + dil + 3 < p.T);
    if (interior) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[k] = fmaf(v0, prelu_t<FAST>(L[k], a1), fmaf(v1, prelu_t<FAST>(C[k], a1), fmaf(v2, prelu_t<FAST>(R[k], a1), bdp)));
        const float y = prelu_t<FAST>(o[k], a2);
        s += y;
        q = fmaf(y, y, q);
      }
    } else {   // z1 = sc*prelu(u) + sh inside [0,T) and 0 outside
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = t + k;
        const float zl = (e - dil >= 0) ? fmaf(sc, prelu_f(L[k], a1), sh) : 0.f;
        const float zc = (e < p.T) ? fmaf(sc, prelu_f(C[k], a1), sh) : 0.f;
        const float zr = (e + dil < p.T) ? fmaf(sc, prelu_f(R[k], a1), sh) : 0.f;
        o[k] = fmaf(w0, zl, fmaf(w1, zc, fmaf(w2, zr, bd)));
        const float y = (e < p.T) ? prelu_f(o[k], a2) : 0.f;
        s += y;
        q = fmaf(y, y, q);
      }
    }
    *reinterpret_cast<float4*>(drow + t) = make_float4(o[0], o[1], o[2], o[3]);
  };
#pragma unroll 1
  for (int i = 4 * sub; i < p.tt; i += 512) {
    const int ta = t0 + i, tb = ta + 256;
    if (ta >= p.T) break;
    const bool has_b = (i + 256 < p.tt) && (tb < p.T);
    float La[4], Ca[4], Ra[4], Lb[4], Cb[4], Rb[4];
    taps4g<DM>(urow, ta, p.T, dil, 0.f, La, Ca, Ra);
    if (has_b) taps4g<DM>(urow, tb, p.T, dil, 0.f, Lb, Cb, Rb);
    emit(ta, La, Ca, Ra);
    if (has_b) emit(tb, Lb, Cb, Rb);
  }
}

template <int DM>
__global__ void __launch_bounds__(ST_THREADS) tcn_dw_fwd_kernel(const StFwdP p) {
  __shared__ float red[2 * 32];
  const int tid = threadIdx.x, sub = tid & 63, chl = tid >> 6;
  const int t0 = blockIdx.x * p.tt, c = blockIdx.y * ST_CH + chl, n = blockIdx.z;
  float s = 0.f, q = 0.f;
  if (c < p.H) {
    if (__ldg(p.a1) <= 1.f && __ldg(p.a2) <= 1.f) dw_fwd_body<DM, true>(p, sub, t0, c, n, s, q);
    else dw_fwd_body<DM, false>(p, sub, t0, c, n, s, q);
  }
  float v[2] = {s, q};
  block_sum<2>(v, red);
  if (tid == 0) {
    atomicAdd(p.stats2 + 2 * n, (double)v[0]);
    atomicAdd(p.stats2 + 2 * n + 1, (double)v[1]);
  }
}

// Backward: the weight gradients are re-indexed so that only dd needs taps:
//   dwd[c][0] = sum_t dd[t] z1[t - dil] = sum_t z1[t] dd[t + dil],  dwd[c][2] = sum_t z1[t] dd[t - dil]
// (dd == 0 outside [0,T)), so u is read once, at the centre.
template <int DM, bool FAST>
__device__ __forceinline__ void dw_bwd_body(const StBwdP& p, int sub, int t0, int c, int n, float (&acc)[8], float& dal) {
  const int dil = p.dil;
  float mu, r;
  gln_mean_rstd(p.stats1 + 2 * n, p.count, GLN_EPS, mu, r);
  const float a1 = __ldg(p.a1), gm = __ldg(p.g1 + c), bt = __ldg(p.be1 + c);
  const float sc = gm * r, sh = bt - gm * mu * r;
  const double iMc = __drcp_rn(p.count);
  const float m1 = (float)(p.rowacc[8 * n + 2] * iMc);
  const float m2 = (float)((p.rowacc[8 * n + 3] - p.rowacc[8 * n + 4]) * iMc);
  const float w0 = __ldg(p.wd + 3 * c), w1 = __ldg(p.wd + 3 * c + 1), w2 = __ldg(p.wd + 3 * c + 2);
  const float* urow = p.u + ((int64_t)n * p.H + c) * p.ld;
  const float* drow = p.dd + ((int64_t)n * p.H + c) * p.ld;
  float* durow = p.du + ((int64_t)n * p.H + c) * p.ld;
  // dy = r*(g1*dz - m1 - yhat1*m2) = cA*dz + cB*y1 + cC with y1 = prelu(u)
  const float cA = r * gm, cB = -r * r * m2, cC = -r * m1 + r * r * m2 * mu;
#pragma unroll 2
  for (int i = 4 * sub; i < p.tt; i += 256) {
    const int t = t0 + i;
    if (t >= p.T) break;
    float dL[4], dC[4], dR[4], uu[4], o[4];
    taps4g<DM>(drow, t, p.T, dil, 0.f, dL, dC, dR);   // dd outside [0,T) is 0
    ld4(urow + t, uu);
    if (t + 3 < p.T) {   // full vector (all but the last vector of a row)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float ddv = dC[k], uv = uu[k];
        const float dz = fmaf(w0, dR[k], fmaf(w1, ddv, w2 * dL[k]));  // dz1[t] = sum_j w[j] dd[t - (j-1) dil]
        const float y1 = prelu_t<FAST>(uv, a1);
        const float z1 = fmaf(sc, y1, sh);
        const float dy = fmaf(cA, dz, fmaf(cB, y1, cC));
        const float duv = uv > 0.f ? dy : a1 * dy;
        o[k] = duv;
        acc[0] += dz;                        // sum dz            -> dbeta1
        acc[1] = fmaf(dz, y1, acc[1]);       // sum dz*y1         -> dgamma1 = r*(sum dz*y1 - mu*sum dz)
        acc[2] = fmaf(z1, dR[k], acc[2]);
        acc[3] = fmaf(z1, ddv, acc[3]);
        acc[4] = fmaf(z1, dL[k], acc[4]);
        acc[5] += ddv;
        acc[6] += duv;
        dal = fmaf(dy, fminf(uv, 0.f), dal);   // sum over u <= 0 of dy*u -> dalpha1
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = t + k < p.T;
        const float ddv = dC[k];
        const float dz = fmaf(w0, dR[k], fmaf(w1, ddv, w2 * dL[k]));
        const float uv = ok ? uu[k] : 1.f;
        const bool pos = uv > 0.f;
        const float y1 = pos ? uv : a1 * uv;
        const float z1 = ok ? fmaf(sc, y1, sh) : 0.f;
        const float dy = fmaf(cA, dz, fmaf(cB, y1, cC));
        float duv = pos ? dy : a1 * dy;
        float dzy = dz * y1, dyu = pos ? 0.f : dy * uv, dzm = dz;
        if (!ok) { duv = 0.f; dzy = 0.f; dyu = 0.f; dzm = 0.f; }
        o[k] = duv;
        acc[0] += dzm;
        acc[1] += dzy;
        acc[2] = fmaf(z1, dR[k], acc[2]);
        acc[3] = fmaf(z1, ddv, acc[3]);
        acc[4] = fmaf(z1, dL[k], acc[4]);
        acc[5] += ddv;
        acc[6] += duv;
        dal += dyu;
      }
    }
    *reinterpret_cast<float4*>(durow + t) = make_float4(o[0], o[1], o[2], o[3]);
  }
  acc[1] = r * (acc[1] - mu * acc[0]);
}

template <int DM>
__global__ void __launch_bounds__(ST_THREADS) tcn_dw_bwd_kernel(const StBwdP p) {
  __shared__ float red[32];
  __shared__ float chred[ST_CH][2][8];
  const int tid = threadIdx.x, sub = tid & 63, chl = tid >> 6, lane = tid & 31;
  const int t0 = blockIdx.x * p.tt, c = blockIdx.y * ST_CH + chl, n = blockIdx.z;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float dal = 0.f;
  if (c < p.H) {
    if (__ldg(p.a1) <= 1.f) dw_bwd_body<DM, true>(p, sub, t0, c, n, acc, dal);
    else dw_bwd_body<DM, false>(p, sub, t0, c, n, acc, dal);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 7; ++i) chred[chl][(tid >> 5) & 1][i] = acc[i];
  }
  float v[1] = {dal};
  block_sum<1>(v, red);
  if (tid == 0 && v[0] != 0.f) atomicAdd(p.da1, v[0]);
  if (sub == 0 && c < p.H) {
    float r7[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) r7[i] = chred[chl][0][i] + chred[chl][1][i];
    atomicAdd(p.dbe1 + c, r7[0]);
    atomicAdd(p.dg1 + c, r7[1]);
    atomicAdd(p.dwd + 3 * c + 0, r7[2]);
    atomicAdd(p.dwd + 3 * c + 1, r7[3]);
    atomicAdd(p.dwd + 3 * c + 2, r7[4]);
    atomicAdd(p.dbd + c, r7[5]);
    atomicAdd(p.sdu + (int64_t)n * p.H + c, r7[6]);
  }
}

// Time steps per CTA of the vectorised stencils: a CTA's fixed cost (parameter / statistics loads up front, seven
// warp reductions + atomics at the end) was ~40 % of its lifetime with 1024-step tiles, so rows are cut into as few
// equal pieces as keep the piece <= 3584 steps (T = 6399 -> 2 x 3200).
static int stencil_tile(int T) {
  const int pieces = cdiv(T, 3584);
  return cdiv(cdiv(T, pieces), 256) * 256;
}

static int launch_dw_fwd(const StFwdP& pin, cudaStream_t st) {
  StFwdP p = pin;
  p.tt = ST_TT;
  const int dm_ = (p.dil % 4 == 0) ? 0 : (p.dil <= 3 ? p.dil : -1);
  if (dm_ >= 0) p.tt = stencil_tile(p.T);
  dim3 grid(cdiv(p.T, p.tt), cdiv(p.H, ST_CH), p.n);
  const int dm = (p.dil % 4 == 0) ? 0 : (p.dil <= 3 ? p.dil : -1);
  if (dm < 0) {
    size_t smem = (size_t)ST_CH * (ST_TT + 2 * p.dil) * sizeof(float);
    WB_CUDA(cudaFuncSetAttribute(tcn_dw_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tcn_dw_fwd_generic_kernel<<<grid, ST_THREADS, smem, st>>>(p);
  } else {
    auto launch = [&](auto k) -> int {
      k<<<grid, ST_THREADS, 0, st>>>(p);
      return 0;
    };
    int rc = dm == 0 ? launch(tcn_dw_fwd_kernel<0>) : dm == 1 ? launch(tcn_dw_fwd_kernel<1>)
                     : dm == 2 ? launch(tcn_dw_fwd_kernel<2>) : launch(tcn_dw_fwd_kernel<3>);
    if (rc) return rc;
  }
  WB_LAUNCH_CHECK("tcn_dw_fwd");
  return 0;
}

static int launch_dw_bwd(const StBwdP& pin, cudaStream_t st) {
  StBwdP p = pin;
  p.tt = ST_TTB;
  const int dm_ = (p.dil % 4 == 0) ? 0 : (p.dil <= 3 ? p.dil : -1);
  if (dm_ >= 0) p.tt = stencil_tile(p.T);
  dim3 grid(cdiv(p.T, p.tt), cdiv(p.H, ST_CH), p.n);
  const int dm = (p.dil % 4 == 0) ? 0 : (p.dil <= 3 ? p.dil : -1);
  if (dm < 0) {
    size_t smem = (size_t)ST_CH * (2 * (ST_TTB + 2 * p.dil) + ST_TTB) * sizeof(float);
    WB_CUDA(cudaFuncSetAttribute(tcn_dw_bwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tcn_dw_bwd_generic_kernel<<<grid, ST_THREADS, smem, st>>>(p);
  } else {
    auto launch = [&](auto k) -> int {
      k<<<grid, ST_THREADS, 0, st>>>(p);
      return 0;
    };
    int rc = dm == 0 ? launch(tcn_dw_bwd_kernel<0>) : dm == 1 ? launch(tcn_dw_bwd_kernel<1>)
                     : dm == 2 ? launch(tcn_dw_bwd_kernel<2>) : launch(tcn_dw_bwd_kernel<3>);
    if (rc) return rc;
  }
  WB_LAUNCH_CHECK("tcn_dw_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------ small kernels
// out[n][c] = sum_t x[n][c][t]; one warp per row.
__global__ void rowsum_kernel(const float* x, int64_t ld, int rows, int T, float* out) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= rows) return;
  const float* r = x + (int64_t)w * ld;
  float s = 0.f;
  int T4 = T & ~3;
  for (int t = lane * 4; t < T4; t += 128) {
    float4 v = __ldg(reinterpret_cast<const float4*>(r + t));
    s += (v.x + v.y) + (v.z + v.w);
  }
  for (int t = T4 + lane; t < T; t += 32) s += __ldg(r + t);
  s = warp_sum(s);
  if (lane == 0) out[w] = s;
}

// gLN2 backward, step 1 (per row): means of h and h*yhat2 from Gn = sum_t g y2^T and sg = sum_t g.
struct F2P {
  int n, B, H; double count;  // count = H*T
  const float* Gn; const float* sg; const float* W3; int64_t ldw3;
  const float* g2; const float* be2; const double* stats2;
  double* rowsc;  // [n][8]: out [0]=mean(h) [1]=mean(h yhat2) [6]=mu2 [7]=r2
  float* dW3; float* dg2; float* dbe2; float* db3;
};
__global__ void __launch_bounds__(256) tcn_f2a_kernel(const F2P p) {
  // grid (cdiv(H,64), n): thread = (channel c, 1 of 4 o-groups); raw sums -> rowsc[8n+0], [8n+1] (double atomics)
  __shared__ float red[2 * 32];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int c = blockIdx.x * 64 + (tid & 63), og = tid >> 6;
  float mu, r;
  gln_mean_rstd(p.stats2 + 2 * n, p.count, GLN_EPS, mu, r);
  const float* G = p.Gn + (int64_t)n * p.B * p.H;
  const float* sg = p.sg + (int64_t)n * p.B;
  float s1 = 0.f, s2 = 0.f;
  if (c < p.H) {
    float a = 0.f, b = 0.f;
#pragma unroll 4
    for (int o = og; o < p.B; o += 4) {
      const float w = __ldg(p.W3 + (int64_t)o * p.ldw3 + c);
      const float sgo = __ldg(sg + o);
      a = fmaf(w, sgo, a);
      b = fmaf(w, r * (__ldg(G + (int64_t)o * p.H + c) - mu * sgo), b);
    }
    const float gm = __ldg(p.g2 + c);
    s1 = gm * a;
    s2 = gm * b;
  }
  float v[2] = {s1, s2};
  block_sum<2>(v, red);
  if (tid == 0) {
    atomicAdd(p.rowsc + 8 * n + 0, (double)v[0]);   // sum_{c,t} h        (consumers divide by count)
    atomicAdd(p.rowsc + 8 * n + 1, (double)v[1]);   // sum_{c,t} h*yhat2
    if (blockIdx.x == 0) {
      p.rowsc[8 * n + 6] = (double)mu;
      p.rowsc[8 * n + 7] = (double)r;
    }
  }
}
// step 2 (per weight element): dW3, dgamma2, dbeta2, db3. grid (cdiv(H,128), B), block 128: thread <-> (o, c);
// the n loads per thread are independent (unrolled) so the kernel is bandwidth- not latency-bound.
__global__ void __launch_bounds__(128) tcn_f2b_kernel(const F2P p) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int o = blockIdx.y;
  if (c >= p.H) return;
  float a1 = 0.f, a2 = 0.f;
  const float* gp = p.Gn + (int64_t)o * p.H + c;
  const int64_t gs = (int64_t)p.B * p.H;
  int n = 0;
  for (; n + 8 <= p.n; n += 8) {
    float gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = __ldg(gp + (n + j) * gs);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mu = (float)p.rowsc[8 * (n + j) + 6], r = (float)p.rowsc[8 * (n + j) + 7];   // written by tcn_f2a_kernel
      const float sgo = __ldg(p.sg + (int64_t)(n + j) * p.B + o);
      a1 = fmaf(r, gv[j] - mu * sgo, a1);
      a2 += sgo;
    }
  }
  for (; n < p.n; ++n) {
    const float mu = (float)p.rowsc[8 * n + 6], r = (float)p.rowsc[8 * n + 7];
    const float sgo = __ldg(p.sg + (int64_t)n * p.B + o);
    a1 = fmaf(r, __ldg(gp + n * gs) - mu * sgo, a1);
    a2 += sgo;
  }
  const float gm = __ldg(p.g2 + c), bt = __ldg(p.be2 + c);
  const float w = __ldg(p.W3 + (int64_t)o * p.ldw3 + c);
  atomicAdd(p.dW3 + (int64_t)o * p.ldw3 + c, gm * a1 + bt * a2);
  atomicAdd(p.dg2 + c, w * a1);
  atomicAdd(p.dbe2 + c, w * a2);
  if (c == 0) atomicAdd(p.db3 + o, a2);
}

// speaker fold of the fuse block: row_bias[n][h] = sum_e W1[h][B+e] aux[n][e]; one warp per (n,h)
__global__ void fuse_rowbias_kernel(const float* W1, int64_t ldw1, int B, int E, int H, int n, const float* aux,
                                    float* row_bias) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n * H) return;
  int r = w / H, h = w % H;
  const float* wr = W1 + (int64_t)h * ldw1 + B;
  const float* a = aux + (int64_t)r * E;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) s = fmaf(__ldg(wr + e), __ldg(a + e), s);
  s = warp_sum(s);
  if (lane == 0) row_bias[w] = s;
}

// tail of the backward: db1, da2 and (fuse block) daux, dW1[:, B:]
struct TailP {
  int n, B, H, E; int64_t ldw1;
  const float* sdu; const double* rowacc; const float* W1; const float* aux;
  float* db1; float* da2; float* daux; float* dW1;
};
__global__ void tcn_bwd_tail_kernel(const TailP p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.H) {
    float s = 0.f;
#pragma unroll 8
    for (int n = 0; n < p.n; ++n) s += p.sdu[(int64_t)n * p.H + i];
    atomicAdd(p.db1 + i, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {   // da2 = sum_n rowacc[n][5]: one warp, independent loads
    double s = 0.0;
    for (int n = threadIdx.x; n < p.n; n += 32) s += p.rowacc[8 * n + 5];
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(p.da2, (float)s);
  }
  if (p.E > 0) {
    if (i < p.n * p.E) {  // daux[n][e] = sum_h W1[h][B+e] sdu[n][h]
      int n = i / p.E, e = i % p.E;
      float s = 0.f;
      for (int h = 0; h < p.H; ++h) s = fmaf(__ldg(p.W1 + (int64_t)h * p.ldw1 + p.B + e), p.sdu[(int64_t)n * p.H + h], s);
      p.daux[i] = s;
    }
    if (i < p.H * p.E) {  // dW1[h][B+e] += sum_n sdu[n][h] aux[n][e]
      int h = i / p.E, e = i % p.E;
      float s = 0.f;
      for (int n = 0; n < p.n; ++n) s = fmaf(p.sdu[(int64_t)n * p.H + h], __ldg(p.aux + (int64_t)n * p.E + e), s);
      atomicAdd(p.dW1 + (int64_t)h * p.ldw1 + p.B + e, s);
    }
  }
}

static int check_tcn(const WesepTcnFwdArgs& a) {
  if (a.n <= 0 || a.B <= 0 || a.H <= 0 || a.T <= 0 || a.dil <= 0) return fail(-1, "tcn: empty shape");
  if ((a.ld & 3) || a.ld < a.T) return fail(-1, "tcn: ld must be a multiple of 4 and >= T");
  if ((a.B & 3) || (a.H & 3)) return fail(-1, "tcn: B and H must be multiples of 4");
  if (a.dil > 4096) return fail(-2, "tcn: dilation too large");
  if ((a.E > 0) != (a.aux != nullptr)) return fail(-1, "tcn: aux and E must be given together");
  if (a.ldw1 < a.B + a.E || (a.ldw1 & 3) || (a.ldw3 & 3)) return fail(-1, "tcn: weight strides");
  return 0;
}

}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_tcn_block_fwd(const WesepTcnFwdArgs* ap, void* stream) {
  const WesepTcnFwdArgs& a = *ap;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_tcn(a)) return rc;
  const double count = (double)a.H * (double)a.T;
  if (a.stats2 == a.stats1 + 2 * a.n) {   // adjacent (the Python host allocates them as one tensor): one memset
    WB_CUDA(cudaMemsetAsync(a.stats1, 0, sizeof(double) * 4 * a.n, st));
  } else {
    WB_CUDA(cudaMemsetAsync(a.stats1, 0, sizeof(double) * 2 * a.n, st));
    WB_CUDA(cudaMemsetAsync(a.stats2, 0, sizeof(double) * 2 * a.n, st));
  }
  if (a.aux) {
    int warps = a.n * a.H;
    fuse_rowbias_kernel<<<cdiv((int64_t)warps * 32, 256), 256, 0, st>>>(a.W1, a.ldw1, a.B, a.E, a.H, a.n, a.aux, a.row_bias);
    WB_LAUNCH_CHECK("fuse_rowbias");
  }
  {  // K2: u = W1 x + b1 (+ row bias); gLN1 statistics of prelu(u)
    GemmWxP p{};
    p.n = a.n; p.M = a.H; p.Kd = a.B; p.T = a.T;
    p.W = a.W1; p.ldw = a.ldw1; p.X = a.x; p.ldx = a.ld; p.bsx = (int64_t)a.B * a.ld;
    p.ep.Y = a.u; p.ep.ldy = a.ld; p.ep.bsy = (int64_t)a.H * a.ld;
    p.ep.bias = a.b1; p.ep.row_bias = a.aux ? a.row_bias : nullptr;
    p.ep.out_stats = a.stats1; p.ep.out_alpha = a.a1;
    p.ws = a.ws; p.ws_bytes = a.ws_bytes;
    if (int rc = launch_gemm_wx(p, false, 0, 0, st)) return rc;
  }
  {  // K3: depthwise dilated conv on gLN1(prelu(u)); gLN2 statistics
    StFwdP p{a.n, a.H, a.T, a.dil, a.ld, a.u, a.d, a.a1, a.g1, a.be1, a.wd, a.bd, a.a2, a.stats1, a.stats2, count};
    if (int rc = launch_dw_fwd(p, st)) return rc;
  }
  {  // K4: out = x + W3 gLN2(prelu(d)) + b3
    GemmWxP p{};
    p.n = a.n; p.M = a.B; p.Kd = a.H; p.T = a.T;
    p.W = a.W3; p.ldw = a.ldw3; p.X = a.d; p.ldx = a.ld; p.bsx = (int64_t)a.H * a.ld;
    p.xf = XformP{a.a2, a.g2, a.be2, a.stats2, count, GLN_EPS};
    p.ep.Y = a.out; p.ep.ldy = a.ld; p.ep.bsy = (int64_t)a.B * a.ld;
    p.ep.bias = a.b3;
    p.ep.R = a.x; p.ep.ldr = a.ld; p.ep.bsr = (int64_t)a.B * a.ld;
    p.ws = a.ws; p.ws_bytes = a.ws_bytes;
    if (int rc = launch_gemm_wx(p, false, 2, 2, st)) return rc;
  }
  return 0;
}

extern "C" int wesep_b200_tcn_block_bwd(const WesepTcnBwdArgs* bp, void* stream) {
  const WesepTcnBwdArgs& b = *bp;
  const WesepTcnFwdArgs& a = b.f;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_tcn(a)) return rc;
  const double count = (double)a.H * (double)a.T;
  // scratch the block zeroes itself: [rowsc | sdu | sg | Gn]; when the caller laid them out back to back, one memset
  const bool packed = b.sdu == reinterpret_cast<float*>(b.rowsc + 8 * a.n) && b.sg == b.sdu + (size_t)a.n * a.H &&
                      b.Gn == b.sg + (size_t)a.n * a.B;
  if (packed) {
    WB_CUDA(cudaMemsetAsync(b.rowsc, 0, sizeof(double) * 8 * a.n + sizeof(float) * ((size_t)a.n * a.H + (size_t)a.n * a.B +
                                                                                     (size_t)a.n * a.B * a.H), st));
  } else {
    WB_CUDA(cudaMemsetAsync(b.Gn, 0, sizeof(float) * (size_t)a.n * a.B * a.H, st));
    WB_CUDA(cudaMemsetAsync(b.sdu, 0, sizeof(float) * (size_t)a.n * a.H, st));
    WB_CUDA(cudaMemsetAsync(b.rowsc, 0, sizeof(double) * 8 * a.n, st));
  }
  {  // Gn[n][o][c] = sum_t g[o][t] * prelu(d[c][t], a2)  and  sg[n][o] = sum_t g[o][t]
    GemmDwP p{};
    p.n = a.n; p.M = a.B; p.N = a.H; p.T = a.T;
    p.A = b.gout; p.lda = a.ld; p.bsa = (int64_t)a.B * a.ld;
    p.B = a.d; p.ldb = a.ld; p.bsb = (int64_t)a.H * a.ld;
    p.C = b.Gn; p.ldc = a.H; p.per_row = 1;
    p.xb = XformP{a.a2, nullptr, nullptr, nullptr, 1.0, 0.f};
    if (gemm_dw_uses_tc(p, 1)) {   // the tcgen05 kernel sums the rows of g while it splits the operand tiles
      if (!packed) WB_CUDA(cudaMemsetAsync(b.sg, 0, sizeof(float) * (size_t)a.n * a.B, st));
      p.a_rowsum = b.sg;
    } else {
      int rows = a.n * a.B;
      rowsum_kernel<<<cdiv((int64_t)rows * 32, 256), 256, 0, st>>>(b.gout, a.ld, rows, a.T, b.sg);
      WB_LAUNCH_CHECK("rowsum");
    }
    if (int rc = launch_gemm_dw(p, 1, st)) return rc;
  }
  {
    F2P p{a.n, a.B, a.H, count, b.Gn, b.sg, a.W3, a.ldw3, a.g2, a.be2, a.stats2, b.rowsc, b.dW3, b.dg2, b.dbe2, b.db3};
    tcn_f2a_kernel<<<dim3(cdiv(a.H, 64), a.n), 256, 0, st>>>(p);
    WB_LAUNCH_CHECK("tcn_f2a");
    tcn_f2b_kernel<<<dim3(cdiv(a.H, 128), a.B), 128, 0, st>>>(p);
    WB_LAUNCH_CHECK("tcn_f2b");
  }
  {  // B2: dd = dL/d(d) from h = g2 * (W3^T g), gLN2 + PReLU_2 backward fused in the epilogue
    GemmWxP p{};
    p.n = a.n; p.M = a.H; p.Kd = a.B; p.T = a.T;
    p.W = a.W3; p.ldw = a.ldw3; p.X = b.gout; p.ldx = a.ld; p.bsx = (int64_t)a.B * a.ld;
    EpiP& e = p.ep;
    e.Y = b.dd; e.ldy = a.ld; e.bsy = (int64_t)a.H * a.ld;
    e.d = a.d; e.ldd = a.ld; e.bsd = (int64_t)a.H * a.ld;
    e.a2 = a.a2; e.g2 = a.g2; e.stats2 = a.stats2; e.count2 = count; e.eps2 = GLN_EPS;
    e.rowsc = b.rowsc; e.rowacc = b.rowsc;
    e.g1 = a.g1; e.be1 = a.be1; e.bd = a.bd; e.wd = a.wd; e.dil = a.dil;
    p.ws = a.ws; p.ws_bytes = a.ws_bytes;
    if (int rc = launch_gemm_wx(p, true, 0, 10, st)) return rc;
  }
  {  // B3: depthwise conv + gLN1 + PReLU_1 backward
    StBwdP p{a.n, a.H, a.T, a.dil, a.ld, a.u, b.dd, b.du, a.a1, a.g1, a.be1, a.wd, a.stats1, count, b.rowsc,
             b.dg1, b.dbe1, b.dwd, b.dbd, b.da1, b.sdu};
    if (int rc = launch_dw_bwd(p, st)) return rc;
  }
  {  // B4: dx = g + W1[:, :B]^T du
    GemmWxP p{};
    p.n = a.n; p.M = a.B; p.Kd = a.H; p.T = a.T;
    p.W = a.W1; p.ldw = a.ldw1; p.X = b.du; p.ldx = a.ld; p.bsx = (int64_t)a.H * a.ld;
    p.ep.Y = b.dx; p.ep.ldy = a.ld; p.ep.bsy = (int64_t)a.B * a.ld;
    p.ep.R = b.gout; p.ep.ldr = a.ld; p.ep.bsr = (int64_t)a.B * a.ld;
    p.ws = a.ws; p.ws_bytes = a.ws_bytes;
    if (int rc = launch_gemm_wx(p, true, 0, 2, st)) return rc;
  }
  {  // dW1[:, :B] += sum_n sum_t du x^T
    GemmDwP p{};
    p.n = a.n; p.M = a.H; p.N = a.B; p.T = a.T;
    p.A = b.du; p.lda = a.ld; p.bsa = (int64_t)a.H * a.ld;
    p.B = a.x; p.ldb = a.ld; p.bsb = (int64_t)a.B * a.ld;
    p.C = b.dW1; p.ldc = a.ldw1; p.per_row = 0;
    if (int rc = launch_gemm_dw(p, 0, st)) return rc;
  }
  {
    TailP p{a.n, a.B, a.H, a.E, a.ldw1, b.sdu, b.rowsc, a.W1, a.aux, b.db1, b.da2, b.daux, b.dW1};
    int work = a.H;
    if (a.E > 0) work = max(work, max(a.n * a.E, a.H * a.E));
    tcn_bwd_tail_kernel<<<cdiv(work, 256), 256, 0, st>>>(p);
    WB_LAUNCH_CHECK("tcn_bwd_tail");
  }
  return 0;
}

extern "C" int wesep_b200_rowsum(const WesepRowSumArgs* a, void* stream) {
  if ((a->ld & 3) || !aligned16(a->x)) return fail(-1, "rowsum: alignment");
  int rows = a->n * a->C;
  rowsum_kernel<<<cdiv((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(a->x, a->ld, rows, a->T, a->out);
  WB_LAUNCH_CHECK("rowsum");
  return 0;
}
