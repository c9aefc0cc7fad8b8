// Persistent cluster kernels for the LSTM recurrence of pBSRNN's ResRNN (wesep/models/bsrnn.py:25-46).
//
// The recurrence a_t = P_t + W_hh . h_{t-1} (P = input projection + biases, one big GEMM done beforehand) is serial in t
// but independent across sequences, so ONE launch runs all S steps of both directions:
//   * a cluster of C = Hd / 32 CTAs owns a group of NQ = 64 sequences of one direction for the whole sequence;
//   * CTA `rank` owns hidden units [32 rank, 32 rank + 32) = 128 gate rows (i | f | g | o blocks of 32) and keeps ITS
//     [128 x Hd] slice of W_hh resident in shared memory for all steps, split into fp16 hi + lo with a per-row power-of-two
//     scale (22 mantissa bits; the scale is undone exactly in the epilogue);
//   * per step it issues D[128 gate rows][64 seqs] = W_slice . h_{t-1} on tcgen05 (kind::f16, M = 128, N = 64, K = Hd,
//     three products hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM), adds P_t, runs the cell for its 32 units and
//     writes the gate activations / c_t / h_t the backward needs;
//   * h_t (|h| < 1, scaled by 2^12 and split into fp16 hi + lo) is the next step's B operand: the CTA writes its
//     [64 seqs x 32 units] slab into its own operand buffer and bulk-copies it (cp.async.bulk shared::cta ->
//     shared::cluster, completing transaction bytes on the receiver's mbarrier) into the other C - 1 CTAs — the
//     all-gather of h over distributed shared memory; a multicast tcgen05.commit tells every CTA of the cluster when a
//     CTA's MMAs have finished reading its operand buffer, so it may be overwritten.
// The backward (BPTT) mirrors it with W_hh^T: each CTA contracts over ITS 128 gate rows (K = 128, M = Hd) and the
// partial dh is reduce-scattered over DSMEM (bf16 hi + lo operands: gradients need the fp32 exponent range).
//
// Tensor layout (time-major act tensors of ops.LstmTmFn): G [S][8 Hd][ld] (forward direction rows [0, 4Hd), reverse
// [4Hd, 8Hd); within a direction i | f | g | o blocks of Hd rows), H / C [S][2 Hd][ld]; columns = sequences.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace wb {
namespace lr {
using namespace tcx;

constexpr int NQ = 64;                 // sequences per cluster
constexpr int ROWS = 128;              // gate rows per CTA (4 gates x 32 hidden units)
constexpr int MAXC = 8;                // cluster size = Hd / 32
constexpr int THREADS = 160;           // warps 0-3: epilogue / cell (TMEM lane quarters), warp 4: MMA issuer + TMEM owner
constexpr int SLAB_W = ROWS * 64;      // 8192 B: one K block (32 hidden units) of the weight slice, hi or lo
constexpr int SLAB_H = NQ * 64;        // 4096 B: one K block of h, hi or lo
constexpr float H_SCALE = 4096.f;      // h in (-1, 1) -> fp16 hi + lo of 4096 h
constexpr int OFF_WHI = 0;
constexpr int OFF_WLO = MAXC * SLAB_W;                 // 65536
constexpr int OFF_H = 2 * MAXC * SLAB_W;               // 131072: C x [hi slab | lo slab]
constexpr int OFF_STG = OFF_H + MAXC * 2 * SLAB_H;     // 196608: [128 rows][64 seqs] fp32, 16-byte chunks XOR-swizzled by row
constexpr int OFF_RS = OFF_STG + ROWS * NQ * 4;        // 229376: per-row descale
constexpr int OFF_BAR = OFF_RS + ROWS * 4;             // 229888
constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;        // + alignment slack
constexpr uint32_t IDESC_FWD = idesc_f16(128, NQ, 0);

struct FwdParams {
  float* G; float* H; float* Cs;
  const float* Whh[2];
  int S, Q, Hd, C;
  int64_t ld, bsG, bsH;
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 1 - 2 / (1 + e^{2x}): two MUFU ops, absolute error ~1e-7 (saturates correctly at +-inf)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ void named_sync_epi() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1) lstm_rec_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, Hd = p.Hd;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid & 1, grp = cid >> 1;
  const int q0 = grp * NQ;
  const uint32_t bar_hfull = base + OFF_BAR, bar_hfree = base + OFF_BAR + 8, bar_acc = base + OFF_BAR + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + OFF_BAR + 24);
  float* rs = reinterpret_cast<float*>(gbase + OFF_RS);

  if (tid == 0) {
    mbar_init(bar_hfull, 1);
    mbar_init(bar_hfree, C);
    mbar_init(bar_acc, 1);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc<64>(smem_u32(tmem_slot));

  // ---- resident weight slice: local row r = gate * 32 + j  <->  W_hh row gate * Hd + 32 rank + j; fp16 hi / lo with a
  //      per-row scale 2^k such that max |w| 2^k is in [2^13, 2^14)
  {
    const float* W = p.Whh[dir];
    for (int r = warp; r < ROWS; r += THREADS / 32) {
      const float* wr = W + ((int64_t)(r >> 5) * Hd + 32 * rank + (r & 31)) * Hd;
      float w[MAXC];
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        w[i] = i < C ? __ldg(wr + 32 * i + lane) : 0.f;
        mx = fmaxf(mx, fabsf(w[i]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int e = 0;
      if (mx > 0.f) frexpf(mx, &e);
      e = max(e, -100);
      const float s = ldexpf(1.f, 14 - e);
      if (lane == 0) rs[r] = ldexpf(1.f, e - 14 - 12);   // 1 / (s * H_SCALE)
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        if (i < C) {
          const float ws = w[i] * s;
          const __half hi = __float2half_rn(ws);
          const __half lo = __float2half_rn(ws - __half2float(hi));
          const uint32_t off = i * SLAB_W + sw64_off(r, lane >> 3) + (lane & 7) * 2;
          *reinterpret_cast<__half*>(gbase + OFF_WHI + off) = hi;
          *reinterpret_cast<__half*>(gbase + OFF_WLO + off) = lo;
        }
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // every CTA's barriers are initialised before any remote traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int S = p.S;

  if (warp == 4) {
    // ================================================= MMA issuer
    if (lane == 0) {
      const uint16_t mask = (uint16_t)((1u << C) - 1u);
      for (int t = 1; t < S; ++t) {
        mbar_wait(bar_hfull, (t - 1) & 1);      // h_{t-1}: own slab written + C - 1 slabs landed
        tc_fence_after();
        for (int kb = 0; kb < C; ++kb) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t a_hi = desc_k_sw64(base + OFF_WHI + kb * SLAB_W + half * 32);
            const uint64_t a_lo = desc_k_sw64(base + OFF_WLO + kb * SLAB_W + half * 32);
            const uint64_t b_hi = desc_k_sw64(base + OFF_H + kb * 2 * SLAB_H + half * 32);
            const uint64_t b_lo = desc_k_sw64(base + OFF_H + kb * 2 * SLAB_H + SLAB_H + half * 32);
            tc_mma_f16(tmem_base, a_lo, b_hi, IDESC_FWD, (kb | half) ? 1u : 0u);
            tc_mma_f16(tmem_base, a_hi, b_lo, IDESC_FWD, 1u);
            tc_mma_f16(tmem_base, a_hi, b_hi, IDESC_FWD, 1u);
          }
        }
        tc_commit(bar_acc);                     // accumulator ready for the epilogue
        tc_commit_mc(bar_hfree, mask);          // and this CTA has finished reading its h buffer: tell the whole cluster
      }
    }
    __syncwarp();
  } else {
    // ================================================= epilogue / cell warps
    const int sl = tid & 63, uh = tid >> 6;     // phase 2: this thread = sequence sl, hidden units [16 uh, 16 uh + 16) of the CTA
    const int q = q0 + sl;
    const bool inb = q < p.ld, valid = q < p.Q;
    const int row_g0 = dir * 4 * Hd + 32 * (int)rank + 16 * uh;     // G row of (gate 0, unit 0 of this thread)
    const int row_h0 = dir * Hd + 32 * (int)rank + 16 * uh;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    const float* stg = reinterpret_cast<const float*>(gbase + OFF_STG);
    for (int t = 0; t < S; ++t) {
      const int s = dir ? S - 1 - t : t;
      float* Gs = p.G + (int64_t)s * p.bsG + (int64_t)row_g0 * p.ld + q;
      float pre[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int u = 0; u < 16; ++u) pre[g][u] = valid ? __ldcs(Gs + ((int64_t)g * Hd + u) * p.ld) : 0.f;
      if (t > 0) {
        mbar_wait(bar_acc, (t - 1) & 1);
        tc_fence_after();
        {   // phase 1: lane = gate row; TMEM -> registers -> descale -> staging (so that phase 2 can read by sequence)
          const int r = warp * 32 + lane;
          uint32_t acc[64];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
          tc_ld32_nowait(taddr, acc);
          tc_ld32_nowait(taddr + 32, acc + 32);
          tc_ld_wait();
          const float sc = rs[r];
          uint8_t* row = gbase + OFF_STG + r * (NQ * 4);
#pragma unroll
          for (int ch = 0; ch < 16; ++ch) {
            float4 v;
            v.x = __uint_as_float(acc[4 * ch + 0]) * sc;
            v.y = __uint_as_float(acc[4 * ch + 1]) * sc;
            v.z = __uint_as_float(acc[4 * ch + 2]) * sc;
            v.w = __uint_as_float(acc[4 * ch + 3]) * sc;
            *reinterpret_cast<float4*>(row + ((ch ^ (r & 15)) << 4)) = v;
          }
        }
        tc_fence_before();
        // every CTA of the cluster has finished the MMAs of step t => all copies of h_{t-1} (mine included) have been
        // consumed: my slab and the remote buffers may be overwritten with h_t
        if (tid == 0) mbar_wait(bar_hfree, (t - 1) & 1);
        named_sync_epi();
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int r = g * 32 + 16 * uh + u;
            pre[g][u] += stg[r * NQ + ((((sl >> 2) ^ (r & 15))) << 2) + (sl & 3)];
          }
      }
      // ---- cell (nn.LSTM gate order i | f | g | o)
      float hq[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float i_ = sigmoid_fast(pre[0][u]), f_ = sigmoid_fast(pre[1][u]);
        const float g_ = tanh_fast(pre[2][u]), o_ = sigmoid_fast(pre[3][u]);
        const float cn = valid ? fmaf(f_, c[u], i_ * g_) : 0.f;
        c[u] = cn;
        const float h_ = valid ? o_ * tanh_fast(cn) : 0.f;
        hq[u] = h_;
        pre[0][u] = valid ? i_ : 0.f; pre[1][u] = valid ? f_ : 0.f; pre[2][u] = valid ? g_ : 0.f; pre[3][u] = valid ? o_ : 0.f;
      }
      if (inb) {
        float* Hs = p.H + (int64_t)s * p.bsH + (int64_t)row_h0 * p.ld + q;
        float* Cc = p.Cs + (int64_t)s * p.bsH + (int64_t)row_h0 * p.ld + q;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int u = 0; u < 16; ++u) __stcs(Gs + ((int64_t)g * Hd + u) * p.ld, pre[g][u]);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          __stcs(Cc + (int64_t)u * p.ld, c[u]);
          Hs[(int64_t)u * p.ld] = hq[u];
        }
      }
      if (t + 1 < S) {
        // ---- h_t -> fp16 hi / lo into my slab (K block `rank`) of my own operand buffer, then all-gather it
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = hq[2 * j] * H_SCALE, b = hq[2 * j + 1] * H_SCALE;
          const __half2 h2 = __floats2half2_rn(a, b);
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        uint8_t* slab = gbase + OFF_H + rank * (2 * SLAB_H);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t off = sw64_off(sl, 2 * uh + e);
          *reinterpret_cast<uint4*>(slab + off) = make_uint4(hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
          *reinterpret_cast<uint4*>(slab + SLAB_H + off) = make_uint4(lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
        }
        fence_proxy_async();
        named_sync_epi();
        if (tid == 0) {
          const uint32_t src = base + OFF_H + rank * (2 * SLAB_H);
          if (C > 1) mbar_arrive_expect_tx(bar_hfull, (uint32_t)(C - 1) * 2 * SLAB_H);
          else mbar_arrive(bar_hfull);
          for (int peer = 0; peer < C; ++peer)
            if (peer != (int)rank) bulk_copy_s2c(mapa(src, peer), src, 2 * SLAB_H, mapa(bar_hfull, peer));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody exits while a peer may still copy into / arrive on its shared memory
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}


// ================================================================================================ backward (BPTT)
// Iterating the steps in the reverse of the forward order, per step t:
//   dh_t = dH_t (from the layers above) + W_hh^T . da_{t+1}          (the second term = the recurrent gradient)
//   cell backward -> da_t = d(pre-activations) [4 gates] (written over the saved activations in G), dc carried in registers
// CTA `rank` owns the 128 gate rows of ITS 32 hidden units, so its share of W_hh^T . da is a K = 128 contraction with the
// da it has just computed (no operand exchange): D[Hd hidden][64 seqs] = A[m][k] . B[seq][k], A = W_hh[rows of this CTA]^T
// resident in shared memory (bf16 hi + lo: gradients need the fp32 exponent range, 16 mantissa bits are plenty), B = da
// (bf16 hi + lo, written by the cell threads).  The partial dh is then reduce-scattered: TMEM lanes [32 w, 32 w + 32) of
// M block mb are exactly the hidden units of CTA 4 mb + w, whose receive buffer gets them with coalesced
// st.shared::cluster stores; every CTA sums the C partials for its 32 units in a fixed order (deterministic).
constexpr int B_OFF_AHI = 0;                              // 4 K slabs x [256 rows x 64 B]
constexpr int B_SLAB_A = 256 * 64;                        // 16384
constexpr int B_OFF_ALO = 4 * B_SLAB_A;                   // 65536
constexpr int B_OFF_B = 8 * B_SLAB_A;                     // 131072: 4 K slabs x [hi | lo] x [64 x 64 B]
constexpr int B_OFF_R = B_OFF_B + 4 * 2 * SLAB_H;         // 163840: C x 8 KB partial-sum slabs
constexpr int B_R_BYTES = 32 * NQ * 4;                    // 8192
constexpr int B_OFF_BAR = B_OFF_R + MAXC * B_R_BYTES;     // 229376
constexpr int B_SMEM_BYTES = B_OFF_BAR + 64 + 1024;
constexpr uint32_t IDESC_BWD = idesc_f16(128, NQ, 1);

struct BwdParams {
  float* G; const float* Cs; const float* dH;
  const float* Whh[2];
  int S, Q, Hd, C;
  int64_t ld, bsG, bsH;
};

__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h2);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__global__ void __launch_bounds__(THREADS, 1) lstm_rec_bwd_kernel(const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, Hd = p.Hd, S = p.S;
  const int MB = (Hd + 127) / 128;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid & 1, grp = cid >> 1;
  const int q0 = grp * NQ;
  const uint32_t bar_bfull = base + B_OFF_BAR, bar_acc = base + B_OFF_BAR + 8, bar_rfull = base + B_OFF_BAR + 16,
                 bar_rfree = base + B_OFF_BAR + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + B_OFF_BAR + 32);

  if (tid == 0) {
    mbar_init(bar_bfull, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_rfull, 32 * C);
    mbar_init(bar_rfree, C);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc<128>(smem_u32(tmem_slot));

  // ---- resident A = W_hh[gate rows of this CTA]^T: A[m][k], k = gate * 32 + j  <->  W_hh[gate * Hd + 32 rank + j][m];
  //      rows m >= Hd of the last M block are zero
  {
    const float* W = p.Whh[dir];
    const int Mrows = MB * 128;
    for (int k = warp; k < ROWS; k += THREADS / 32) {
      const float* wr = W + ((int64_t)(k >> 5) * Hd + 32 * rank + (k & 31)) * Hd;
      const int j = k & 31;
      uint8_t* sl_hi = gbase + B_OFF_AHI + (k >> 5) * B_SLAB_A;
      uint8_t* sl_lo = gbase + B_OFF_ALO + (k >> 5) * B_SLAB_A;
      for (int m = lane; m < Mrows; m += 32) {
        const float w = m < Hd ? __ldg(wr + m) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const uint32_t off = sw64_off(m, j >> 3) + (j & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sl_hi + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sl_lo + off) = lo;
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================================================= MMA issuer: one product per step except the last processed one
    if (lane == 0) {
      for (int n = 0; n + 1 < S; ++n) {
        mbar_wait(bar_bfull, n & 1);
        tc_fence_after();
        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t a_off = kb * B_SLAB_A + mb * (128 * 64) + half * 32;
              const uint64_t a_hi = desc_k_sw64(base + B_OFF_AHI + a_off);
              const uint64_t a_lo = desc_k_sw64(base + B_OFF_ALO + a_off);
              const uint64_t b_hi = desc_k_sw64(base + B_OFF_B + kb * 2 * SLAB_H + half * 32);
              const uint64_t b_lo = desc_k_sw64(base + B_OFF_B + kb * 2 * SLAB_H + SLAB_H + half * 32);
              const uint32_t d = tmem_base + mb * NQ;
              tc_mma_f16(d, a_lo, b_hi, IDESC_BWD, (kb | half) ? 1u : 0u);
              tc_mma_f16(d, a_hi, b_lo, IDESC_BWD, 1u);
              tc_mma_f16(d, a_hi, b_hi, IDESC_BWD, 1u);
            }
          }
        }
        tc_commit(bar_acc);
      }
    }
    __syncwarp();
  } else {
    const int sl = tid & 63, uh = tid >> 6;
    const int q = q0 + sl;
    const bool inb = q < p.ld, valid = q < p.Q;
    const int row_g0 = dir * 4 * Hd + 32 * (int)rank + 16 * uh;
    const int row_h0 = dir * Hd + 32 * (int)rank + 16 * uh;
    float dc[16], c_cur[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { dc[u] = 0.f; c_cur[u] = 0.f; }
    {   // c of the first processed step (t = S - 1)
      const int s = dir ? 0 : S - 1;
      const float* Cc = p.Cs + (int64_t)s * p.bsH + (int64_t)row_h0 * p.ld + q;
#pragma unroll
      for (int u = 0; u < 16; ++u) c_cur[u] = valid ? __ldcs(Cc + (int64_t)u * p.ld) : 0.f;
    }
    for (int n = 0; n < S; ++n) {
      const int t = S - 1 - n;                                  // forward processing index of this step
      const int s = dir ? S - 1 - t : t;
      const int sp = dir ? S - t : t - 1;                       // time index of the step processed before t in the forward
      float* Gs = p.G + (int64_t)s * p.bsG + (int64_t)row_g0 * p.ld + q;
      const float* dHs = p.dH + (int64_t)s * p.bsH + (int64_t)row_h0 * p.ld + q;
      float act[4][16], dh[16], c_prev[16];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int u = 0; u < 16; ++u) act[g][u] = valid ? __ldcs(Gs + ((int64_t)g * Hd + u) * p.ld) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) dh[u] = valid ? __ldcs(dHs + (int64_t)u * p.ld) : 0.f;
      if (t > 0) {
        const float* Cp = p.Cs + (int64_t)sp * p.bsH + (int64_t)row_h0 * p.ld + q;
#pragma unroll
        for (int u = 0; u < 16; ++u) c_prev[u] = valid ? __ldcs(Cp + (int64_t)u * p.ld) : 0.f;
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) c_prev[u] = 0.f;
      }
      if (n > 0) {
        // ---- recurrent gradient: sum of the C partial products of round n - 1 for my 16 units
        mbar_wait_cluster(bar_rfull, (n - 1) & 1);
        const float* R = reinterpret_cast<const float*>(gbase + B_OFF_R);
        const int q4 = sl >> 2;
        for (int src = 0; src < C; ++src) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int ju = 16 * uh + u;
            dh[u] += R[src * (B_R_BYTES / 4) + q4 * 128 + ((ju ^ (q4 & 7)) << 2) + (sl & 3)];
          }
        }
      }
      // ---- cell backward
      float da[4][16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float i_ = act[0][u], f_ = act[1][u], g_ = act[2][u], o_ = act[3][u];
        const float tc = tanh_fast(c_cur[u]);
        const float dhu = dh[u];
        const float dcu = fmaf(dhu * o_, 1.f - tc * tc, dc[u]);
        da[0][u] = dcu * g_ * i_ * (1.f - i_);
        da[1][u] = dcu * c_prev[u] * f_ * (1.f - f_);
        da[2][u] = dcu * i_ * (1.f - g_ * g_);
        da[3][u] = dhu * tc * o_ * (1.f - o_);
        dc[u] = dcu * f_;
        c_cur[u] = c_prev[u];
      }
      if (inb) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int u = 0; u < 16; ++u) Gs[((int64_t)g * Hd + u) * p.ld] = valid ? da[g][u] : 0.f;
      }
      if (n + 1 < S) {
        // ---- da -> bf16 hi / lo B operand: row = sequence, K index = gate * 32 + unit
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_bf16x2(da[g][2 * j], da[g][2 * j + 1], hi[j], lo[j]);
          uint8_t* slab = gbase + B_OFF_B + g * (2 * SLAB_H);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t off = sw64_off(sl, 2 * uh + e);
            *reinterpret_cast<uint4*>(slab + off) = make_uint4(hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
            *reinterpret_cast<uint4*>(slab + SLAB_H + off) = make_uint4(lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
          }
        }
        fence_proxy_async();
        named_sync_epi();
        if (tid == 0) {
          mbar_arrive(bar_bfull);
          if (n > 0)                                   // round n - 1 has been read by all my threads: its senders may reuse R
            for (int peer = 0; peer < C; ++peer) mbar_arrive_remote_release(mapa(bar_rfree, peer));
        }
        // ---- phase A: my partial W^T da -> the CTAs that own those hidden units
        mbar_wait(bar_acc, n & 1);
        tc_fence_after();
        if (n > 0) mbar_wait_cluster(bar_rfree, (n - 1) & 1);
        for (int mb = 0; mb < MB; ++mb) {
          const int m0 = mb * 128 + warp * 32;         // hidden units of this warp's TMEM lanes (warp-uniform)
          if (m0 < Hd) {
            uint32_t acc[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NQ;
            tc_ld32_nowait(taddr, acc);
            tc_ld32_nowait(taddr + 32, acc + 32);
            tc_ld_wait();
            const uint32_t dst = (uint32_t)(m0 >> 5);
            const uint32_t rbase = mapa(base + B_OFF_R + rank * B_R_BYTES, dst);
#pragma unroll
            for (int q4 = 0; q4 < 16; ++q4) {
              float4 v;
              v.x = __uint_as_float(acc[4 * q4 + 0]); v.y = __uint_as_float(acc[4 * q4 + 1]);
              v.z = __uint_as_float(acc[4 * q4 + 2]); v.w = __uint_as_float(acc[4 * q4 + 3]);
              st_cluster_v4(rbase + q4 * 512 + ((lane ^ (q4 & 7)) << 4), v);
            }
            mbar_arrive_remote_release(mapa(bar_rfull, dst));
          }
        }
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace lr
}  // namespace wb

using namespace wb;

static int check_rec(const WesepLstmRecArgs* a, bool bwd) {
  if (a->S <= 0 || a->Q <= 0 || a->Hd <= 0) return fail(-1, "lstm_rec: empty shape");
  if (a->Hd % 32 || a->Hd / 32 > lr::MAXC) return fail(-2, "lstm_rec: hidden size must be 32 * (1..8)");
  if ((a->ld & 3) || a->ld < a->Q) return fail(-1, "lstm_rec: ld must be a multiple of 4 and >= Q");
  if (!a->G || !a->H || !a->C || !a->Whh_f || !a->Whh_r) return fail(-1, "lstm_rec: null pointer");
  if (a->bsG < 8 * (int64_t)a->Hd * a->ld || a->bsH < 2 * (int64_t)a->Hd * a->ld) return fail(-1, "lstm_rec: step strides");
  if (bwd && !a->dH) return fail(-1, "lstm_rec_bwd: dH missing");
  return 0;
}

extern "C" int wesep_b200_lstm_rec_supported(int Hd) { return (Hd > 0 && Hd % 32 == 0 && Hd / 32 <= lr::MAXC) ? 1 : 0; }

extern "C" int wesep_b200_lstm_rec_fwd(const WesepLstmRecArgs* a, void* stream) {
  if (int rc = check_rec(a, false)) return rc;
  lr::FwdParams p{};
  p.G = a->G; p.H = a->H; p.Cs = a->C;
  p.Whh[0] = a->Whh_f; p.Whh[1] = a->Whh_r;
  p.S = a->S; p.Q = a->Q; p.Hd = a->Hd; p.C = a->Hd / 32;
  p.ld = a->ld; p.bsG = a->bsG; p.bsH = a->bsH;
  const int groups = cdiv(a->Q, lr::NQ);
  WB_CUDA(cudaFuncSetAttribute(lr::lstm_rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lr::SMEM_BYTES));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * groups * p.C));
  cfg.blockDim = dim3(lr::THREADS);
  cfg.dynamicSmemBytes = lr::SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  WB_CUDA(cudaLaunchKernelEx(&cfg, lr::lstm_rec_fwd_kernel, p));
  WB_LAUNCH_CHECK("lstm_rec_fwd");
  return 0;
}

extern "C" int wesep_b200_lstm_rec_bwd(const WesepLstmRecArgs* a, void* stream) {
  if (int rc = check_rec(a, true)) return rc;
  lr::BwdParams p{};
  p.G = a->G; p.Cs = a->C; p.dH = a->dH;
  p.Whh[0] = a->Whh_f; p.Whh[1] = a->Whh_r;
  p.S = a->S; p.Q = a->Q; p.Hd = a->Hd; p.C = a->Hd / 32;
  p.ld = a->ld; p.bsG = a->bsG; p.bsH = a->bsH;
  const int groups = cdiv(a->Q, lr::NQ);
  WB_CUDA(cudaFuncSetAttribute(lr::lstm_rec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lr::B_SMEM_BYTES));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * groups * p.C));
  cfg.blockDim = dim3(lr::THREADS);
  cfg.dynamicSmemBytes = lr::B_SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  WB_CUDA(cudaLaunchKernelEx(&cfg, lr::lstm_rec_bwd_kernel, p));
  WB_LAUNCH_CHECK("lstm_rec_bwd");
  return 0;
}
