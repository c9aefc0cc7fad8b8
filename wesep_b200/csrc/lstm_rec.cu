// Persistent cluster kernels for the LSTM recurrence of pBSRNN's ResRNN (wesep/models/bsrnn.py:25-46).
//
// The recurrence a_t = P_t + W_hh . h_{t-1} (P = input projection + biases, one big GEMM done beforehand) is serial in t
// but independent across sequences, so ONE launch runs all S steps of both directions:
//   * a cluster of C = Hd / 32 CTAs owns NG groups of NQ = 64 sequences of one direction for the whole sequence;
//   * CTA `rank` owns hidden units [32 rank, 32 rank + 32) = 128 gate rows (i | f | g | o blocks of 32).  ITS [128 x Hd]
//     slice of W_hh stays resident in TENSOR MEMORY for all steps as the A operand of the step product (fp16 hi + lo with a
//     per-row power-of-two scale: 22 mantissa bits, undone exactly in the epilogue) — shared memory then only holds the
//     h operand, and the MMA reads 2 KB instead of 6 KB of shared memory per instruction;
//   * per step and group: D[128 gate rows][NQ seqs] = W_slice . h_{t-1} on tcgen05 (kind::f16, M = 128, N = NQ, K = Hd,
//     three products hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM); the 8 cell warps move D through a swizzled staging
//     tile (lane = gate row -> lane = sequence), add P_t, run the cell for the CTA's 32 units and write the gate
//     activations / c_t / h_t the backward needs (all global accesses coalesced along the sequence axis);
//   * h_t (|h| < 1, scaled by 2^12, fp16 hi + lo) is the next step's B operand: the CTA writes its [NQ x 32 units] slab into
//     its own operand buffer and bulk-copies it (cp.async.bulk shared::cta -> shared::cluster, completing transaction
//     bytes on the receiver's mbarrier) into the other C - 1 CTAs — the all-gather of h over distributed shared memory; a
//     multicast tcgen05.commit tells every CTA of the cluster when a CTA's MMAs have finished reading its operand buffer;
//   * with NG = 2 the two groups are independent recurrences interleaved on the same CTA: the MMAs of one group run
//     while the cell warps work on the other (the per-step dependency chain MMA -> cell -> exchange is latency-bound).
// The backward (BPTT) mirrors it with W_hh^T (bf16 hi + lo in tensor memory: gradients need the fp32 exponent range):
// each CTA contracts over ITS 128 gate rows (K = 128, M = Hd) with the d(pre-activations) it has just computed, and the
// partial dh is reduce-scattered over DSMEM (coalesced st.shared::cluster), summed in a fixed order.
//
// Tensor layout (time-major act tensors of ops.LstmTmFn): G [S][8 Hd][ld] (forward direction rows [0, 4Hd), reverse
// [4Hd, 8Hd); within a direction i | f | g | o blocks of Hd rows), H / C / dH [S][2 Hd][ld]; columns = sequences.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace wb {
namespace lr {
using namespace tcx;

constexpr int MAXC = 8;                // cluster size = Hd / 32
constexpr int NW = 16;                 // cell warps: 512 threads = 32 hidden units x 16 sequence quads
constexpr int THREADS = 32 * NW + 64;  // warps 0-15: TMEM drain / cell, warp 16: MMA issuer + TMEM owner, warp 17: publisher
constexpr int MMA_WARP = NW, PUB_WARP = NW + 1;   // (576 threads cost the same registers as 544: allocation is per 4 warps)
constexpr float H_SCALE = 4096.f;      // h in (-1, 1) -> fp16 hi + lo of 4096 h
constexpr float LOG2E = 1.4426950408889634f;

__host__ __device__ constexpr uint32_t pow2_cols(int need) {
  return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
}

constexpr int NQ = 64;                 // sequences per group (N of the step product)
constexpr int SLAB = 32 * 128;         // one K block (32 K rows x 64 sequences x 2 B) of a B operand, hi or lo (MN-major, SW128)

template <int C, int NG>
struct FwdCfg {
  static constexpr int HBUF = C * 2 * SLAB;          // operand buffer of one group: C x [hi slab | lo slab]
  static constexpr int OFF_H = 0;
  static constexpr int OFF_OUT = NG * HBUF;          // outgoing copy of my slab, double-buffered per group: [NG][2][hi | lo]
  static constexpr int OFF_STG = OFF_OUT + NG * 2 * 2 * SLAB;   // 2 x [128 rows][NQ] fp32, 16-byte chunks XOR-swizzled by row
  static constexpr int STG_BYTES = 128 * NQ * 4;
  static constexpr int OFF_RS = OFF_STG + 2 * STG_BYTES;
  static constexpr int OFF_BAR = OFF_RS + 512;
  static constexpr int SMEM = OFF_BAR + 128 + 1024;  // + alignment slack
  static constexpr int COL_WHI = 0, COL_WLO = 16 * C, COL_D = 32 * C;   // TMEM columns
  static constexpr uint32_t TCOLS = pow2_cols(32 * C + NG * NQ);
};

struct FwdParams {
  float* G; float* H; float* Cs;
  const float* Whh[2];
  int S, Q, Hd;
  int64_t ld, bsG, bsH;
  long long* prof;
};

// optional per-phase SM-clock stamps of CTA 0 (group 0) for steps [PROF_T0, PROF_T0 + PROF_N): prof[(t - T0) * 16 + slot]
constexpr int PROF_T0 = 8, PROF_N = 16;
#define LR_STAMP(slot)                                                                                            \
  do {                                                                                                            \
    if (do_prof && g == 0 && (unsigned)(t - PROF_T0) < (unsigned)PROF_N) p.prof[(t - PROF_T0) * 16 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// 1 - 2 / (1 + e^{2x}): two MUFU ops, absolute error ~1e-7 (saturates correctly at +-inf)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcpf(1.f + ex2f(x * (2.f * LOG2E))); }
// The four gate activations with ONE reciprocal: 1 / (1 + e_k) = (prod of the other three denominators) / (prod of all four).
// Inputs are clamped to +-20 (sigma(-20) = 2e-9, tanh(10) = 1 - 4e-9: below fp32 resolution of the results) so that the
// product of four denominators (<= (1 + e^20)^4 = 5.5e34) cannot overflow.  5 MUFU ops instead of 8.
// The arguments arrive PRE-SCALED for ex2: xi = -log2(e) a_i (same for f, o), xg = 2 log2(e) a_g (the scale is folded into
// the descale of the accumulator and one FFMA on the input projection); only the upper clamp is needed (ex2 -> 0 is fine).
constexpr float EX2_CLAMP = 20.f * LOG2E;
__device__ __forceinline__ void gates_fast(float xi, float xf, float xg, float xo, float& i_, float& f_, float& g_, float& o_) {
  const float di = 1.f + ex2f(fminf(xi, EX2_CLAMP));
  const float df = 1.f + ex2f(fminf(xf, EX2_CLAMP));
  const float dO = 1.f + ex2f(fminf(xo, EX2_CLAMP));
  const float dg = 1.f + ex2f(fminf(xg, EX2_CLAMP));
  const float pif = di * df, pog = dO * dg;
  const float r = rcpf(pif * pog);
  const float rp = r * pif, rq = r * pog;
  i_ = rq * df;
  f_ = rq * di;
  o_ = rp * dg;
  g_ = fmaf(-2.f * rp, dO, 1.f);
}

__device__ __forceinline__ void named_sync_epi() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }
// barrier 2: the cell warps only ARRIVE (they go on to the next item), warp 0 waits for all of them before it publishes
// the slab / signals the MMA thread
// (one barrier id per group: consecutive publications of DIFFERENT groups are not separated by a full barrier, and an
// early second arrival on the same id would corrupt its phase)
// All 16 cell warps arrive and go on; the publisher warp (no cell work of its own, so no cell warp becomes the critical
// path) waits for them and then publishes.
__device__ __forceinline__ void named_arrive_pub(int g) { asm volatile("bar.arrive %0, 544;\n" ::"r"(2 + g) : "memory"); }
__device__ __forceinline__ void named_sync_pub(int g) { asm volatile("bar.sync %0, 544;\n" ::"r"(2 + g) : "memory"); }
static_assert(NW == 16, "named_sync_epi / column split assume 16 cell warps");

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// D[tmem] (+)= A[tmem] . B[smem]: A = [128 lanes = rows][K = 16: 8 columns of packed 16-bit pairs]
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;\n" ::: "memory"); }

// ================================================================================================ forward
template <int C, int NG>
__global__ void __launch_bounds__(THREADS, 1) lstm_rec_fwd_kernel(const FwdParams p) {
  using K = FwdCfg<C, NG>;
  constexpr int HBUF = K::HBUF, Hd = 32 * C;
  constexpr uint32_t IDESC = idesc_f16(128, NQ, 0, 1);          // fp16 operands, B MN-major
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid & 1, cgrp = cid >> 1;
  const int S = p.S;
  const bool do_prof = p.prof != nullptr && blockIdx.x == 0 && (tid == 0 || warp == MMA_WARP);
  auto bar_hfull = [&](int g) { return base + K::OFF_BAR + 8u * g; };
  auto bar_hfree = [&](int g) { return base + K::OFF_BAR + 16u + 8u * g; };
  auto bar_acc = [&](int g) { return base + K::OFF_BAR + 32u + 8u * g; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + K::OFF_BAR + 48);
  float* rs = reinterpret_cast<float*>(gbase + K::OFF_RS);

  if (tid == 0) {
    for (int g = 0; g < NG; ++g) {
      mbar_init(bar_hfull(g), 1);
      mbar_init(bar_hfree(g), C);
      mbar_init(bar_acc(g), 1);
    }
    mbar_init_fence();
  }
  if (warp == MMA_WARP) tmem_alloc<K::TCOLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- resident weight slice -> TMEM: lane r = gate * 32 + j  <->  W_hh row gate * Hd + 32 rank + j; column c holds the
  //      fp16 pair (k = 2c, 2c + 1); hi and lo of 2^e w with max |w| 2^e in [2^13, 2^14) per row
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const float* wr = p.Whh[dir] + ((int64_t)warp * Hd + 32 * rank + lane) * Hd;
    float mx = 0.f;
    for (int k = 0; k < Hd; k += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(wr + k));
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    e = max(e, -100);
    const float s = ldexpf(1.f, 14 - e);
    rs[r] = ldexpf(1.f, e - 14 - 12) * (warp == 2 ? 2.f * LOG2E : -LOG2E);   // 1 / (s * H_SCALE) x the gate's ex2 scale
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int kc = 0; kc < C; ++kc) {                         // 32 k values -> 16 columns
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wr + 32 * kc + 4 * i));
        const float a0 = v.x * s, a1 = v.y * s, a2 = v.z * s, a3 = v.w * s;
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
        hi[2 * i] = *reinterpret_cast<const uint32_t*>(&h01);
        hi[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        lo[2 * i] = *reinterpret_cast<const uint32_t*>(&l01);
        lo[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&l23);
      }
      tc_st16(trow + K::COL_WHI + 16 * kc, hi);
      tc_st16(trow + K::COL_WLO + 16 * kc, lo);
    }
    tc_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // every CTA's barriers are initialised before any remote traffic
  tc_fence_after();

  if (warp == MMA_WARP) {
    // ================================================= MMA issuer
    if (elect_one()) {
      for (int t = 1; t < S; ++t) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          mbar_wait(bar_hfull(g), (t - 1) & 1);   // h_{t-1} of group g: own slab written + C - 1 slabs landed
          tc_fence_after();
          LR_STAMP(8);
          const uint32_t d = tmem_base + K::COL_D + g * NQ;
          const uint32_t hb = base + K::OFF_H + g * HBUF;
#pragma unroll
          for (int j = 0; j < 2 * C; ++j) {       // K step of 16 = two 8-row atoms: K block j / 2, half j % 2
            const uint64_t b_hi = desc_mn_sw128(hb + (j >> 1) * 2 * SLAB + (j & 1) * 2048);
            const uint64_t b_lo = desc_mn_sw128(hb + (j >> 1) * 2 * SLAB + SLAB + (j & 1) * 2048);
            const uint32_t a_hi = tmem_base + K::COL_WHI + 8 * j, a_lo = tmem_base + K::COL_WLO + 8 * j;
            tc_mma_f16_ts(d, a_lo, b_hi, IDESC, j ? 1u : 0u);
            tc_mma_f16_ts(d, a_hi, b_lo, IDESC, 1u);
            tc_mma_f16_ts(d, a_hi, b_hi, IDESC, 1u);
          }
          tc_commit(bar_acc(g));                  // accumulator ready for the cell warps
          LR_STAMP(9);
        }
      }
    }
    __syncwarp();
  } else if (warp < NW) {
    // ================================================= TMEM drain + cell warps.  Cell phase: thread = (hidden unit u of the CTA,
    // quad of 4 consecutive sequences) so that every global / shared access is a 16-byte vector (the cell is bound by the
    // number of LSU instructions, not by bytes)
    const int u = tid >> 4, qd = tid & 15;
    const int64_t row_g = (int64_t)(dir * 4 * Hd + 32 * (int)rank + u) * p.ld;   // gate 0 row of this unit
    const int64_t row_h = (int64_t)(dir * Hd + 32 * (int)rank + u) * p.ld;
    const int64_t gstride = (int64_t)Hd * p.ld;
    float c[NG][4];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int k = 0; k < 4; ++k) c[g][k] = 0.f;
    // (544 threads cap the kernel at 96 registers: the pre-activations of an item are loaded at its start - the latency
    // hides behind the accumulator wait, the TMEM drain and the barrier - after an L2 prefetch issued one item earlier)
    float4 nxt[4];
    auto load_pre = [&](int t, int g) {
      const int s = dir ? S - 1 - t : t;
      const int q = (cgrp * NG + g) * NQ + 4 * qd;
      const float* Gs = p.G + (int64_t)s * p.bsG + row_g + q;
#pragma unroll
      for (int gt = 0; gt < 4; ++gt)
        nxt[gt] = q < p.Q ? __ldcs(reinterpret_cast<const float4*>(Gs + gt * gstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto prefetch_pre = [&](int t, int g) {
      const int s = dir ? S - 1 - t : t;
      const int q = (cgrp * NG + g) * NQ + 4 * qd;
      const float* Gs = p.G + (int64_t)s * p.bsG + row_g + q;
      if ((qd & 7) == 0 && q < p.Q) {                         // one request per 128-byte line
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Gs + gt * gstride));
      }
    };
    uint32_t item = 0;                                        // staging tile parity
    for (int t = 0; t < S; ++t) {
      const int s = dir ? S - 1 - t : t;
#pragma unroll
      for (int g = 0; g < NG; ++g, ++item) {
        float pre[4][4];
        uint8_t* stg = gbase + K::OFF_STG + (item & 1) * K::STG_BYTES;
        load_pre(t, g);
        LR_STAMP(0);
        if (t > 0) {
          mbar_wait(bar_acc(g), (t - 1) & 1);
          tc_fence_after();
          LR_STAMP(1);
          // my MMAs of (t, g) are done, i.e. this CTA has finished reading its copy of h_{t-1}: tell every CTA of the
          // cluster (plain remote arrives: a multicast tcgen05.commit was measured to land > 1000 cycles later)
          if (warp == 1 && lane < C) mbar_arrive_remote(mapa(bar_hfree(g), lane));
          {   // phase 1: lane = gate row; TMEM -> registers -> descale -> staging (so that the cell can read by sequence)
            constexpr int NCOL = NQ / 4;               // warps w, w + 4, w + 8, w + 12 share a lane quarter and split the columns
            const int r = (warp & 3) * 32 + lane, c0 = (warp >> 2) * NCOL;
            uint32_t acc[NCOL];
            tc_ld16_nowait(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + K::COL_D + g * NQ + c0, acc);
            tc_ld_wait();
            const float sc = rs[r];
            uint8_t* row = stg + r * (NQ * 4);
#pragma unroll
            for (int ch = 0; ch < NCOL / 4; ++ch) {
              float4 v;
              v.x = __uint_as_float(acc[4 * ch + 0]) * sc;
              v.y = __uint_as_float(acc[4 * ch + 1]) * sc;
              v.z = __uint_as_float(acc[4 * ch + 2]) * sc;
              v.w = __uint_as_float(acc[4 * ch + 3]) * sc;
              *reinterpret_cast<float4*>(row + (((c0 / 4 + ch) ^ (r & 7)) << 4)) = v;
            }
          }
          tc_fence_before();
          LR_STAMP(2);
          named_sync_epi();
          LR_STAMP(3);
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) {
            const int r = gt * 32 + u;
            const float4 v = *reinterpret_cast<const float4*>(stg + r * (NQ * 4) + ((qd ^ (r & 7)) << 4));
            const float4 n4 = nxt[gt];
            const float kg = gt == 2 ? 2.f * LOG2E : -LOG2E;
            pre[gt][0] = fmaf(n4.x, kg, v.x); pre[gt][1] = fmaf(n4.y, kg, v.y);
            pre[gt][2] = fmaf(n4.z, kg, v.z); pre[gt][3] = fmaf(n4.w, kg, v.w);
          }
        } else {
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) {
            const float4 n4 = nxt[gt];
            const float kg = gt == 2 ? 2.f * LOG2E : -LOG2E;
            pre[gt][0] = n4.x * kg; pre[gt][1] = n4.y * kg; pre[gt][2] = n4.z * kg; pre[gt][3] = n4.w * kg;
          }
        }
        // ---- cell (nn.LSTM gate order i | f | g | o), global writes, h_t -> fp16 hi / lo into my slab of the operand buffer
        const int q = (cgrp * NG + g) * NQ + 4 * qd;
        const int nv = p.Q - q;                                  // valid sequences of this quad (<= 0: none)
        float hq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float i_, f_, g_, o_;
          gates_fast(pre[0][k], pre[1][k], pre[2][k], pre[3][k], i_, f_, g_, o_);
          const float cn = fmaf(f_, c[g][k], i_ * g_);
          c[g][k] = cn;
          hq[k] = o_ * tanh_fast(cn);
          pre[0][k] = i_; pre[1][k] = f_; pre[2][k] = g_; pre[3][k] = o_;
        }
        if (nv < 4) {                                            // pad columns carry zeros (cold path: last quad of the batch)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k >= nv) { c[g][k] = 0.f; hq[k] = 0.f; pre[0][k] = 0.f; pre[1][k] = 0.f; pre[2][k] = 0.f; pre[3][k] = 0.f; }
        }
        if (t + 1 < S) {
          // B operand of the next step, MN-major: K row = my unit, 4 consecutive sequences = 8 bytes
          uint32_t hi[2], lo[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float a = hq[2 * j] * H_SCALE, bb = hq[2 * j + 1] * H_SCALE;
            const __half2 h2 = __floats2half2_rn(a, bb);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(a - hf.x, bb - hf.y);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          // my own operand buffer (my MMAs of (t, g) are done: bar_acc) and the outgoing copy (buffer t & 1: the copies of
          // step t - 2 out of it completed before the hfree wait of step t - 1)
          uint8_t* slab = gbase + K::OFF_H + g * HBUF + rank * (2 * SLAB);
          uint8_t* outb = gbase + K::OFF_OUT + (g * 2 + (t & 1)) * (2 * SLAB);
          const uint32_t off = sw128_off(u, qd >> 1) + (qd & 1) * 8;
          *reinterpret_cast<uint2*>(slab + off) = make_uint2(hi[0], hi[1]);
          *reinterpret_cast<uint2*>(slab + SLAB + off) = make_uint2(lo[0], lo[1]);
          if (C > 1) {
            *reinterpret_cast<uint2*>(outb + off) = make_uint2(hi[0], hi[1]);
            *reinterpret_cast<uint2*>(outb + SLAB + off) = make_uint2(lo[0], lo[1]);
          }
        }
        LR_STAMP(5);
        if (t + 1 < S) {
          fence_proxy_async();             // (a membar: keep the prefetch below it, or it waits for it)
          named_arrive_pub(g);
          LR_STAMP(6);
          prefetch_pre(t + 1, g);          // this group's next pre-activations -> L2
        }
        // the activations / cell state / h of this item go to global memory AFTER h_t has been handed to the publisher: they
        // are off the step's critical path and overlap the next item's accumulator wait
        if (q < p.ld) {
          float* Gs = p.G + (int64_t)s * p.bsG + row_g + q;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
            __stcs(reinterpret_cast<float4*>(Gs + gt * gstride), make_float4(pre[gt][0], pre[gt][1], pre[gt][2], pre[gt][3]));
          __stcs(reinterpret_cast<float4*>(p.Cs + (int64_t)s * p.bsH + row_h + q), make_float4(c[g][0], c[g][1], c[g][2], c[g][3]));
          *reinterpret_cast<float4*>(p.H + (int64_t)s * p.bsH + row_h + q) = make_float4(hq[0], hq[1], hq[2], hq[3]);
        }
      }
    }
  }
  if (warp == PUB_WARP) {
    // ================================================= publisher: h_t of (t, g) is complete in shared memory -> all-gather it
    for (int t = 0; t + 1 < S; ++t) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        named_sync_pub(g);              // every cell thread has written its part of the slab (and fenced it for the async proxy)
        // every CTA of the cluster has finished the MMAs of (t, g) => all copies of h_{t-1} have been consumed and the
        // remote operand buffers of this group may be overwritten with h_t
        if (t > 0) mbar_wait(bar_hfree(g), (t - 1) & 1);
        const uint32_t src = base + K::OFF_OUT + (g * 2 + (t & 1)) * (2 * SLAB);
        const uint32_t dst = base + K::OFF_H + g * HBUF + rank * (2 * SLAB);
        if (lane == 0) {
          if (C > 1) mbar_arrive_expect_tx(bar_hfull(g), (uint32_t)(C - 1) * 2 * SLAB);
          else mbar_arrive(bar_hfull(g));
        }
        if (lane < C && lane != (int)rank) bulk_copy_s2c(mapa(dst, lane), src, 2 * SLAB, mapa(bar_hfull(g), lane));
        if (lane == 0 && p.prof && blockIdx.x == 0 && g == 0 && (unsigned)(t - PROF_T0) < (unsigned)PROF_N)
          p.prof[(t - PROF_T0) * 16 + 7] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody exits while a peer may still copy into / arrive on its shared memory
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<K::TCOLS>(tmem_base);
  }
}

// ================================================================================================ backward (BPTT)
// Iterating the steps in the reverse of the forward order, per step n and group g:
//   phase B: dh = dH (from the layers above) + sum over source CTAs of the partial W_hh^T . da of the previous iteration;
//            cell backward -> da = d(pre-activations) [4 gates] (written over the saved activations in G and, as bf16
//            hi / lo, into the B operand buffer: row = sequence, K = gate * 32 + unit); dc carried in registers
//   MMA    : D[Hd hidden][NQ] = A[m][k] . B[seq][k], A = W_hh[gate rows of this CTA]^T resident in TMEM, K = 128
//   phase A: TMEM lanes [32 w', 32 w' + 32) of M block mb are the hidden units of CTA 4 mb + w': coalesced
//            st.shared::cluster into that CTA's receive slab for this source, one fence + one remote arrive per warp.
template <int C, int NG>
struct BwdCfg {
  static constexpr int MB = (32 * C + 127) / 128;        // M blocks of 128 hidden units
  static constexpr int BBUF = 4 * 2 * SLAB;              // B operand of one group: 4 K blocks (gates) x [hi | lo]
  static constexpr int RSLAB = 32 * NQ * 4;              // partial sums of one source CTA for my 32 units
  static constexpr int RBUF = C * RSLAB;
  static constexpr int OFF_B = 0;
  static constexpr int OFF_R = NG * BBUF;
  static constexpr int OFF_BAR = OFF_R + NG * RBUF;
  static constexpr int SMEM = OFF_BAR + 128 + 1024;
  static constexpr int COL_AHI = 0, COL_ALO = 64 * MB, COL_D = 128 * MB;   // D of (g, mb) at COL_D + (g * MB + mb) * NQ
  static constexpr uint32_t TCOLS = pow2_cols(128 * MB + NG * MB * NQ);
};

struct BwdParams {
  float* G; const float* Cs; const float* dH;
  const float* Whh[2];
  int S, Q, Hd;
  int64_t ld, bsG, bsH;
  long long* prof;
};

__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h2);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

template <int C, int NG>
__global__ void __launch_bounds__(THREADS, 1) lstm_rec_bwd_kernel(const BwdParams p) {
  using K = BwdCfg<C, NG>;
  constexpr int MB = K::MB, Hd = 32 * C;
  constexpr uint32_t IDESC = idesc_f16(128, NQ, 1, 1);          // bf16 operands, B MN-major
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid & 1, cgrp = cid >> 1;
  const int S = p.S;
  const bool do_prof = p.prof != nullptr && blockIdx.x == 0 && (tid == 0 || warp == MMA_WARP);
  auto bar_bfull = [&](int g) { return base + K::OFF_BAR + 8u * g; };
  auto bar_acc = [&](int g) { return base + K::OFF_BAR + 16u + 8u * g; };
  auto bar_rfull = [&](int g) { return base + K::OFF_BAR + 32u + 8u * g; };
  auto bar_rfree = [&](int g) { return base + K::OFF_BAR + 48u + 8u * g; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + K::OFF_BAR + 64);

  if (tid == 0) {
    for (int g = 0; g < NG; ++g) {
      mbar_init(bar_bfull(g), 1);
      mbar_init(bar_acc(g), 1);
      mbar_init(bar_rfull(g), 2 * C);            // two sender warps (column halves) per source CTA
      mbar_init(bar_rfree(g), C);
    }
    mbar_init_fence();
  }
  if (warp == MMA_WARP) tmem_alloc<K::TCOLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- resident A = W_hh[gate rows of this CTA]^T -> TMEM: lane m (hidden unit, M block warp / 4), K index
  //      k = gate * 32 + j  <->  W_hh[gate * Hd + 32 rank + j][m]; column c holds the bf16 pair (k = 2c, 2c + 1)
  if (warp < 4 * MB) {
    const int mb = warp >> 2;
    const int m = mb * 128 + (warp & 3) * 32 + lane;
    const bool mok = m < Hd;
    const float* W = p.Whh[dir] + (int64_t)(32 * rank) * Hd + (mok ? m : 0);
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + mb * 64;
#pragma unroll 1
    for (int kc = 0; kc < 4; ++kc) {                      // gate kc: 32 k values -> 16 columns
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float w0 = mok ? __ldg(W + ((int64_t)kc * Hd + 2 * i) * Hd) : 0.f;
        const float w1 = mok ? __ldg(W + ((int64_t)kc * Hd + 2 * i + 1) * Hd) : 0.f;
        split_bf16x2(w0, w1, hi[i], lo[i]);
      }
      tc_st16(trow + K::COL_AHI + 16 * kc, hi);
      tc_st16(trow + K::COL_ALO + 16 * kc, lo);
    }
    tc_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  if (warp == MMA_WARP) {
    // ================================================= MMA issuer: one product per step except the last processed one
    if (elect_one()) {
      for (int t = 0; t + 1 < S; ++t) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          mbar_wait(bar_bfull(g), t & 1);
          tc_fence_after();
          LR_STAMP(8);
          const uint32_t bb = base + K::OFF_B + g * K::BBUF;
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t d = tmem_base + K::COL_D + (g * MB + mb) * NQ;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t b_hi = desc_mn_sw128(bb + (j >> 1) * 2 * SLAB + (j & 1) * 2048);
              const uint64_t b_lo = desc_mn_sw128(bb + (j >> 1) * 2 * SLAB + SLAB + (j & 1) * 2048);
              const uint32_t a_hi = tmem_base + K::COL_AHI + mb * 64 + 8 * j, a_lo = tmem_base + K::COL_ALO + mb * 64 + 8 * j;
              tc_mma_f16_ts(d, a_lo, b_hi, IDESC, j ? 1u : 0u);
              tc_mma_f16_ts(d, a_hi, b_lo, IDESC, 1u);
              tc_mma_f16_ts(d, a_hi, b_hi, IDESC, 1u);
            }
          }
          tc_commit(bar_acc(g));
          LR_STAMP(9);
        }
      }
    }
    __syncwarp();
  } else if (warp == PUB_WARP) {
    // ================================================= publisher: da of (t, g) is complete -> start its product; the receive
    // slabs of round t - 1 have been read by all my cell threads -> their senders may reuse them
    for (int t = 0; t + 1 < S; ++t) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        named_sync_pub(g);
        if (lane == 0) mbar_arrive(bar_bfull(g));
        if (t > 0 && lane < C) mbar_arrive_remote(mapa(bar_rfree(g), lane));
      }
    }
  } else {
    // cell threads: (hidden unit u of the CTA, quad of 4 consecutive sequences), 16-byte accesses everywhere
    const int u = tid >> 4, qd = tid & 15;
    const int64_t row_g = (int64_t)(dir * 4 * Hd + 32 * (int)rank + u) * p.ld;
    const int64_t row_h = (int64_t)(dir * Hd + 32 * (int)rank + u) * p.ld;
    const int64_t gstride = (int64_t)Hd * p.ld;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float dc[NG][4], c_cur[NG][4];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int q = (cgrp * NG + g) * NQ + 4 * qd;
      const int s = dir ? 0 : S - 1;                           // first processed step
      const float4 v = q < p.Q ? __ldcs(reinterpret_cast<const float4*>(p.Cs + (int64_t)s * p.bsH + row_h + q)) : zero4;
      c_cur[g][0] = v.x; c_cur[g][1] = v.y; c_cur[g][2] = v.z; c_cur[g][3] = v.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) dc[g][k] = 0.f;
    }
    for (int t = 0; t < S; ++t) {                                // t = backward iteration; tf = forward processing index
      const int tf = S - 1 - t;
      const int s = dir ? S - 1 - tf : tf;
      const int sp = dir ? S - tf : tf - 1;                      // time index of the step processed before tf in the forward
      // ------------------------------------------------------------------ phase B of every group
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        LR_STAMP(0);
        const int q = (cgrp * NG + g) * NQ + 4 * qd;
        const int nv = p.Q - q;
        const bool any = q < p.Q;
        float* Gs = p.G + (int64_t)s * p.bsG + row_g + q;
        float4 a4[4];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) a4[gt] = any ? __ldcs(reinterpret_cast<const float4*>(Gs + gt * gstride)) : zero4;
        const float4 dh4 = any ? __ldcs(reinterpret_cast<const float4*>(p.dH + (int64_t)s * p.bsH + row_h + q)) : zero4;
        const float4 cp4 = (any && tf > 0) ? __ldcs(reinterpret_cast<const float4*>(p.Cs + (int64_t)sp * p.bsH + row_h + q)) : zero4;
        float dh[4] = {dh4.x, dh4.y, dh4.z, dh4.w};
        const float c_prev[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
        if (t > 0) {
          // ---- recurrent gradient: sum of the C partial products of round t - 1 for my unit
          mbar_wait_cluster(bar_rfull(g), (t - 1) & 1);
          LR_STAMP(1);
          const uint8_t* R = gbase + K::OFF_R + g * K::RBUF + qd * 512 + ((u ^ (qd & 7)) << 4);
#pragma unroll
          for (int src = 0; src < C; ++src) {
            const float4 v = *reinterpret_cast<const float4*>(R + src * K::RSLAB);
            dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
          }
        }
        LR_STAMP(2);
        // ---- cell backward, da -> global (over the activations) and -> bf16 hi / lo B operand (MN-major: K row = gate * 32 + u)
        const float ai[4] = {a4[0].x, a4[0].y, a4[0].z, a4[0].w}, af[4] = {a4[1].x, a4[1].y, a4[1].z, a4[1].w};
        const float ag[4] = {a4[2].x, a4[2].y, a4[2].z, a4[2].w}, ao[4] = {a4[3].x, a4[3].y, a4[3].z, a4[3].w};
        float da[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool ok = k < nv;                               // pad columns: zeros
          const float i_ = ai[k], f_ = af[k], g_ = ag[k], o_ = ao[k];
          const float tc = tanh_fast(c_cur[g][k]);
          const float dhu = ok ? dh[k] : 0.f;
          const float dcu = ok ? fmaf(dhu * o_, 1.f - tc * tc, dc[g][k]) : 0.f;
          da[0][k] = dcu * g_ * i_ * (1.f - i_);
          da[1][k] = dcu * c_prev[k] * f_ * (1.f - f_);
          da[2][k] = dcu * i_ * (1.f - g_ * g_);
          da[3][k] = dhu * tc * o_ * (1.f - o_);
          dc[g][k] = dcu * f_;
          c_cur[g][k] = c_prev[k];
        }
        if (q < p.ld) {
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
            *reinterpret_cast<float4*>(Gs + gt * gstride) = make_float4(da[gt][0], da[gt][1], da[gt][2], da[gt][3]);
        }
        if (t + 1 < S) {
          uint8_t* bbuf = gbase + K::OFF_B + g * K::BBUF;
          const uint32_t off = sw128_off(u, qd >> 1) + (qd & 1) * 8;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) {
            uint32_t h0, l0, h1, l1;
            split_bf16x2(da[gt][0], da[gt][1], h0, l0);
            split_bf16x2(da[gt][2], da[gt][3], h1, l1);
            *reinterpret_cast<uint2*>(bbuf + gt * 2 * SLAB + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(bbuf + gt * 2 * SLAB + SLAB + off) = make_uint2(l0, l1);
          }
        }
        LR_STAMP(3);
        if (t + 1 < S) {
          fence_proxy_async();
          named_arrive_pub(g);
        }
        LR_STAMP(4);
      }
      // ------------------------------------------------------------------ phase A of every group
      if (t + 1 < S) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          mbar_wait(bar_acc(g), t & 1);
          tc_fence_after();
          LR_STAMP(5);
          if (t > 0) mbar_wait_cluster(bar_rfree(g), (t - 1) & 1);
          LR_STAMP(6);
          const int mb = (warp >> 2) & 1, ch = warp >> 3;      // M block, column half
          const int m0 = mb * 128 + (warp & 3) * 32;            // hidden units of this warp's TMEM lanes (warp-uniform)
          if (mb < MB && m0 < Hd) {
            constexpr int NCOL = NQ / 2;
            const uint32_t dst = (uint32_t)(m0 >> 5);
            const uint32_t rbase = mapa(base + K::OFF_R + g * K::RBUF + rank * K::RSLAB, dst);
            uint32_t acc[NCOL];
            tc_ld32_nowait(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + K::COL_D + (g * MB + mb) * NQ + ch * NCOL, acc);
            tc_ld_wait();
#pragma unroll
            for (int i = 0; i < NCOL / 4; ++i) {
              const int q4 = ch * (NCOL / 4) + i;
              float4 v;
              v.x = __uint_as_float(acc[4 * i + 0]); v.y = __uint_as_float(acc[4 * i + 1]);
              v.z = __uint_as_float(acc[4 * i + 2]); v.w = __uint_as_float(acc[4 * i + 3]);
              st_cluster_v4(rbase + q4 * 512 + ((lane ^ (q4 & 7)) << 4), v);
            }
            __syncwarp();
            if (lane == 0) {
              fence_cluster();
              mbar_arrive_remote(mapa(bar_rfull(g), dst));
            }
          }
          tc_fence_before();
          LR_STAMP(7);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<K::TCOLS>(tmem_base);
  }
}

// ================================================================================================ host side
template <typename Params, typename Kern>
static int launch_cluster(Kern kern, int smem, int C, int clusters, const Params& p, cudaStream_t st) {
  WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * C));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  WB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

template <typename Kern>
static int max_clusters_of(Kern kern, int smem, int C) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(C * 64));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

// Sequences per cluster (capacity) = 64 NG.  Pick the smaller one if its cluster count fits in one wave (co-resident
// clusters, cached per hidden size); larger problems take 128 (two interleaved groups).
static int pick_capacity(int Q, int maxc, int force) {
  if (force == 64 || force == 128) return force;
  return 2 * cdiv(Q, 64) <= maxc ? 64 : 128;
}

template <int C>
static int max_clusters_c(bool bwd) {
  static int cached[2] = {0, 0};
  if (!cached[bwd]) {
    int n = bwd ? max_clusters_of(lstm_rec_bwd_kernel<C, 2>, BwdCfg<C, 2>::SMEM, C)
                : max_clusters_of(lstm_rec_fwd_kernel<C, 2>, FwdCfg<C, 2>::SMEM, C);
    cached[bwd] = n > 0 ? n : 1;
  }
  return cached[bwd];
}

template <int C>
static int run_fwd(const FwdParams& p, int force, cudaStream_t st) {
  const int cap = pick_capacity(p.Q, max_clusters_c<C>(false), force);
  const int clusters = 2 * cdiv(p.Q, cap);
  if (cap == 64) return launch_cluster(lstm_rec_fwd_kernel<C, 1>, FwdCfg<C, 1>::SMEM, C, clusters, p, st);
  return launch_cluster(lstm_rec_fwd_kernel<C, 2>, FwdCfg<C, 2>::SMEM, C, clusters, p, st);
}
template <int C>
static int run_bwd(const BwdParams& p, int force, cudaStream_t st) {
  const int cap = pick_capacity(p.Q, max_clusters_c<C>(true), force);
  const int clusters = 2 * cdiv(p.Q, cap);
  if (cap == 64) return launch_cluster(lstm_rec_bwd_kernel<C, 1>, BwdCfg<C, 1>::SMEM, C, clusters, p, st);
  return launch_cluster(lstm_rec_bwd_kernel<C, 2>, BwdCfg<C, 2>::SMEM, C, clusters, p, st);
}

}  // namespace lr
}  // namespace wb

using namespace wb;

extern "C" int wesep_b200_lstm_rec_supported(int Hd) {
  return (Hd == 32 || Hd == 64 || Hd == 128 || Hd == 192 || Hd == 256) ? 1 : 0;
}

static int check_rec(const WesepLstmRecArgs* a, bool bwd) {
  if (a->S <= 0 || a->Q <= 0 || a->Hd <= 0) return fail(-1, "lstm_rec: empty shape");
  if (!wesep_b200_lstm_rec_supported(a->Hd)) return fail(-2, "lstm_rec: hidden size must be 32, 64, 128, 192 or 256");
  if ((a->ld & 3) || a->ld < a->Q) return fail(-1, "lstm_rec: ld must be a multiple of 4 and >= Q");
  if (!a->G || !a->H || !a->C || !a->Whh_f || !a->Whh_r) return fail(-1, "lstm_rec: null pointer");
  if (!aligned16(a->Whh_f) || !aligned16(a->Whh_r)) return fail(-1, "lstm_rec: W_hh must be 16-byte aligned");
  if (a->bsG < 8 * (int64_t)a->Hd * a->ld || a->bsH < 2 * (int64_t)a->Hd * a->ld) return fail(-1, "lstm_rec: step strides");
  if (bwd && !a->dH) return fail(-1, "lstm_rec_bwd: dH missing");
  return 0;
}

// How many clusters of the (largest configuration of the) kernel can be resident at once on the current device.
extern "C" int wesep_b200_lstm_rec_max_clusters(int Hd, int bwd) {
  switch (Hd / 32) {
    case 1: return lr::max_clusters_c<1>(bwd != 0);
    case 2: return lr::max_clusters_c<2>(bwd != 0);
    case 4: return lr::max_clusters_c<4>(bwd != 0);
    case 6: return lr::max_clusters_c<6>(bwd != 0);
    case 8: return lr::max_clusters_c<8>(bwd != 0);
  }
  return 0;
}

extern "C" int wesep_b200_lstm_rec_fwd(const WesepLstmRecArgs* a, void* stream) {
  if (int rc = check_rec(a, false)) return rc;
  lr::FwdParams p{};
  p.G = a->G; p.H = a->H; p.Cs = a->C;
  p.Whh[0] = a->Whh_f; p.Whh[1] = a->Whh_r;
  p.S = a->S; p.Q = a->Q; p.Hd = a->Hd;
  p.ld = a->ld; p.bsG = a->bsG; p.bsH = a->bsH;
  p.prof = (long long*)a->prof;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = -2;
  switch (a->Hd / 32) {
    case 1: rc = lr::run_fwd<1>(p, a->seqs_per_cluster, st); break;
    case 2: rc = lr::run_fwd<2>(p, a->seqs_per_cluster, st); break;
    case 4: rc = lr::run_fwd<4>(p, a->seqs_per_cluster, st); break;
    case 6: rc = lr::run_fwd<6>(p, a->seqs_per_cluster, st); break;
    case 8: rc = lr::run_fwd<8>(p, a->seqs_per_cluster, st); break;
  }
  if (rc) return rc;
  WB_LAUNCH_CHECK("lstm_rec_fwd");
  return 0;
}

extern "C" int wesep_b200_lstm_rec_bwd(const WesepLstmRecArgs* a, void* stream) {
  if (int rc = check_rec(a, true)) return rc;
  lr::BwdParams p{};
  p.G = a->G; p.Cs = a->C; p.dH = a->dH;
  p.Whh[0] = a->Whh_f; p.Whh[1] = a->Whh_r;
  p.S = a->S; p.Q = a->Q; p.Hd = a->Hd;
  p.ld = a->ld; p.bsG = a->bsG; p.bsH = a->bsH;
  p.prof = (long long*)a->prof;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = -2;
  switch (a->Hd / 32) {
    case 1: rc = lr::run_bwd<1>(p, a->seqs_per_cluster, st); break;
    case 2: rc = lr::run_bwd<2>(p, a->seqs_per_cluster, st); break;
    case 4: rc = lr::run_bwd<4>(p, a->seqs_per_cluster, st); break;
    case 6: rc = lr::run_bwd<6>(p, a->seqs_per_cluster, st); break;
    case 8: rc = lr::run_bwd<8>(p, a->seqs_per_cluster, st); break;
  }
  if (rc) return rc;
  WB_LAUNCH_CHECK("lstm_rec_bwd");
  return 0;
}
