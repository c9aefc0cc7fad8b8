// tcgen05 (UMMA) pointwise-conv GEMM for sm_100a:  Y[n][o][t] = sum_k W[o][k] * f(X[n][k][t])  (+ epilogue)
//
//   D[M = 128 output channels][N = 256 time steps] accumulates in TMEM (fp32), operands in shared memory:
//     A = W^T tile  [k][o]  (MN-major, 32 o = 128 B per row)   pre-split into hi / lo by a prep kernel
//     B = X   tile  [k][t]  (MN-major, 32 t = 128 B per row)   TMA'd raw, split into hi / lo by transform warps
//   both with the 128-byte TMA/UMMA swizzle, so ONE layout rule covers both operands and the transform pass is a
//   layout-agnostic elementwise sweep (the swizzle only permutes 16-byte chunks inside a 128-byte row).
//   3xTF32: D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (kind::tf32, K = 8 per instruction), fp32-grade accuracy.
//
//   Warp roles (448 threads, 1 CTA/SM, persistent over tiles):
//     warps 0-7   epilogue   TMEM -> registers (tcgen05.ld 32x32b) -> fused epilogue -> global (lane = channel;
//                            warp w owns TMEM lanes 32*(w%4).. and columns 128*(w/4)..)
//     warp  8     TMA producer (4-stage ring, BK = 16) + TMEM alloc/dealloc
//     warp  9     MMA issuer (one thread), tcgen05.commit -> mbarriers
//     warps 10-17 transform  hi/lo split (+ gLN/PReLU operand prologue), fence.proxy.async
//   TMEM: 2 x 256 columns (double-buffered accumulators: epilogue of tile i overlaps the MMAs of tile i+1).
#include <cuda.h>

#include "gemm_mma.cuh"

namespace wb {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 16, TC_STAGES = 4;
constexpr int TC_THREADS = 576;   // 8 epilogue + TMA + MMA + 8 transform warps
constexpr int TC_W_TMA = 8, TC_W_MMA = 9, TC_W_XF = 10;
constexpr int TC_BOX_BYTES = 32 * TC_BK * 4;             // one TMA box: 32 floats x 16 rows = 2048 B
constexpr int TC_W_BYTES = (TC_BM / 32) * TC_BOX_BYTES;  // 8192
constexpr int TC_X_BYTES = (TC_BN / 32) * TC_BOX_BYTES;  // 16384
constexpr int TC_OFF_WHI = 0, TC_OFF_WLO = TC_W_BYTES, TC_OFF_XHI = 2 * TC_W_BYTES, TC_OFF_XLO = 2 * TC_W_BYTES + TC_X_BYTES;
constexpr int TC_STAGE_BYTES = 2 * TC_W_BYTES + 2 * TC_X_BYTES;  // 49152
constexpr int TC_TX_BYTES = 2 * TC_W_BYTES + TC_X_BYTES;         // bytes the TMA delivers per stage
constexpr int TC_MAXK = 1024;
constexpr int TC_SMEM_AUX = 2 * TC_MAXK * 4 + 256;                // sc/sh + barriers
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_SMEM_AUX + 1024;

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// TMA store of one box (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// Shared-memory matrix descriptor: MN-major 32-bit (tf32) operand.  For MN-major tf32 the only legal smem layout
// is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92; cute Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over
// 32 floats x 4 K-rows): 128-byte rows, 32-byte swizzle granules, K atoms of 4 rows (512 B).  The matching TMA mode
// is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
//   LBO = byte stride between MN atoms (next 32 floats of M/N) = one TMA box; SBO = stride between K atoms = 512 B.
//   (cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout_type[61,64)=1)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(TC_BOX_BYTES >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both MN-major, N=256, M=128.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

// ------------------------------------------------------------------------------------------ prep kernel
// hi[oa][k][oi] = tf32-truncated W[o = 32*oa + oi][k] (or W[k][o] if w_trans), lo = W - hi; rows o >= M are zero.
// Tiling by 32-channel atoms makes the CTA's 128 x 16 weight tile ONE 3-D TMA box (32, 16, 4).
__global__ void split_w_kernel(const float* __restrict__ W, int64_t ldw, int w_trans, int M, int Kd, float* __restrict__ hi,
                               float* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int M32 = (M + 31) & ~31;
  if (idx >= M32 * Kd) return;
  const int oi = idx & 31, k = (idx >> 5) % Kd, oa = (idx >> 5) / Kd;
  const int o = oa * 32 + oi;
  float w = 0.f;
  if (o < M) w = w_trans ? W[(int64_t)k * ldw + o] : W[(int64_t)o * ldw + k];
  const float h = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
  hi[idx] = h;
  lo[idx] = w - h;
}

struct TcParams {
  GemmWxP g;
  int n_ob, n_tt, n_tiles;
  int skip_hi_store;   // PRO 0 only: leave the raw fp32 tile as the "hi" operand (valid iff the MMA truncates to tf32)
  int xf_groups;       // transform warps split into this many groups (1, 2, 4); group g handles stages with it % groups == g
  int dbg;             // timing experiments only (2-CTA kernel): 1 = no epilogue loads, 2 = no epilogue stores, 4 = no transform
  int mixed;           // 2-CTA kernel: tf32 leading term + bf16 cross terms (see tc_mma_bf16_2cta) instead of 3xTF32
  const uint8_t* wimg; // mixed: bf16 weight images, 8 KB per (128-channel block, K block)
};
int g_tc_flags = 0;

// SM count of the CURRENT device (a process may drive several devices: no single function-static value)
static int sm_count(int* out) {
  static int cache[64] = {0};
  int dev = 0;
  WB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(-3, "device index out of range");
  if (!cache[dev]) WB_CUDA(cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev));
  *out = cache[dev];
  return 0;
}

// ------------------------------------------------------------------------------------------ main kernel
template <int PRO, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
    gemm_wx_tc_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                      const __grid_constant__ CUtensorMap map_x, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const GemmWxP& p = P.g;
  // 1024-byte aligned base (swizzle atoms)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  float* sc = reinterpret_cast<float*>(gbase + TC_STAGES * TC_STAGE_BYTES);
  float* sh = sc + TC_MAXK;
  const uint32_t bar0 = base + TC_STAGES * TC_STAGE_BYTES + 2 * TC_MAXK * 4;
  // barrier addresses (8 B each)
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * TC_STAGES + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (3 * TC_STAGES + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (3 * TC_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + TC_STAGES * TC_STAGE_BYTES + 2 * TC_MAXK * 4 + 8 * (3 * TC_STAGES + 4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = (p.Kd + TC_BK - 1) / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 8 / P.xf_groups);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == TC_W_TMA) {  // TMEM allocation: 512 columns (2 accumulator buffers)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TC_W_TMA) {
    // =========================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int ob = tile % P.n_ob, rest = tile / P.n_ob;
        const int tt = rest % P.n_tt, n = rest / P.n_tt;
        const int o0 = ob * TC_BM, t0 = tt * TC_BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t sb = base + s * TC_STAGE_BYTES;
          mbar_expect_tx(bar_full(s), TC_TX_BYTES);
          const int k0 = kb * TC_BK;
          // 3 TMA operations per stage (per-box overhead, not bytes, bounds the small-box variant)
          tma_load_3d(sb + TC_OFF_WHI, &map_whi, bar_full(s), 0, k0, o0 >> 5);
          tma_load_3d(sb + TC_OFF_WLO, &map_wlo, bar_full(s), 0, k0, o0 >> 5);
          tma_load_4d(sb + TC_OFF_XHI, &map_x, bar_full(s), 0, k0, t0 >> 5, n);
        }
      }
    }
  } else if (warp == TC_W_MMA) {
    // =========================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1;
        mbar_wait(bar_acce(a), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * TC_BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint32_t sb = base + s * TC_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t a_hi = make_desc_mn_sw128(sb + TC_OFF_WHI + ks * 1024);
            const uint64_t a_lo = make_desc_mn_sw128(sb + TC_OFF_WLO + ks * 1024);
            const uint64_t b_hi = make_desc_mn_sw128(sb + TC_OFF_XHI + ks * 1024);
            const uint64_t b_lo = make_desc_mn_sw128(sb + TC_OFF_XLO + ks * 1024);
            tc_mma_tf32(d_tmem, a_lo, b_hi, TC_IDESC, (kb | ks) != 0 ? 1u : 0u);
            tc_mma_tf32(d_tmem, a_hi, b_lo, TC_IDESC, 1u);
            tc_mma_tf32(d_tmem, a_hi, b_hi, TC_IDESC, 1u);
          }
          tc_commit(bar_empty(s));   // stage reusable once these MMAs have read it
        }
        tc_commit(bar_accf(a));      // accumulator complete
      }
    }
  } else if (warp >= TC_W_XF) {
    // =========================================================================== transform warps
    const int tt_id = tid - TC_W_XF * 32;  // 0..255
    const int xf_groups = P.xf_groups, xf_nthr = 256 / xf_groups;
    const int xf_gid = tt_id / xf_nthr, xf_tid = tt_id % xf_nthr;   // group of this warp, thread index inside the group
    float alpha = 1.f;
    if constexpr (PRO >= 1) alpha = p.xf.alpha ? __ldg(p.xf.alpha) : 1.f;
    uint32_t it = 0;
    int cur_n = -1;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const int n = (tile / P.n_ob) / P.n_tt;
      if constexpr (PRO >= 2) {
        if (n != cur_n) {  // per-row gLN constants: sc[k] = gamma*rstd, sh[k] = beta - gamma*mean*rstd
          // all transform warps of this CTA see the same tile sequence: sync them around the table rewrite
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          float mu = 0.f, r = 1.f;
          if (p.xf.row_stats) gln_mean_rstd(p.xf.row_stats + 2 * n, p.xf.count, p.xf.eps, mu, r);
          for (int k = tt_id; k < p.Kd; k += 256) {
            const float gm = p.xf.ch_scale ? __ldg(p.xf.ch_scale + k) : 1.f;
            const float bt = p.xf.ch_shift ? __ldg(p.xf.ch_shift + k) : 0.f;
            sc[k] = gm * r;
            sh[k] = bt - gm * mu * r;
          }
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          cur_n = n;
        }
      }
      for (int kb = 0; kb < KB; ++kb, ++it) {
        if ((int)(it % (uint32_t)xf_groups) != xf_gid) continue;   // another transform group owns this stage
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        uint8_t* xs_hi = gbase + s * TC_STAGE_BYTES + TC_OFF_XHI;
        uint8_t* xs_lo = gbase + s * TC_STAGE_BYTES + TC_OFF_XLO;
        for (int i = 0; i < TC_X_BYTES / 16 / xf_nthr; ++i) {
          const int off = (xf_tid + xf_nthr * i) * 16;
          float4 v = *reinterpret_cast<const float4*>(xs_hi + off);
          if constexpr (PRO >= 1) {
            float c = 1.f, d = 0.f;
            if constexpr (PRO >= 2) {
              const int kk = kb * TC_BK + ((off % TC_BOX_BYTES) >> 7);  // row inside the box = channel
              const bool kok = kk < p.Kd;
              c = kok ? sc[kk] : 0.f;
              d = kok ? sh[kk] : 0.f;
            }
            if constexpr (PRO == 3) {   // BatchNorm apply, then PReLU
              v.x = prelu_f(fmaf(c, v.x, d), alpha);
              v.y = prelu_f(fmaf(c, v.y, d), alpha);
              v.z = prelu_f(fmaf(c, v.z, d), alpha);
              v.w = prelu_f(fmaf(c, v.w, d), alpha);
            } else {                    // PReLU, then gLN apply
              v.x = fmaf(c, prelu_f(v.x, alpha), d);
              v.y = fmaf(c, prelu_f(v.y, alpha), d);
              v.z = fmaf(c, prelu_f(v.z, alpha), d);
              v.w = fmaf(c, prelu_f(v.w, alpha), d);
            }
          }
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
          if (PRO != 0 || !P.skip_hi_store) *reinterpret_cast<float4*>(xs_hi + off) = h;
          *reinterpret_cast<float4*>(xs_lo + off) = l;
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(s));
      }
    }
  } else {
    // =========================================================================== epilogue warps 0..7
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int chalf = warp >> 2;   // which 128-column half of the accumulator
    const EpiP& e = p.ep;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
      const int ob = tile % P.n_ob, rest = tile / P.n_ob;
      const int tt = rest % P.n_tt, n = rest / P.n_tt;
      const int o_raw = ob * TC_BM + q * 32 + lane;
      const bool o_ok = o_raw < p.M;                 // partial last channel block
      const int o = o_ok ? o_raw : p.M - 1;          // clamp for the constant loads; stores are masked
      const int t0 = tt * TC_BN;
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;

      // per-channel constants
      float bias_o = 0.f;
      if (e.bias) bias_o = __ldg(e.bias + o);
      if (e.row_bias) bias_o += __ldg(e.row_bias + (int64_t)n * p.M + o);
      float out_alpha = 1.f;
      if constexpr (EPI == 0) out_alpha = e.out_alpha ? __ldg(e.out_alpha) : 1.f;
      float mu2 = 0.f, r2 = 1.f, a2 = 1.f, gam2 = 0.f, mh = 0.f, mhy = 0.f, gam1 = 0.f, bet1 = 0.f, bdm = 0.f, w0 = 0.f, w1 = 0.f,
            w2 = 0.f;
      if constexpr (EPI == 10) {
        gln_mean_rstd(e.stats2 + 2 * n, e.count2, e.eps2, mu2, r2);
        a2 = __ldg(e.a2); gam2 = __ldg(e.g2 + o);
        mh = (float)(e.rowsc[8 * n + 0] / e.count2); mhy = (float)(e.rowsc[8 * n + 1] / e.count2);
        gam1 = __ldg(e.g1 + o); bet1 = __ldg(e.be1 + o); bdm = __ldg(e.bd + o);
        w0 = __ldg(e.wd + 3 * o); w1 = __ldg(e.wd + 3 * o + 1); w2 = __ldg(e.wd + 3 * o + 2);
      }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, sL = 0.f, sR = 0.f;
      float cA = 0.f, cB = 0.f, cC = 0.f;
      if constexpr (EPI == 10) {
        cA = r2 * gam2;
        cB = -r2 * r2 * mhy;
        cC = -r2 * mh + r2 * r2 * mhy * mu2;
      }

      mbar_wait(bar_accf(a), aph);
      tc_fence_after();
      float* yrow = e.Y + n * e.bsy + (int64_t)o * e.ldy;
#pragma unroll 1
      for (int c0 = chalf * (TC_BN / 2); c0 < (chalf + 1) * (TC_BN / 2); c0 += 32) {
        if (t0 + c0 >= p.T) break;   // warp-uniform
        const bool edge_chunk = (EPI == 10) && ((t0 + c0 < e.dil) || (t0 + c0 + 32 > p.T - e.dil));
        // issue this chunk's global operand loads before the TMEM load so their latency overlaps it
        float4 gop[8];
        if constexpr (EPI == 2 || EPI == 3 || EPI == 10) {
          const float* gsrc = (EPI == 10) ? (e.d + n * e.bsd + (int64_t)o * e.ldd) : (e.R + n * e.bsr + (int64_t)o * e.ldr);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int t = t0 + c0 + 4 * g;
            gop[g] = (t < p.T && o_ok) ? *reinterpret_cast<const float4*>(gsrc + t) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_BN + c0), r);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int t = t0 + c0 + 4 * g;
          if (t < p.T && o_ok) {
            float v[4] = {__uint_as_float(r[4 * g]) + bias_o, __uint_as_float(r[4 * g + 1]) + bias_o,
                          __uint_as_float(r[4 * g + 2]) + bias_o, __uint_as_float(r[4 * g + 3]) + bias_o};
            if constexpr (EPI == 0) {
              *reinterpret_cast<float4*>(yrow + t) = make_float4(v[0], v[1], v[2], v[3]);
              if (e.out_stats) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float y = (t + i < p.T) ? prelu_f(v[i], out_alpha) : 0.f;
                  s0 += y;
                  s1 = fmaf(y, y, s1);
                }
              }
              if (e.ch_stats) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float y = (t + i < p.T) ? v[i] : 0.f;
                  s2 += y;
                  s3 = fmaf(y, y, s3);
                }
              }
            } else if constexpr (EPI == 1) {
              *reinterpret_cast<float4*>(yrow + t) = make_float4(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
            } else if constexpr (EPI == 3) {   // decoder masks: Y2 = relu(v), Y = aux * relu(v)
              const float4 rr = gop[g];
              const float m0 = fmaxf(v[0], 0.f), m1 = fmaxf(v[1], 0.f), m2 = fmaxf(v[2], 0.f), m3 = fmaxf(v[3], 0.f);
              *reinterpret_cast<float4*>(e.Y2 + n * e.bsy2 + (int64_t)o * e.ldy2 + t) = make_float4(m0, m1, m2, m3);
              *reinterpret_cast<float4*>(yrow + t) = make_float4(rr.x * m0, rr.y * m1, rr.z * m2, rr.w * m3);
            } else if constexpr (EPI == 2) {
              const float4 rr = gop[g];
              *reinterpret_cast<float4*>(yrow + t) = make_float4(v[0] + rr.x, v[1] + rr.y, v[2] + rr.z, v[3] + rr.w);
            } else if constexpr (EPI == 10) {
              // dy2 = r2*(g2*v - mh - yhat2*mhy) = cA*v + cB*y2 + cC ;  dd = dy2 * prelu'(d)
              const float4 d4 = gop[g];
              const float draw[4] = {d4.x, d4.y, d4.z, d4.w};
              float dd[4];
              if (!edge_chunk) {   // interior columns: no masks, no edge bookkeeping (warp-uniform)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float dvi = draw[i];
                  const bool pos = dvi > 0.f;
                  const float y2 = pos ? dvi : a2 * dvi;
                  const float dy2 = fmaf(cA, v[i], fmaf(cB, y2, cC));
                  const float ddv = pos ? dy2 : a2 * dy2;
                  dd[i] = ddv;
                  s0 += ddv;                        // S
                  s1 = fmaf(ddv, dvi, s1);          // sum dd*d
                  s3 += pos ? 0.f : dy2 * dvi;      // dalpha2
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int tt_ = t + i;
                  const bool ok = tt_ < p.T;
                  const float dvi = ok ? draw[i] : 1.f;       // padding columns may hold NaN: neutralise the operands
                  const bool pos = dvi > 0.f;
                  const float y2 = pos ? dvi : a2 * dvi;
                  const float dy2 = ok ? fmaf(cA, v[i], fmaf(cB, y2, cC)) : 0.f;
                  const float ddv = pos ? dy2 : a2 * dy2;
                  dd[i] = ddv;
                  s0 += ddv;
                  s1 = fmaf(ddv, dvi, s1);
                  s3 += pos ? 0.f : dy2 * dvi;
                  sL += tt_ < e.dil ? ddv : 0.f;              // columns whose left tap falls off the sequence
                  sR += tt_ >= p.T - e.dil ? ddv : 0.f;       // ... right tap
                }
              }
              *reinterpret_cast<float4*>(yrow + t) = make_float4(dd[0], dd[1], dd[2], dd[3]);
            }
          }
        }
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(a));
      // per-row statistics: one (double) atomic per warp and quantity
      if constexpr (EPI == 0) {
        if (e.out_stats) {
          s0 = warp_sum(s0);
          s1 = warp_sum(s1);
          if (lane == 0) {
            atomicAdd(e.out_stats + 2 * n, (double)s0);
            atomicAdd(e.out_stats + 2 * n + 1, (double)s1);
          }
        }
        if (e.ch_stats && o_ok) {   // BatchNorm batch statistics: this thread owns channel o
          atomicAdd(e.ch_stats + 2 * o, (double)s2);
          atomicAdd(e.ch_stats + 2 * o + 1, (double)s3);
        }
      }
      if constexpr (EPI == 10) {
        // P1 = sum dd*g1*kappa, P2 = sum dd*(d - bd), P3 = sum dd*be1*kappa with kappa = w0+w1+w2 minus the taps
        // that fall outside [0,T) (interior columns all share kappa = w0+w1+w2)
        const float S = s0, SD = s1;
        const float kS = (w0 + w1 + w2) * S - w0 * sL - w2 * sR;
        s0 = gam1 * kS;
        s1 = SD - bdm * S;
        s2 = bet1 * kS;
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) {
          atomicAdd(e.rowacc + 8 * n + 2, (double)s0);
          atomicAdd(e.rowacc + 8 * n + 3, (double)s1);
          atomicAdd(e.rowacc + 8 * n + 4, (double)s2);
          atomicAdd(e.rowacc + 8 * n + 5, (double)s3);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_W_TMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), (const cuuint64_t*)dims,
                  (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  return 0;
}

bool gemm_wx_tc_eligible(const GemmWxP& p, int pro, int epi) {
  if ((p.M & 3) || (p.Kd & 3) || p.Kd > TC_MAXK) return false;   // partial tiles: TMA zero-fill + epilogue row mask
  if ((p.ldx & 31) || (reinterpret_cast<uintptr_t>(p.X) & 127) || (p.bsx & 3)) return false;   // 4-D map over 128-byte atoms
  if (!(pro == 0 || pro == 2 || pro == 3)) return false;
  if (!(epi == 0 || epi == 1 || epi == 2 || epi == 3 || epi == 10)) return false;
  if (epi == 3 && ((p.ep.ldy2 & 3) || !aligned16(p.ep.Y2) || (p.ep.bsy2 & 3) || (p.ep.ldr & 3) || !aligned16(p.ep.R) || (p.ep.bsr & 3))) return false;
  if ((p.ep.ldy & 3) || !aligned16(p.ep.Y) || (p.ep.bsy & 3)) return false;
  if (epi == 2 && ((p.ep.ldr & 3) || !aligned16(p.ep.R) || (p.ep.bsr & 3))) return false;
  if (epi == 10 && ((p.ep.ldd & 3) || !aligned16(p.ep.d) || (p.ep.bsd & 3))) return false;
  return true;
}

size_t gemm_wx_tc_ws_bytes(int M, int Kd) {   // hi + lo fp32 copies, or hi + the bf16 images (K padded to 16) of the mixed mode
  return (size_t)((M + 31) & ~31) * (Kd + ((Kd + 15) & ~15)) * sizeof(float);
}

template <int PRO, int EPI>
static int launch_tc_t(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx, const TcParams& P, cudaStream_t st) {
  auto k = gemm_wx_tc_kernel<PRO, EPI>;
  WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  int n_sm = 0;
  if (int rc = sm_count(&n_sm)) return rc;     // per current device (cached per device index)
  int grid = P.n_tiles < n_sm ? P.n_tiles : n_sm;
  k<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(mh, ml, mx, P);
  WB_LAUNCH_CHECK("gemm_wx_tc");
  return 0;
}

bool gemm_wx_tc2_eligible(const GemmWxP& p, int pro, int epi);
bool gemm_wx_tc2_instantiated(int pro, int epi);
__global__ void split_w_mixed_kernel(const float* __restrict__ W, int64_t ldw, int w_trans, int M, int Kd, float* __restrict__ hi,
                                     uint8_t* __restrict__ img);
int launch_gemm_wx_tc2(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx, TcParams P, int pro, int epi,
                       cudaStream_t st);

// W is [M][Kd] (ldw) or, a_trans, [Kd][M] (ldw). ws: >= gemm_wx_tc_ws_bytes(M, Kd), 16-byte aligned.
int launch_gemm_wx_tc(const GemmWxP& p, bool a_trans, int pro, int epi, void* ws, cudaStream_t st) {
  if (!gemm_wx_tc_eligible(p, pro, epi)) return fail(-2, "gemm_wx_tc: shape/config not eligible");
  if (!ws || !aligned16(ws)) return fail(-1, "gemm_wx_tc: workspace missing or misaligned");
  const int M32 = (p.M + 31) & ~31;
  float* whi = reinterpret_cast<float*>(ws);
  float* wlo = whi + (size_t)M32 * p.Kd;
  // the 2-CTA kernel runs the mixed tf32 + bf16 split product (its weight operand is prepared differently)
  const bool use2 = gemm_wx_tc2_eligible(p, pro, epi) && gemm_wx_tc2_instantiated(pro, epi);
  // (measured, profiles/r02_*: with an operand prologue (PRO >= 2: gLN-apply + PReLU, result stored back as the tf32 operand) the
  // transform warps of the 4-stage ring become the critical path and the mixed mode is 18 % slower than 3xTF32 there)
  const bool mixed = use2 && pro < 2 && !(g_tc_flags & 1024);
  if (!p.ws_presplit) {   // loops that reuse one weight split it once and set ws_presplit
    if (mixed) {
      const int total = (p.M >> 2) * (((p.Kd + 15) >> 4) * 16);
      split_w_mixed_kernel<<<cdiv(total, 256), 256, 0, st>>>(p.W, p.ldw, a_trans ? 1 : 0, p.M, p.Kd, whi,
                                                             reinterpret_cast<uint8_t*>(wlo));
    } else {
      const int total = M32 * p.Kd;
      split_w_kernel<<<cdiv(total, 256), 256, 0, st>>>(p.W, p.ldw, a_trans ? 1 : 0, p.M, p.Kd, whi, wlo);
    }
    WB_LAUNCH_CHECK("split_w");
  }
  CUtensorMap mh, ml, mx, mx2;
  {  // weights, pre-tiled [o/32][k][32]: dims (32, Kd, M/32), box (32, 16, 4) = the CTA's whole 128 x 16 tile
    uint64_t dims[3] = {32, (uint64_t)p.Kd, (uint64_t)(M32 / 32)};
    uint64_t strides[2] = {128, (uint64_t)p.Kd * 128};
    uint32_t box[3] = {32, TC_BK, TC_BM / 32};
    if (int rc = encode_map(&mh, whi, 3, dims, strides, box)) return rc;
    if (int rc = encode_map(&ml, wlo, 3, dims, strides, box)) return rc;
  }
  {  // activations [n][k][ld] viewed as (32 t, Kd, ld/32 atoms, n): atoms are 128-byte row segments (ld % 32 == 0),
     // so one box (32, 16, 8, 1) lands the 16 x 256 tile as 8 swizzle-atom columns; atoms beyond T are zero-filled
    uint64_t dims[4] = {32, (uint64_t)p.Kd, (uint64_t)cdiv(p.T, 32), (uint64_t)p.n};
    uint64_t strides[3] = {(uint64_t)p.ldx * 4, 128, (uint64_t)p.bsx * 4};
    uint32_t box[4] = {32, TC_BK, TC_BN / 32, 1};
    if (int rc = encode_map(&mx, p.X, 4, dims, strides, box)) return rc;
    uint32_t box2[4] = {32, TC_BK, TC_BN / 64, 1};
    if (int rc = encode_map(&mx2, p.X, 4, dims, strides, box2)) return rc;
  }
  TcParams P;
  P.g = p;
  P.n_ob = cdiv(p.M, TC_BM);
  P.n_tt = cdiv(p.T, TC_BN);
  P.n_tiles = P.n_ob * P.n_tt * p.n;
  P.skip_hi_store = (g_tc_flags & 1) ? 0 : 1;   // default: raw tile is the hi operand (HW truncates tf32 inputs; measured)
  // transform-warp groups: bits 2-3 of the flags (0 = default 2 groups, 1 -> 1, 2 -> 2, 3 -> 4).  The group count MUST
  // divide the stage count (a slot must always belong to the same group, or a group could run two mbarrier phases
  // ahead and pass a parity wait spuriously): 1-CTA ring = 4 stages, 2-CTA ring = 6 stages.
  {
    const int sel = (g_tc_flags >> 2) & 3;
    P.xf_groups = sel == 1 ? 1 : sel == 3 ? 4 : 2;
#ifdef WESEP_TC_DEBUG   // the wrong-result timing switches (bits 4-6) exist in debug builds only
    P.dbg = (g_tc_flags >> 4) & 15;
#else
    P.dbg = (g_tc_flags >> 4) & 8;    // release: only 'single store box' (correct results) is reachable
#endif
  }
  P.mixed = mixed ? 1 : 0;
  P.wimg = reinterpret_cast<const uint8_t*>(wlo);
  if (use2) {                                   // 2-CTA (cta_group::2) kernel for 256-channel multiples
    int rc = launch_gemm_wx_tc2(mh, ml, mx2, P, pro, epi, st);
    if (rc != -100) return rc;
  }
  if (pro == 0 && epi == 0) return launch_tc_t<0, 0>(mh, ml, mx, P, st);
  if (pro == 0 && epi == 2) return launch_tc_t<0, 2>(mh, ml, mx, P, st);
  if (pro == 0 && epi == 10) return launch_tc_t<0, 10>(mh, ml, mx, P, st);
  if (pro == 2 && epi == 2) return launch_tc_t<2, 2>(mh, ml, mx, P, st);
  if (pro == 2 && epi == 0) return launch_tc_t<2, 0>(mh, ml, mx, P, st);
  if (pro == 3 && epi == 0) return launch_tc_t<3, 0>(mh, ml, mx, P, st);
  if (pro == 0 && epi == 1) return launch_tc_t<0, 1>(mh, ml, mx, P, st);
  if (pro == 0 && epi == 3) return launch_tc_t<0, 3>(mh, ml, mx, P, st);
  return fail(-2, "gemm_wx_tc: unsupported (pro, epi)");
}

}  // namespace wb

extern "C" int wesep_b200_set_tc_flags(int flags) {
#ifndef WESEP_TC_DEBUG
  if (flags & (16 | 32 | 64)) return wb::fail(-2, "set_tc_flags: bits 4-6 (timing experiments with wrong results) need a -DWESEP_TC_DEBUG build");
#endif
  wb::g_tc_flags = flags;
  return 0;
}
extern "C" int64_t wesep_b200_gemm_ws_bytes(int M, int Kd) { return (int64_t)wb::gemm_wx_tc_ws_bytes(M, Kd); }

// ================================================================================================
// Weight-gradient GEMM on tcgen05:  C[o][c] += sum_t A[n][o][t] * f(B[n][c][t])   (contraction over time)
//   D[M = 128 rows of A][N = 256 rows of B] in TMEM; both operands K-major (time contiguous): rows of
//   16 floats (64 B) with the 64-byte TMA/UMMA swizzle, 8-row atoms of 512 B.  Both operands are
//   activations, so both are split hi/lo by the transform warps (f = identity or PReLU on B).
//   One tile (and one K range) per CTA; partial results are added to C with fp32 atomics.
// ================================================================================================
namespace wb {

constexpr int DW_BM = 128, DW_BN = 256, DW_BK = 16, DW_STAGES = 4, DW_THREADS = 448;
constexpr int DW_A_BYTES = DW_BM * DW_BK * 4;   // 8192
constexpr int DW_B_BYTES = DW_BN * DW_BK * 4;   // 16384
constexpr int DW_OFF_AHI = 0, DW_OFF_ALO = DW_A_BYTES, DW_OFF_BHI = 2 * DW_A_BYTES, DW_OFF_BLO = 2 * DW_A_BYTES + DW_B_BYTES;
constexpr int DW_STAGE_BYTES = 2 * DW_A_BYTES + 2 * DW_B_BYTES;  // 49152
constexpr int DW_TX_BYTES = DW_A_BYTES + DW_B_BYTES;
constexpr int DW_SMEM_BYTES = DW_STAGES * DW_STAGE_BYTES + 256 + 1024;

// K-major operand, SWIZZLE_64B: rows of 64 B, 8-row atoms (512 B) -> SBO = 512 B; LBO unused (1).
__device__ __forceinline__ uint64_t make_desc_k_sw64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D=f32, A=B=tf32, both K-major, N=256, M=128
constexpr uint32_t DW_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(DW_BN >> 3) << 17) | ((uint32_t)(DW_BM >> 4) << 24);

struct DwTcParams {
  GemmDwP g;
  int n_ob, n_cb, n_tiles;
  int skip_hi_store;
  int mixed;           // 2-CTA kernel: tf32 leading term + bf16 cross terms instead of 3xTF32
};

// Stream-K over (tile, k-block) units: the grid is one CTA per SM and CTA i owns the contiguous unit range
// [i*U/G, (i+1)*U/G) of the flattened (tile-major) space, so all SMs carry the same number of k-blocks no matter
// how the tile count divides the SM count (128 tiles on 148 SMs left 14 % of the machine idle).  A range may span
// tiles; every (tile, k-range) segment ends with an atomic add of its partial D into C, which the accumulate-into-C
// contract already required.  The two TMEM accumulators let a segment's epilogue overlap the next segment's MMAs.
template <int PRO_B, bool RS>
__global__ void __launch_bounds__(DW_THREADS, 1)
    gemm_dw_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const DwTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const GemmDwP& p = P.g;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar0 = base + DW_STAGES * DW_STAGE_BYTES;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (DW_STAGES + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * DW_STAGES + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (3 * DW_STAGES + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (3 * DW_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + DW_STAGES * DW_STAGE_BYTES + 8 * (3 * DW_STAGES + 4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KBT = (p.T + DW_BK - 1) / DW_BK;
  const int64_t units = (int64_t)P.n_tiles * KBT;
  const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
  // unit -> (tile, k-block); tile = (row * n_cb + cb) * n_ob + ob
  auto decode = [&](int tile, int& ob, int& cb, int& row) {
    ob = tile % P.n_ob;
    const int rest = tile / P.n_ob;
    cb = rest % P.n_cb;
    row = rest / P.n_cb;
  };

  if (tid == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 4);   // 2 transform groups of 4 warps; group g owns stages with it % 2 == g
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);
      mbar_init(bar_acce(a), 4);    // 4 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = u0; u < u1;) {
        const int tile = (int)(u / KBT), kb_lo = (int)(u % KBT);
        const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
        int ob, cb, row;
        decode(tile, ob, cb, row);
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % DW_STAGES;
          const uint32_t ph = (it / DW_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t sb = base + s * DW_STAGE_BYTES;
          mbar_expect_tx(bar_full(s), DW_TX_BYTES);
          tma_load_3d(sb + DW_OFF_AHI, &map_a, bar_full(s), kb * DW_BK, ob * DW_BM, row);
          tma_load_3d(sb + DW_OFF_BHI, &map_b, bar_full(s), kb * DW_BK, cb * DW_BN, row);
        }
        u += kb_hi - kb_lo;
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t it = 0, sg = 0;
      for (int64_t u = u0; u < u1; ++sg) {
        const int kb_lo = (int)(u % KBT);
        const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
        const int a = sg & 1;
        mbar_wait(bar_acce(a), ((sg >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * DW_BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % DW_STAGES;
          const uint32_t ph = (it / DW_STAGES) & 1;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint32_t sb = base + s * DW_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < DW_BK / 8; ++ks) {
            const uint64_t a_hi = make_desc_k_sw64(sb + DW_OFF_AHI + ks * 32);
            const uint64_t a_lo = make_desc_k_sw64(sb + DW_OFF_ALO + ks * 32);
            const uint64_t b_hi = make_desc_k_sw64(sb + DW_OFF_BHI + ks * 32);
            const uint64_t b_lo = make_desc_k_sw64(sb + DW_OFF_BLO + ks * 32);
            tc_mma_tf32(d_tmem, a_lo, b_hi, DW_IDESC, (kb != kb_lo || ks != 0) ? 1u : 0u);
            tc_mma_tf32(d_tmem, a_hi, b_lo, DW_IDESC, 1u);
            tc_mma_tf32(d_tmem, a_hi, b_hi, DW_IDESC, 1u);
          }
          tc_commit(bar_empty(s));
        }
        tc_commit(bar_accf(a));
        u += kb_hi - kb_lo;
      }
    }
  } else if (warp >= 6) {
    const int tt_id = tid - 6 * 32;
    float alpha = 1.f;
    if constexpr (PRO_B >= 1) alpha = p.xb.alpha ? __ldg(p.xb.alpha) : 1.f;
    const int xf_gid = tt_id >> 7, xf_tid = tt_id & 127;   // DW_STAGES % 2 == 0: a slot always belongs to the same group
    const int n_it = (int)(u1 - u0);
    // optional by-product: row sums of A over time.  A float4 at index idx of the (swizzled) A tile always belongs
    // to tile row idx / 4, and this thread meets the same 4 rows in every stage, so the partial sums live in
    // registers and are flushed when the CTA's range moves to another tile.  Only column-block 0 tiles contribute
    // (the other column blocks see the same A rows again); frames >= T were zero-filled by the TMA.
    constexpr int A_F4 = DW_A_BYTES / 16 / 128;       // A float4s per transform thread per stage (4)
    constexpr int B_F4 = DW_B_BYTES / 16 / 128;       // B float4s per transform thread per stage (8)
    float rs[A_F4];
#pragma unroll
    for (int i = 0; i < A_F4; ++i) rs[i] = 0.f;
    auto flush_rs = [&](int tile) {
      int ob, cb, row;
      decode(tile, ob, cb, row);
      if (RS && cb == 0) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
          const int o = ob * DW_BM + (xf_tid + 128 * i) / 4;
          if (o < p.M) atomicAdd(p.a_rowsum + (int64_t)row * p.M + o, rs[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < A_F4; ++i) rs[i] = 0.f;
    };
    // PRO_B 2 / 3: per-channel (= per B-tile row) scale / shift of this thread's 8 rows, rebuilt per tile
    float bsc[B_F4], bsh[B_F4];
#pragma unroll
    for (int j = 0; j < B_F4; ++j) { bsc[j] = 1.f; bsh[j] = 0.f; }
    auto load_scsh = [&](int tile) {
      if constexpr (PRO_B >= 2) {
        int ob, cb, row;
        decode(tile, ob, cb, row);
        float mu = 0.f, r = 1.f;
        if (p.xb.row_stats) gln_mean_rstd(p.xb.row_stats + 2 * row, p.xb.count, p.xb.eps, mu, r);
#pragma unroll
        for (int j = 0; j < B_F4; ++j) {
          const int c = cb * DW_BN + (xf_tid + 128 * j) / 4;
          const float gm = (p.xb.ch_scale && c < p.N) ? __ldg(p.xb.ch_scale + c) : 1.f;
          const float bt = (p.xb.ch_shift && c < p.N) ? __ldg(p.xb.ch_shift + c) : 0.f;
          bsc[j] = gm * r;
          bsh[j] = bt - gm * mu * r;
        }
      }
    };
    int cur_tile = (int)(u0 / KBT);
    int next_b = KBT - (int)(u0 % KBT);               // first `it` that belongs to the next tile
    load_scsh(cur_tile);
    for (int it = xf_gid; it < n_it; it += 2) {
      const int s = it % DW_STAGES;
      const uint32_t ph = (it / DW_STAGES) & 1;
      while (it >= next_b) {
        if constexpr (RS) flush_rs(cur_tile);
        ++cur_tile;
        next_b += KBT;
        load_scsh(cur_tile);
      }
      mbar_wait(bar_full(s), ph);
      uint8_t* st = gbase + s * DW_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < A_F4 + B_F4; ++i) {
        const int idx = xf_tid + 128 * i;             // float4 index over [A tile | B tile]
        const bool is_b = i >= A_F4;
        const int off = is_b ? (idx * 16 - DW_A_BYTES) : idx * 16;
        uint8_t* hi_p = st + (is_b ? DW_OFF_BHI : DW_OFF_AHI) + off;
        uint8_t* lo_p = st + (is_b ? DW_OFF_BLO : DW_OFF_ALO) + off;
        float4 v = *reinterpret_cast<const float4*>(hi_p);
        if constexpr (RS) {
          if (!is_b) rs[is_b ? 0 : i] += (v.x + v.y) + (v.z + v.w);
        }
        if constexpr (PRO_B == 1) {
          if (is_b) { v.x = prelu_f(v.x, alpha); v.y = prelu_f(v.y, alpha); v.z = prelu_f(v.z, alpha); v.w = prelu_f(v.w, alpha); }
        }
        if constexpr (PRO_B == 2) {
          if (is_b) {
            const float c_ = bsc[is_b ? i - A_F4 : 0], d_ = bsh[is_b ? i - A_F4 : 0];
            v.x = fmaf(c_, prelu_f(v.x, alpha), d_); v.y = fmaf(c_, prelu_f(v.y, alpha), d_);
            v.z = fmaf(c_, prelu_f(v.z, alpha), d_); v.w = fmaf(c_, prelu_f(v.w, alpha), d_);
          }
        }
        if constexpr (PRO_B == 3) {
          if (is_b) {
            const float c_ = bsc[is_b ? i - A_F4 : 0], d_ = bsh[is_b ? i - A_F4 : 0];
            v.x = prelu_f(fmaf(c_, v.x, d_), alpha); v.y = prelu_f(fmaf(c_, v.y, d_), alpha);
            v.z = prelu_f(fmaf(c_, v.z, d_), alpha); v.w = prelu_f(fmaf(c_, v.w, d_), alpha);
          }
        }
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        if (!P.skip_hi_store || (PRO_B >= 1 && is_b)) *reinterpret_cast<float4*>(hi_p) = h;
        *reinterpret_cast<float4*>(lo_p) = l;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready(s));
    }
    if constexpr (RS) {
      if (n_it > 0) flush_rs(cur_tile);
    }
  } else {
    // epilogue warps 0..3: C[o][c] += D for every segment of this CTA's range
    const int q = warp;
    uint32_t sg = 0;
    for (int64_t u = u0; u < u1; ++sg) {
      const int tile = (int)(u / KBT), kb_lo = (int)(u % KBT);
      const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
      u += kb_hi - kb_lo;
      int ob, cb, row;
      decode(tile, ob, cb, row);
      const int a = sg & 1;
      const int o = ob * DW_BM + q * 32 + lane;
      const bool o_ok = o < p.M;                      // rows beyond M were zero-filled by the TMA
      float* C = p.C + (p.per_row ? (int64_t)row * p.M * p.ldc : 0) + (int64_t)(o_ok ? o : 0) * p.ldc + cb * DW_BN;
      const int ncols = min(DW_BN, p.N - cb * DW_BN);
      mbar_wait(bar_accf(a), (sg >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < DW_BN; c0 += 32) {
        if (c0 >= ncols) break;                       // warp-uniform
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * DW_BN + c0), r);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (o_ok && c0 + i < ncols) atomicAdd(C + c0 + i, __uint_as_float(r[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(a));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int encode_map_sw(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), (const cuuint64_t*)dims,
                  (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  return 0;
}

bool gemm_dw_tc_eligible(const GemmDwP& p, int pro_b) {
  if (pro_b < 0 || pro_b > 3) return false;
  if ((p.lda & 3) || (p.ldb & 3) || (p.bsa & 3) || (p.bsb & 3) || !aligned16(p.A) || !aligned16(p.B)) return false;
  return true;
}

bool gemm_dw_tc2_eligible(const GemmDwP& p, int pro_b);
int launch_gemm_dw_tc2(const GemmDwP& p, int pro_b, cudaStream_t st);

int launch_gemm_dw_tc(const GemmDwP& p, int pro_b, cudaStream_t st) {
  if (!gemm_dw_tc_eligible(p, pro_b)) return fail(-2, "gemm_dw_tc: not eligible");
  if (gemm_dw_tc2_eligible(p, pro_b)) return launch_gemm_dw_tc2(p, pro_b, st);   // 2-CTA kernel for 256-row multiples
  CUtensorMap ma, mb;
  {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)p.M, (uint64_t)p.n};
    uint64_t strides[2] = {(uint64_t)p.lda * 4, (uint64_t)p.bsa * 4};
    uint32_t box[3] = {DW_BK, DW_BM, 1};
    if (int rc = encode_map_sw(&ma, p.A, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)p.N, (uint64_t)p.n};
    uint64_t strides[2] = {(uint64_t)p.ldb * 4, (uint64_t)p.bsb * 4};
    uint32_t box[3] = {DW_BK, DW_BN, 1};
    if (int rc = encode_map_sw(&mb, p.B, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  DwTcParams P;
  P.g = p;
  P.n_ob = cdiv(p.M, DW_BM);
  P.n_cb = cdiv(p.N, DW_BN);
  P.n_tiles = P.n_ob * P.n_cb * p.n;
  P.skip_hi_store = (g_tc_flags & 1) ? 0 : 1;
  P.mixed = 0;
  int n_sm = 0;
  if (int rc = sm_count(&n_sm)) return rc;     // per current device (cached per device index)
  // stream-K grid: one CTA per SM, but never fewer than 8 k-blocks per CTA
  const int64_t units = (int64_t)P.n_tiles * cdiv(p.T, DW_BK);
  int grid = n_sm;
  if (units < (int64_t)grid * 8) grid = (int)(units / 8 > 0 ? units / 8 : 1);
  // measured (tools/ab_block.py): with 128 tiles on 148 SMs, one whole tile per CTA is 6 % FASTER than the balanced
  // 148-CTA split (the machine is power-capped; fewer partial-tile epilogues and better L2 sharing of the operand
  // tiles win), so stream-K is kept for the cases that would leave most of the SMs idle.  bit 8 of the flags forces it.
  if (!(g_tc_flags & 256) && P.n_tiles <= n_sm && P.n_tiles * 4 >= n_sm * 3) grid = P.n_tiles;
  auto launch = [&](auto k) -> int {
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));
    k<<<grid, DW_THREADS, DW_SMEM_BYTES, st>>>(ma, mb, P);
    return 0;
  };
  int rc;
  if (p.a_rowsum) {   // row sums of A are only wired for the two prologues the TCN block uses
    if (pro_b > 1) return fail(-2, "gemm_dw_tc: a_rowsum needs pro_b 0 or 1");
    rc = pro_b == 0 ? launch(gemm_dw_tc_kernel<0, true>) : launch(gemm_dw_tc_kernel<1, true>);
  } else {
    rc = pro_b == 0 ? launch(gemm_dw_tc_kernel<0, false>) : pro_b == 1 ? launch(gemm_dw_tc_kernel<1, false>)
         : pro_b == 2 ? launch(gemm_dw_tc_kernel<2, false>) : launch(gemm_dw_tc_kernel<3, false>);
  }
  if (rc) return rc;
  WB_LAUNCH_CHECK("gemm_dw_tc");
  return 0;
}

}  // namespace wb

// ================================================================================================
// 2-CTA variant of gemm_wx_tc (tcgen05 cta_group::2): a cluster of two CTAs (one TPC) computes a
// 256-channel x 256-frame tile.  Each CTA stages ITS 128 weight rows and ITS 128-frame half of the
// activation tile, so per-SM shared-memory traffic (MMA operand reads + TMA fills + hi/lo transform) drops from
// ~177 B/clk to ~116 B/clk — below the 128 B/clk/SM limit that bounds the 1-CTA kernel — and the smaller
// stages (32 KB) allow a 6-deep TMA ring.  The leader CTA's MMA thread issues tcgen05.mma.cta_group::2
// (M = 256) reading both CTAs' smem; tcgen05.commit multicasts to both CTAs' barriers; the peer's transform /
// epilogue warps signal the leader's barriers with remote mbarrier arrives.
// ================================================================================================
namespace wb {

// Ring depth NS and epilogue staging (2 KB boxes of 32 channels x 16 frames, 64-byte swizzle, per epilogue warp):
//   PRO 0, EPI 0 : NS = 6, two store boxes per warp (no gLN scale/shift tables)               (192 + 32 KB)
//   otherwise    : NS = 4, two store boxes + (EPI 2, 10) T2_NL load boxes (R / d prefetch)    (128 + 32..80 KB)
// The epilogue moves its global traffic with TMA (coalesced 64-byte row segments) instead of one 16-byte access per
// lane per row, which was measured to cost more than the MMAs themselves (profiles/r01_*: 448 vs 183 us).
constexpr int T2_NL = 4, T2_NSB = 2;   // T2_NL = max load boxes per warp (barrier slots); t2_nl(epi) are used
constexpr int T2_EBOX = 2048, T2_ECOLS = 16;
__host__ __device__ constexpr int t2_nl(int epi) { return epi == 10 ? 4 : 3; }   // EPI 10 (M = 512, K = 256) has the most epilogue per MMA
__host__ __device__ constexpr int t2_nbuf(int epi) { return (epi == 2 || epi == 3 || epi == 10) ? T2_NSB + t2_nl(epi) : T2_NSB; }
__host__ __device__ constexpr int t2_stages(int pro, int epi) { return (pro == 0 && epi == 0) ? 6 : 4; }
__host__ __device__ constexpr int t2_scsh_bytes(int pro) { return pro >= 2 ? 2 * TC_MAXK * 4 : 0; }
constexpr int T2_BAR_BYTES = 512;
constexpr int T2_XH_BYTES = (TC_BN / 2 / 32) * TC_BOX_BYTES;  // this CTA's half of the X tile: 4 boxes = 8192
constexpr int T2_OFF_WHI = 0, T2_OFF_WLO = TC_W_BYTES, T2_OFF_XHI = 2 * TC_W_BYTES, T2_OFF_XLO = 2 * TC_W_BYTES + T2_XH_BYTES;
constexpr int T2_STAGE_BYTES = 2 * TC_W_BYTES + 2 * T2_XH_BYTES;  // 32768
constexpr int T2_TX_BYTES = 2 * TC_W_BYTES + T2_XH_BYTES;          // per CTA per stage
__host__ __device__ constexpr int t2_smem_bytes(int pro, int epi) {
  return t2_stages(pro, epi) * T2_STAGE_BYTES + 8 * t2_nbuf(epi) * T2_EBOX + t2_scsh_bytes(pro) + T2_BAR_BYTES + 1024;
}
// D=f32, A=B=tf32, both MN-major, N=256, M=256 (cta_group::2)
constexpr uint32_t T2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- mixed-precision split product (2-CTA kernels): x.y = xt.yt + xl.yt + xt.yl with xt = tf32(x), xl = x - xt.
// The leading term runs as kind::tf32 on the raw fp32 tiles (the tensor core truncates its inputs to tf32); the two
// cross terms are 2^-11 of the product, so they run as kind::f16 on bf16 operands (bf16(x) for the big factor, bf16(xl)
// for the small one: error 2^-9 . 2^-11 = 2^-20 relative, the same order as the dropped xl.yl term) at twice the tf32
// rate and half the operand bytes.  One K = 16 slice costs 2 tf32 + 2 bf16 instructions = 4 time units instead of the 6 of
// 3xTF32, at the same fp32-grade accuracy (measured rel. L2 ~1e-6) and the same shared-memory footprint:
//   3xTF32 stage: [hi fp32][lo fp32]        mixed stage: [raw fp32][bf16(x) | bf16(x - tf32(x))]
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 hi / lo of four fp32 values -> two 8-byte packets
__device__ __forceinline__ void split_bf16_quad(const float4& v, uint2& bh, uint2& bl) {
  const float lx = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), ly = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  const float lz = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), lw = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  uint32_t h0, h1, l0, l1;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(h0) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(h1) : "f"(v.w), "f"(v.z));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(l0) : "f"(ly), "f"(lx));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(l1) : "f"(lw), "f"(lz));
  bh = make_uint2(h0, h1);
  bl = make_uint2(l0, l1);
}
// MN-major 16-bit operand, 128-byte swizzle: [MN atom of 64 elements][K atom of 8 rows][8 rows][128 B]; LBO = stride between
// MN atoms (2048 B), SBO = stride between the two K atoms of a K = 16 instruction (1024 B)
__device__ __forceinline__ uint64_t make_desc_mn16_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(2048 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of 4 consecutive MN elements (mn % 4 == 0) at K row k of a [128 MN x 16 K] bf16 tile in that layout
__device__ __host__ __forceinline__ uint32_t mn16_off(uint32_t mn, uint32_t k) {
  const uint32_t row = k & 7u;
  return (mn >> 6) * 2048u + (k >> 3) * 1024u + row * 128u + ((((mn & 63u) >> 3) ^ row) << 4) + (mn & 4u) * 2u;
}
// K-major 16-bit operand with rows of 16 elements (32 B), 32-byte swizzle: 8-row atoms of 256 B (SBO), 16-byte chunk ^= (row >> 2) & 1
__device__ __forceinline__ uint64_t make_desc_k16_sw32(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
// D=f32, A=B=bf16, N=256, M=256 (cta_group::2): both MN-major (wx) / both K-major (dw)
constexpr uint32_t T2_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) |
                                   ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t D2_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr int TC_FLAG_NO_MIXED = 1024;   // set_tc_flags bit 10: 2-CTA kernels fall back to 3xTF32 (A/B timing, tests)

// prep for the mixed mode: hi = tf32-truncated W pre-tiled as split_w_kernel does, img = per (128-channel block, 16-row K
// block) an 8 KB shared-memory image [bf16(W) tile | bf16(W - tf32(W)) tile] in the MN-major layout above (ONE bulk copy
// per stage); K rows >= Kd are zero
__global__ void split_w_mixed_kernel(const float* __restrict__ W, int64_t ldw, int w_trans, int M, int Kd, float* __restrict__ hi,
                                     uint8_t* __restrict__ img) {
  const int KB = (Kd + 15) >> 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // over (o / 4, k padded)
  const int total = (M >> 2) * KB * 16;
  if (idx >= total) return;
  const int o4 = idx % (M >> 2), k = idx / (M >> 2);
  const int o = o4 * 4;
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < Kd) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = w_trans ? W[(int64_t)k * ldw + o + i] : W[(int64_t)(o + i) * ldw + k];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                               // the fp32 (tf32 leading term) copy, [o / 32][k][32] tiling
      const int oo = o + i;
      hi[((int64_t)(oo >> 5) * Kd + k) * 32 + (oo & 31)] = __uint_as_float(__float_as_uint(w[i]) & 0xFFFFE000u);
    }
  }
  uint2 bh, bl;
  split_bf16_quad(make_float4(w[0], w[1], w[2], w[3]), bh, bl);
  uint8_t* tile = img + ((int64_t)(o >> 7) * KB + (k >> 4)) * 8192;
  const uint32_t off = mn16_off((uint32_t)(o & 127), (uint32_t)(k & 15));
  *reinterpret_cast<uint2*>(tile + off) = bh;
  *reinterpret_cast<uint2*>(tile + 4096 + off) = bl;
}

template <int PRO, int EPI, int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    gemm_wx_tc2_kernel(const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                       const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_y,
                       const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_y2,
                       const TcParams P) {
  constexpr int T2_STAGES = NS;
  constexpr int NBUF = t2_nbuf(EPI);
  constexpr int EB_OFF = NS * T2_STAGE_BYTES;                 // epilogue staging boxes
  constexpr int AUX_OFF = EB_OFF + 8 * NBUF * T2_EBOX;        // sc / sh, then the barriers
  extern __shared__ uint8_t smem_raw[];
  const GemmWxP& p = P.g;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  constexpr int SCSH = t2_scsh_bytes(PRO);
  float* sc = reinterpret_cast<float*>(gbase + AUX_OFF);   // only PRO >= 2 has (and touches) the tables
  float* sh = sc + TC_MAXK;
  const uint32_t bar0 = base + AUX_OFF + SCSH;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (T2_STAGES + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * T2_STAGES + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (3 * T2_STAGES + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (3 * T2_STAGES + 2 + a); };
  auto bar_ld = [&](int w, int b_) { return bar0 + 8u * (3 * T2_STAGES + 4 + w * T2_NL + b_); };   // per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + AUX_OFF + SCSH + 8 * (3 * T2_STAGES + 4 + 8 * T2_NL));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int KB = (p.Kd + TC_BK - 1) / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(bar_full(s), 1);     // local TMA transaction barrier
      mbar_init(bar_ready(s), 16 / P.xf_groups);   // (leader's copy is used) transform warps of the owning group x 2 CTAs
      mbar_init(bar_empty(s), 1);    // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);     // multicast tcgen05.commit
      mbar_init(bar_acce(a), 16);    // (leader's copy is used) 8 epilogue warps x 2 CTAs
    }
    for (int w = 0; w < 8; ++w)
      for (int b_ = 0; b_ < T2_NL; ++b_) mbar_init(bar_ld(w, b_), 1);   // epilogue R / d prefetch boxes
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == TC_W_TMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile decode shared by all roles: tile -> (channel pair, frame tile, row)
  auto decode = [&](int tile, int& o0, int& t0, int& n) {
    const int ob2 = tile % P.n_ob, rest = tile / P.n_ob;
    t0 = (rest % P.n_tt) * TC_BN;
    n = rest / P.n_tt;
    o0 = ob2 * 256 + (int)rank * TC_BM;
  };

  if (warp == TC_W_TMA) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < P.n_tiles; tile += n_clusters) {
        int o0, t0, n;
        decode(tile, o0, t0, n);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % T2_STAGES;
          const uint32_t ph = (it / T2_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t sb = base + s * T2_STAGE_BYTES;
          mbar_expect_tx(bar_full(s), T2_TX_BYTES);
          const int k0 = kb * TC_BK;
          tma_load_3d(sb + T2_OFF_WHI, &map_whi, bar_full(s), 0, k0, o0 >> 5);
          if (P.mixed) {      // [bf16(W) | bf16(W - tf32(W))] image of this (channel block, K block): one 8 KB bulk copy
            const uint8_t* src = P.wimg + ((int64_t)(o0 >> 7) * KB + kb) * 8192;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sb + T2_OFF_WLO),
                         "l"(src), "r"(8192u), "r"(bar_full(s))
                         : "memory");
          } else {
            tma_load_3d(sb + T2_OFF_WLO, &map_wlo, bar_full(s), 0, k0, o0 >> 5);
          }
          tma_load_4d(sb + T2_OFF_XHI, &map_x2, bar_full(s), 0, k0, (t0 >> 5) + 4 * (int)rank, n);
        }
      }
    }
  } else if (warp == TC_W_MMA) {
    if (lane == 0 && rank == 0) {   // leader CTA issues for the pair
      uint32_t it = 0, ti = 0;
      for (int tile = cluster_id; tile < P.n_tiles; tile += n_clusters, ++ti) {
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1;
        mbar_wait(bar_acce(a), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * TC_BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % T2_STAGES;
          const uint32_t ph = (it / T2_STAGES) & 1;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint32_t sb = base + s * T2_STAGE_BYTES;
          if (P.mixed) {
            tc_mma_bf16_2cta(d_tmem, make_desc_mn16_sw128(sb + T2_OFF_WLO + 4096), make_desc_mn16_sw128(sb + T2_OFF_XLO), T2_IDESC_BF16,
                             kb != 0 ? 1u : 0u);                                      // (W - tf32 W) . X
            tc_mma_bf16_2cta(d_tmem, make_desc_mn16_sw128(sb + T2_OFF_WLO), make_desc_mn16_sw128(sb + T2_OFF_XLO + 4096), T2_IDESC_BF16,
                             1u);                                                     // W . (X - tf32 X)
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks)                                    // tf32 W . tf32 X
              tc_mma_tf32_2cta(d_tmem, make_desc_mn_sw128(sb + T2_OFF_WHI + ks * 1024), make_desc_mn_sw128(sb + T2_OFF_XHI + ks * 1024),
                               T2_IDESC, 1u);
          } else {
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
              const uint64_t a_hi = make_desc_mn_sw128(sb + T2_OFF_WHI + ks * 1024);
              const uint64_t a_lo = make_desc_mn_sw128(sb + T2_OFF_WLO + ks * 1024);
              const uint64_t b_hi = make_desc_mn_sw128(sb + T2_OFF_XHI + ks * 1024);
              const uint64_t b_lo = make_desc_mn_sw128(sb + T2_OFF_XLO + ks * 1024);
              tc_mma_tf32_2cta(d_tmem, a_lo, b_hi, T2_IDESC, (kb | ks) != 0 ? 1u : 0u);
              tc_mma_tf32_2cta(d_tmem, a_hi, b_lo, T2_IDESC, 1u);
              tc_mma_tf32_2cta(d_tmem, a_hi, b_hi, T2_IDESC, 1u);
            }
          }
          tc_commit_mc2(bar_empty(s));   // frees stage s in BOTH CTAs
        }
        tc_commit_mc2(bar_accf(a));      // accumulators complete in BOTH CTAs
      }
    }
  } else if (warp >= TC_W_XF) {
    const int tt_id = tid - TC_W_XF * 32;
    const int xf_groups = P.xf_groups, xf_nthr = 256 / xf_groups;
    const int xf_gid = tt_id / xf_nthr, xf_tid = tt_id % xf_nthr;
    float alpha = 1.f;
    if constexpr (PRO >= 1) alpha = p.xf.alpha ? __ldg(p.xf.alpha) : 1.f;
    uint32_t it = 0;
    int cur_n = -1;
    for (int tile = cluster_id; tile < P.n_tiles; tile += n_clusters) {
      int o0, t0, n;
      decode(tile, o0, t0, n);
      if constexpr (PRO >= 2) {
        if (n != cur_n) {
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          float mu = 0.f, r = 1.f;
          if (p.xf.row_stats) gln_mean_rstd(p.xf.row_stats + 2 * n, p.xf.count, p.xf.eps, mu, r);
          for (int k = tt_id; k < p.Kd; k += 256) {
            const float gm = p.xf.ch_scale ? __ldg(p.xf.ch_scale + k) : 1.f;
            const float bt = p.xf.ch_shift ? __ldg(p.xf.ch_shift + k) : 0.f;
            sc[k] = gm * r;
            sh[k] = bt - gm * mu * r;
          }
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          cur_n = n;
        }
      }
      for (int kb = 0; kb < KB; ++kb, ++it) {
        if ((int)(it % (uint32_t)xf_groups) != xf_gid) continue;
        const int s = it % T2_STAGES;
        const uint32_t ph = (it / T2_STAGES) & 1;
        mbar_wait(bar_full(s), ph);
        uint8_t* xs_hi = gbase + s * T2_STAGE_BYTES + T2_OFF_XHI;
        uint8_t* xs_lo = gbase + s * T2_STAGE_BYTES + T2_OFF_XLO;
        for (int i = 0; i < ((P.dbg & 4) ? 0 : T2_XH_BYTES / 16 / xf_nthr); ++i) {
          const int off = (xf_tid + xf_nthr * i) * 16;
          float4 v = *reinterpret_cast<const float4*>(xs_hi + off);
          if constexpr (PRO >= 1) {
            float c = 1.f, d = 0.f;
            if constexpr (PRO >= 2) {
              const int kk = kb * TC_BK + ((off % TC_BOX_BYTES) >> 7);
              const bool kok = kk < p.Kd;
              c = kok ? sc[kk] : 0.f;
              d = kok ? sh[kk] : 0.f;
            }
            if constexpr (PRO == 3) {
              v.x = prelu_f(fmaf(c, v.x, d), alpha); v.y = prelu_f(fmaf(c, v.y, d), alpha);
              v.z = prelu_f(fmaf(c, v.z, d), alpha); v.w = prelu_f(fmaf(c, v.w, d), alpha);
            } else {
              v.x = fmaf(c, prelu_f(v.x, alpha), d); v.y = fmaf(c, prelu_f(v.y, alpha), d);
              v.z = fmaf(c, prelu_f(v.z, alpha), d); v.w = fmaf(c, prelu_f(v.w, alpha), d);
            }
          }
          if (P.mixed) {
            // raw tile: 4 boxes (32 frames each) x [16 K rows x 128 B], 32-byte granules XORed with (k & 3): recover the
            // logical (k, t) of this float4 and drop its bf16 hi / lo packets into the MN-major 16-bit tiles
            const uint32_t k = ((uint32_t)off >> 7) & 15u, c16 = ((uint32_t)off >> 4) & 7u;
            const uint32_t tt = ((uint32_t)off >> 11) * 32u + ((((c16 >> 1) ^ (k & 3u)) << 3) | ((c16 & 1u) << 2));
            uint2 bh, bl;
            split_bf16_quad(v, bh, bl);
            const uint32_t o16 = mn16_off(tt, k);
            if (PRO != 0) *reinterpret_cast<float4*>(xs_hi + off) = v;      // the tensor core truncates to tf32 itself
            *reinterpret_cast<uint2*>(xs_lo + o16) = bh;
            *reinterpret_cast<uint2*>(xs_lo + 4096 + o16) = bl;
          } else {
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            if (PRO != 0 || !P.skip_hi_store) *reinterpret_cast<float4*>(xs_hi + off) = h;
            *reinterpret_cast<float4*>(xs_lo + off) = l;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {   // signal the LEADER's ready barrier (it gates the pair's MMAs)
          if (rank == 0) mbar_arrive(bar_ready(s));
          else mbar_arrive_cluster(bar_ready(s), 0);
        }
      }
    }
  } else {
    // epilogue warps 0..7: TMEM -> registers (lane = channel, 16 frames per pass) -> fused epilogue -> swizzled 2 KB
    // box in shared memory -> TMA store.  The R / d operand of EPI 2 / 10 arrives the same way (TMA load boxes,
    // prefetched T2_NL passes ahead, across tile boundaries), so every global access of the epilogue is a full
    // 64-byte row segment issued by the TMA unit rather than 32 scattered 16-byte accesses per warp instruction.
    constexpr bool HAS_LD = (EPI == 2 || EPI == 3 || EPI == 10);
    constexpr int NLB = t2_nl(EPI);   // load boxes in flight per warp
    const int q = warp & 3;
    const int chalf = warp >> 2;
    const EpiP& e = p.ep;
    uint8_t* sbox = gbase + EB_OFF + warp * (NBUF * T2_EBOX);          // T2_NSB store boxes, then the load boxes
    const uint32_t sbox_u = base + EB_OFF + warp * (NBUF * T2_EBOX);
    uint32_t sc_i = 0;                                                 // store passes issued by this warp
    const uint32_t lrow = (uint32_t)lane * 64u, lsw = (uint32_t)(lane >> 1) & 3u;   // 64-byte swizzle: chunk ^= (row / 2) % 4
    const bool ld_on = HAS_LD && !(P.dbg & 1);
    // prefetch iterator over this warp's (tile, pass) sequence; passes whose first frame is >= T do not exist
    int p_tile = cluster_id, p_j = 0;
    uint32_t li = 0, lc = 0;
    auto issue_next = [&]() {
      while (p_tile < P.n_tiles) {
        int po0, pt0, pn;
        decode(p_tile, po0, pt0, pn);
        const int pc0 = chalf * (TC_BN / 2) + T2_ECOLS * p_j;
        if (p_j < (TC_BN / 2) / T2_ECOLS && pt0 + pc0 < p.T) {
          const int b_ = (int)(li % NLB);
          if (lane == 0) {
            mbar_expect_tx(bar_ld(warp, b_), T2_EBOX);
            tma_load_3d(sbox_u + (T2_NSB + b_) * T2_EBOX, &map_r, bar_ld(warp, b_), pt0 + pc0, po0 + q * 32, pn);
          }
          ++li;
          ++p_j;
          return;
        }
        p_j = 0;
        p_tile += n_clusters;
      }
    };
    if (ld_on) {
#pragma unroll 1
      for (int b_ = 0; b_ < NLB; ++b_) issue_next();
    }
    uint32_t ti = 0;
    for (int tile = cluster_id; tile < P.n_tiles; tile += n_clusters, ++ti) {
      int o0, t0, n;
      decode(tile, o0, t0, n);
      const int o = o0 + q * 32 + lane;   // M % 256 == 0: always valid
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;
      float bias_o = 0.f;
      if (e.bias) bias_o = __ldg(e.bias + o);
      if (e.row_bias) bias_o += __ldg(e.row_bias + (int64_t)n * p.M + o);
      float out_alpha = 1.f;
      if constexpr (EPI == 0) out_alpha = e.out_alpha ? __ldg(e.out_alpha) : 1.f;
      float mu2 = 0.f, r2 = 1.f, a2 = 1.f, gam2 = 0.f, mh = 0.f, mhy = 0.f, gam1 = 0.f, bet1 = 0.f, bdm = 0.f, w0 = 0.f, w1 = 0.f,
            w2 = 0.f;
      if constexpr (EPI == 10) {
        // per-row scalars were finalised by tcn_f2a_kernel (rowsc[6] = mean2, rowsc[7] = rstd2): no fp64 division /
        // sqrt per thread per tile here
        mu2 = (float)e.rowsc[8 * n + 6]; r2 = (float)e.rowsc[8 * n + 7];
        a2 = __ldg(e.a2); gam2 = __ldg(e.g2 + o);
        const float inv_cnt = 1.f / (float)e.count2;
        mh = (float)e.rowsc[8 * n + 0] * inv_cnt; mhy = (float)e.rowsc[8 * n + 1] * inv_cnt;
        gam1 = __ldg(e.g1 + o); bet1 = __ldg(e.be1 + o); bdm = __ldg(e.bd + o);
        w0 = __ldg(e.wd + 3 * o); w1 = __ldg(e.wd + 3 * o + 1); w2 = __ldg(e.wd + 3 * o + 2);
      }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, sL = 0.f, sR = 0.f;
      float cA = 0.f, cB = 0.f, cC = 0.f;
      if constexpr (EPI == 10) {
        cA = r2 * gam2;
        cB = -r2 * r2 * mhy;
        cC = -r2 * mh + r2 * r2 * mhy * mu2;
      }
      mbar_wait(bar_accf(a), aph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = chalf * (TC_BN / 2); c0 < (chalf + 1) * (TC_BN / 2); c0 += T2_ECOLS) {
        if (t0 + c0 >= p.T) break;
        const bool edge_chunk = (EPI == 10) && ((t0 + c0 < e.dil) || (t0 + c0 + T2_ECOLS > p.T - e.dil));
        float4 gop[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) gop[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (HAS_LD) {
          if (ld_on) {
            const int b_ = (int)(lc % NLB);
            mbar_wait(bar_ld(warp, b_), (lc / NLB) & 1);
            const uint8_t* lb = sbox + (T2_NSB + b_) * T2_EBOX + lrow;
#pragma unroll
            for (int g = 0; g < 4; ++g) gop[g] = *reinterpret_cast<const float4*>(lb + (((uint32_t)g ^ lsw) << 4));
            ++lc;
            __syncwarp();      // every lane has read the box: it can be refilled
            issue_next();
          }
        }
        uint32_t r[16];
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_BN + c0), r);
        float4 outv[4];
        [[maybe_unused]] float4 outv2[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int t = t0 + c0 + 4 * g;
          float v[4] = {__uint_as_float(r[4 * g]) + bias_o, __uint_as_float(r[4 * g + 1]) + bias_o,
                        __uint_as_float(r[4 * g + 2]) + bias_o, __uint_as_float(r[4 * g + 3]) + bias_o};
          if constexpr (EPI == 0) {
            outv[g] = make_float4(v[0], v[1], v[2], v[3]);
            if (e.out_stats) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float y = (t + i < p.T) ? prelu_f(v[i], out_alpha) : 0.f;
                s0 += y;
                s1 = fmaf(y, y, s1);
              }
            }
            if (e.ch_stats) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float y = (t + i < p.T) ? v[i] : 0.f;
                s2 += y;
                s3 = fmaf(y, y, s3);
              }
            }
          } else if constexpr (EPI == 1) {
            outv[g] = make_float4(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
          } else if constexpr (EPI == 3) {   // decoder masks: Y2 = relu(v) (stored first), Y = aux * relu(v)
            const float4 rr = gop[g];
            const float m0 = fmaxf(v[0], 0.f), m1 = fmaxf(v[1], 0.f), m2 = fmaxf(v[2], 0.f), m3 = fmaxf(v[3], 0.f);
            outv2[g] = make_float4(m0, m1, m2, m3);
            outv[g] = make_float4(rr.x * m0, rr.y * m1, rr.z * m2, rr.w * m3);
          } else if constexpr (EPI == 2) {
            const float4 rr = gop[g];
            outv[g] = make_float4(v[0] + rr.x, v[1] + rr.y, v[2] + rr.z, v[3] + rr.w);
          } else if constexpr (EPI == 10) {
            const float4 d4 = gop[g];
            const float draw[4] = {d4.x, d4.y, d4.z, d4.w};
            float dd[4];
            if (!edge_chunk) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float dvi = draw[i];
                const bool pos = dvi > 0.f;
                const float y2 = pos ? dvi : a2 * dvi;
                const float dy2 = fmaf(cA, v[i], fmaf(cB, y2, cC));
                const float ddv = pos ? dy2 : a2 * dy2;
                dd[i] = ddv;
                s0 += ddv;
                s1 = fmaf(ddv, dvi, s1);
                s3 += pos ? 0.f : dy2 * dvi;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int tt_ = t + i;
                const bool ok = tt_ < p.T;
                const float dvi = ok ? draw[i] : 1.f;
                const bool pos = dvi > 0.f;
                const float y2 = pos ? dvi : a2 * dvi;
                const float dy2 = ok ? fmaf(cA, v[i], fmaf(cB, y2, cC)) : 0.f;
                const float ddv = pos ? dy2 : a2 * dy2;
                dd[i] = ddv;
                s0 += ddv;
                s1 = fmaf(ddv, dvi, s1);
                s3 += pos ? 0.f : dy2 * dvi;
                sL += tt_ < e.dil ? ddv : 0.f;
                sR += tt_ >= p.T - e.dil ? ddv : 0.f;
              }
            }
            outv[g] = make_float4(dd[0], dd[1], dd[2], dd[3]);
          }
        }
        if constexpr (EPI == 3) {   // second output (the masks themselves) goes out through the other store box
          const uint32_t sb2 = (sc_i & 1u) * T2_EBOX;
          ++sc_i;
          if (lane == 0) bulk_wait_read1();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(sbox + sb2 + lrow + (((uint32_t)g ^ lsw) << 4)) = outv2[g];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_y2, sbox_u + sb2, t0 + c0, o0 + q * 32, n);
            bulk_commit();
          }
        }
        // registers -> store box (the TMA store issued two passes ago must have finished READING it) -> TMA store
        // (frames >= T are clipped)
        const uint32_t sb_off = (P.dbg & 8) ? 0u : (sc_i & 1u) * T2_EBOX;   // dbg 8: single store box (A/B timing)
        ++sc_i;
        if (lane == 0) {
          if (P.dbg & 8) bulk_wait_read0();
          else bulk_wait_read1();
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(sbox + sb_off + lrow + (((uint32_t)g ^ lsw) << 4)) = outv[g];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(P.dbg & 2)) {
          tma_store_3d(&map_y, sbox_u + sb_off, t0 + c0, o0 + q * 32, n);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // release the accumulator buffer: the LEADER's barrier gates the pair's next MMAs
        if (rank == 0) mbar_arrive(bar_acce(a));
        else mbar_arrive_cluster(bar_acce(a), 0);
      }
      if constexpr (EPI == 0) {
        if (e.out_stats) {
          s0 = warp_sum(s0);
          s1 = warp_sum(s1);
          if (lane == 0) {
            atomicAdd(e.out_stats + 2 * n, (double)s0);
            atomicAdd(e.out_stats + 2 * n + 1, (double)s1);
          }
        }
        if (e.ch_stats) {   // BatchNorm batch statistics: this lane owns channel o
          atomicAdd(e.ch_stats + 2 * o, (double)s2);
          atomicAdd(e.ch_stats + 2 * o + 1, (double)s3);
        }
      }
      if constexpr (EPI == 10) {
        const float S = s0, SD = s1;
        const float kS = (w0 + w1 + w2) * S - w0 * sL - w2 * sR;
        s0 = gam1 * kS;
        s1 = SD - bdm * S;
        s2 = bet1 * kS;
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) {
          atomicAdd(e.rowacc + 8 * n + 2, (double)s0);
          atomicAdd(e.rowacc + 8 * n + 3, (double)s1);
          atomicAdd(e.rowacc + 8 * n + 4, (double)s2);
          atomicAdd(e.rowacc + 8 * n + 5, (double)s3);
        }
      }
    }
    if (lane == 0) bulk_wait0();   // all of this warp's TMA stores are complete before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's smem / signalling its barriers until here
  if (warp == TC_W_TMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool gemm_wx_tc2_eligible(const GemmWxP& p, int pro, int epi) {
  if (g_tc_flags & 2) return false;              // debug switch: force the 1-CTA kernel
  if (p.M % 256) return false;
  if (!(epi == 0 || epi == 1 || epi == 2 || epi == 3 || epi == 10)) return false;
  return gemm_wx_tc_eligible(p, pro, epi);       // includes the 16-byte alignment of Y / R / d the TMA boxes need
}

template <int PRO, int EPI>
static int launch_tc2_t(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx, const CUtensorMap& my,
                        const CUtensorMap& mr, const CUtensorMap& my2, const TcParams& P, cudaStream_t st) {
  constexpr int NS = t2_stages(PRO, EPI);
  constexpr int SMEM = t2_smem_bytes(PRO, EPI);
  static_assert(SMEM <= 232448, "2-CTA kernel exceeds the 227 KB shared-memory limit");
  static_assert((3 * NS + 4 + 8 * T2_NL) * 8 + 4 <= T2_BAR_BYTES, "barrier area too small");
  auto k = gemm_wx_tc2_kernel<PRO, EPI, NS>;
  WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  int n_sm = 0;
  if (int rc = sm_count(&n_sm)) return rc;     // per current device (cached per device index)
  int clusters = P.n_tiles < n_sm / 2 ? P.n_tiles : n_sm / 2;
  k<<<2 * clusters, TC_THREADS, SMEM, st>>>(mh, ml, mx, my, mr, my2, P);
  WB_LAUNCH_CHECK("gemm_wx_tc2");
  return 0;
}

bool gemm_wx_tc2_instantiated(int pro, int epi) {
  return (pro == 0 && (epi == 0 || epi == 1 || epi == 2 || epi == 3 || epi == 10)) || (pro == 2 && epi == 2) || (pro == 3 && epi == 0);
}

int launch_gemm_wx_tc2(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx, TcParams P, int pro, int epi,
                       cudaStream_t st) {
  if (!gemm_wx_tc2_instantiated(pro, epi)) return -100;   // caller falls back to the 1-CTA kernel
  if (t2_stages(pro, epi) % P.xf_groups) P.xf_groups = 2;   // must divide the ring depth (see launch_gemm_wx_tc)
  P.n_ob = P.g.M / 256;                       // channel PAIRS
  P.n_tiles = P.n_ob * P.n_tt * P.g.n;
  // epilogue boxes: (16 frames, 32 channels, 1 row) with the 64-byte swizzle; frames >= T are clipped / zero-filled
  const GemmWxP& p = P.g;
  CUtensorMap my, mr, my2;
  const uint32_t ebox[3] = {T2_ECOLS, 32, 1};
  {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)p.M, (uint64_t)p.n};
    uint64_t strides[2] = {(uint64_t)p.ep.ldy * 4, (uint64_t)p.ep.bsy * 4};
    if (int rc = encode_map_sw(&my, p.ep.Y, 3, dims, strides, ebox, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    mr = my;
    my2 = my;
    if (epi == 3) {
      uint64_t s2[2] = {(uint64_t)p.ep.ldy2 * 4, (uint64_t)p.ep.bsy2 * 4};
      if (int rc = encode_map_sw(&my2, p.ep.Y2, 3, dims, s2, ebox, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
    if (epi == 2 || epi == 3) {
      uint64_t sr[2] = {(uint64_t)p.ep.ldr * 4, (uint64_t)p.ep.bsr * 4};
      if (int rc = encode_map_sw(&mr, p.ep.R, 3, dims, sr, ebox, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    } else if (epi == 10) {
      uint64_t sd[2] = {(uint64_t)p.ep.ldd * 4, (uint64_t)p.ep.bsd * 4};
      if (int rc = encode_map_sw(&mr, p.ep.d, 3, dims, sd, ebox, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    }
  }
  if (pro == 0 && epi == 0) return launch_tc2_t<0, 0>(mh, ml, mx, my, mr, my2, P, st);
  if (pro == 0 && epi == 1) return launch_tc2_t<0, 1>(mh, ml, mx, my, mr, my2, P, st);
  if (pro == 0 && epi == 2) return launch_tc2_t<0, 2>(mh, ml, mx, my, mr, my2, P, st);
  if (pro == 0 && epi == 3) return launch_tc2_t<0, 3>(mh, ml, mx, my, mr, my2, P, st);
  if (pro == 0 && epi == 10) return launch_tc2_t<0, 10>(mh, ml, mx, my, mr, my2, P, st);
  if (pro == 2 && epi == 2) return launch_tc2_t<2, 2>(mh, ml, mx, my, mr, my2, P, st);
  return launch_tc2_t<3, 0>(mh, ml, mx, my, mr, my2, P, st);
}

}  // namespace wb

// ================================================================================================
// 2-CTA variant of gemm_dw_tc (tcgen05 cta_group::2): a cluster of two CTAs accumulates a 256 x 256 tile of
// C += A f(B)^T.  Each CTA stages ITS 128 rows of A and ITS 128 rows of B per k-block (hi + lo: 32 KB per stage
// instead of 48 KB), so the ring is 6 deep instead of 4 and the per-SM shared-memory traffic of a stage (TMA fill +
// hi/lo transform + MMA operand reads) drops from ~144 KB to ~96 KB per 1060-cycle stage — the 1-CTA kernel sits
// right at the 128 B/clk limit.  Roles, stream-K unit ranges (per cluster) and barriers follow gemm_dw_tc_kernel /
// gemm_wx_tc2_kernel: the leader's MMA thread issues for the pair, commits are multicast, the peer's transform and
// epilogue warps signal the leader's barriers with remote arrives.
// ================================================================================================
namespace wb {

constexpr int D2_STAGES = 6;
constexpr int D2_A_BYTES = 128 * DW_BK * 4, D2_B_BYTES = 128 * DW_BK * 4;   // per CTA per stage: 8 KB each
constexpr int D2_OFF_AHI = 0, D2_OFF_ALO = D2_A_BYTES, D2_OFF_BHI = 2 * D2_A_BYTES, D2_OFF_BLO = 2 * D2_A_BYTES + D2_B_BYTES;
constexpr int D2_STAGE_BYTES = 2 * D2_A_BYTES + 2 * D2_B_BYTES;             // 32768
constexpr int D2_TX_BYTES = D2_A_BYTES + D2_B_BYTES;
constexpr int D2_SMEM_BYTES = D2_STAGES * D2_STAGE_BYTES + 256 + 1024;
// D=f32, A=B=tf32, both K-major, N=256, M=256 (cta_group::2)
constexpr uint32_t D2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

template <int PRO_B, bool RS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DW_THREADS, 1)
    gemm_dw_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const DwTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const GemmDwP& p = P.g;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar0 = base + D2_STAGES * D2_STAGE_BYTES;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (D2_STAGES + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * D2_STAGES + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (3 * D2_STAGES + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (3 * D2_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + D2_STAGES * D2_STAGE_BYTES + 8 * (3 * D2_STAGES + 4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int KBT = (p.T + DW_BK - 1) / DW_BK;
  const int64_t units = (int64_t)P.n_tiles * KBT;
  const int64_t u0 = units * cluster_id / n_clusters, u1 = units * (cluster_id + 1) / n_clusters;
  // unit -> (pair tile, k-block); tile = (row * n_cb + cb) * n_ob + ob2 (n_ob counts 256-row pairs)
  auto decode = [&](int tile, int& ob, int& cb, int& row) {
    ob = tile % P.n_ob;
    const int rest = tile / P.n_ob;
    cb = rest % P.n_cb;
    row = rest / P.n_cb;
  };

  if (tid == 0) {
    for (int s = 0; s < D2_STAGES; ++s) {
      mbar_init(bar_full(s), 1);     // local TMA transaction barrier
      mbar_init(bar_ready(s), 8);    // (leader's copy is used) 4 transform warps of the owning group x 2 CTAs
      mbar_init(bar_empty(s), 1);    // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf(a), 1);     // multicast tcgen05.commit
      mbar_init(bar_acce(a), 8);     // (leader's copy is used) 4 epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = u0; u < u1;) {
        const int tile = (int)(u / KBT), kb_lo = (int)(u % KBT);
        const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
        int ob, cb, row;
        decode(tile, ob, cb, row);
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % D2_STAGES;
          const uint32_t ph = (it / D2_STAGES) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          const uint32_t sb = base + s * D2_STAGE_BYTES;
          mbar_expect_tx(bar_full(s), D2_TX_BYTES);
          tma_load_3d(sb + D2_OFF_AHI, &map_a, bar_full(s), kb * DW_BK, ob * 256 + (int)rank * 128, row);
          tma_load_3d(sb + D2_OFF_BHI, &map_b, bar_full(s), kb * DW_BK, cb * 256 + (int)rank * 128, row);
        }
        u += kb_hi - kb_lo;
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && rank == 0) {   // leader CTA issues for the pair
      uint32_t it = 0, sg = 0;
      for (int64_t u = u0; u < u1; ++sg) {
        const int kb_lo = (int)(u % KBT);
        const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
        const int a = sg & 1;
        mbar_wait(bar_acce(a), ((sg >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * 256;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % D2_STAGES;
          const uint32_t ph = (it / D2_STAGES) & 1;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint32_t sb = base + s * D2_STAGE_BYTES;
          if (P.mixed) {      // [bf16(x) | bf16(x - tf32 x)] tiles of 128 rows x 32 B sit where the fp32 lo tile was
            tc_mma_bf16_2cta(d_tmem, make_desc_k16_sw32(sb + D2_OFF_ALO + 4096), make_desc_k16_sw32(sb + D2_OFF_BLO), D2_IDESC_BF16,
                             kb != kb_lo ? 1u : 0u);
            tc_mma_bf16_2cta(d_tmem, make_desc_k16_sw32(sb + D2_OFF_ALO), make_desc_k16_sw32(sb + D2_OFF_BLO + 4096), D2_IDESC_BF16, 1u);
#pragma unroll
            for (int ks = 0; ks < DW_BK / 8; ++ks)
              tc_mma_tf32_2cta(d_tmem, make_desc_k_sw64(sb + D2_OFF_AHI + ks * 32), make_desc_k_sw64(sb + D2_OFF_BHI + ks * 32), D2_IDESC, 1u);
          } else {
#pragma unroll
            for (int ks = 0; ks < DW_BK / 8; ++ks) {
              const uint64_t a_hi = make_desc_k_sw64(sb + D2_OFF_AHI + ks * 32);
              const uint64_t a_lo = make_desc_k_sw64(sb + D2_OFF_ALO + ks * 32);
              const uint64_t b_hi = make_desc_k_sw64(sb + D2_OFF_BHI + ks * 32);
              const uint64_t b_lo = make_desc_k_sw64(sb + D2_OFF_BLO + ks * 32);
              tc_mma_tf32_2cta(d_tmem, a_lo, b_hi, D2_IDESC, (kb != kb_lo || ks != 0) ? 1u : 0u);
              tc_mma_tf32_2cta(d_tmem, a_hi, b_lo, D2_IDESC, 1u);
              tc_mma_tf32_2cta(d_tmem, a_hi, b_hi, D2_IDESC, 1u);
            }
          }
          tc_commit_mc2(bar_empty(s));   // frees stage s in BOTH CTAs
        }
        tc_commit_mc2(bar_accf(a));      // accumulators complete in BOTH CTAs
        u += kb_hi - kb_lo;
      }
    }
  } else if (warp >= 6) {
    const int tt_id = tid - 6 * 32;
    float alpha = 1.f;
    if constexpr (PRO_B >= 1) alpha = p.xb.alpha ? __ldg(p.xb.alpha) : 1.f;
    const int xf_gid = tt_id >> 7, xf_tid = tt_id & 127;   // D2_STAGES % 2 == 0: a slot always belongs to the same group
    const int n_it = (int)(u1 - u0);
    constexpr int A_F4 = D2_A_BYTES / 16 / 128;       // 4
    constexpr int B_F4 = D2_B_BYTES / 16 / 128;       // 4
    float rs[A_F4];
#pragma unroll
    for (int i = 0; i < A_F4; ++i) rs[i] = 0.f;
    auto flush_rs = [&](int tile) {
      int ob, cb, row;
      decode(tile, ob, cb, row);
      if (RS && cb == 0) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
          const int o = ob * 256 + (int)rank * 128 + (xf_tid + 128 * i) / 4;
          if (o < p.M) atomicAdd(p.a_rowsum + (int64_t)row * p.M + o, rs[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < A_F4; ++i) rs[i] = 0.f;
    };
    float bsc[B_F4], bsh[B_F4];
#pragma unroll
    for (int j = 0; j < B_F4; ++j) { bsc[j] = 1.f; bsh[j] = 0.f; }
    auto load_scsh = [&](int tile) {
      if constexpr (PRO_B >= 2) {
        int ob, cb, row;
        decode(tile, ob, cb, row);
        float mu = 0.f, r = 1.f;
        if (p.xb.row_stats) gln_mean_rstd(p.xb.row_stats + 2 * row, p.xb.count, p.xb.eps, mu, r);
#pragma unroll
        for (int j = 0; j < B_F4; ++j) {
          const int c = cb * 256 + (int)rank * 128 + (xf_tid + 128 * j) / 4;
          const float gm = (p.xb.ch_scale && c < p.N) ? __ldg(p.xb.ch_scale + c) : 1.f;
          const float bt = (p.xb.ch_shift && c < p.N) ? __ldg(p.xb.ch_shift + c) : 0.f;
          bsc[j] = gm * r;
          bsh[j] = bt - gm * mu * r;
        }
      }
    };
    int cur_tile = (int)(u0 / KBT);
    int next_b = KBT - (int)(u0 % KBT);
    if (n_it > 0) load_scsh(cur_tile);
    for (int it = xf_gid; it < n_it; it += 2) {
      const int s = it % D2_STAGES;
      const uint32_t ph = (it / D2_STAGES) & 1;
      while (it >= next_b) {
        if constexpr (RS) flush_rs(cur_tile);
        ++cur_tile;
        next_b += KBT;
        load_scsh(cur_tile);
      }
      mbar_wait(bar_full(s), ph);
      uint8_t* st = gbase + s * D2_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < A_F4 + B_F4; ++i) {
        const int idx = xf_tid + 128 * i;
        const bool is_b = i >= A_F4;
        const int off = is_b ? (idx * 16 - D2_A_BYTES) : idx * 16;
        uint8_t* hi_p = st + (is_b ? D2_OFF_BHI : D2_OFF_AHI) + off;
        uint8_t* lo_p = st + (is_b ? D2_OFF_BLO : D2_OFF_ALO) + off;
        float4 v = *reinterpret_cast<const float4*>(hi_p);
        if constexpr (RS) {
          if (!is_b) rs[is_b ? 0 : i] += (v.x + v.y) + (v.z + v.w);
        }
        if constexpr (PRO_B == 1) {
          if (is_b) { v.x = prelu_f(v.x, alpha); v.y = prelu_f(v.y, alpha); v.z = prelu_f(v.z, alpha); v.w = prelu_f(v.w, alpha); }
        }
        if constexpr (PRO_B == 2) {
          if (is_b) {
            const float c_ = bsc[is_b ? i - A_F4 : 0], d_ = bsh[is_b ? i - A_F4 : 0];
            v.x = fmaf(c_, prelu_f(v.x, alpha), d_); v.y = fmaf(c_, prelu_f(v.y, alpha), d_);
            v.z = fmaf(c_, prelu_f(v.z, alpha), d_); v.w = fmaf(c_, prelu_f(v.w, alpha), d_);
          }
        }
        if constexpr (PRO_B == 3) {
          if (is_b) {
            const float c_ = bsc[is_b ? i - A_F4 : 0], d_ = bsh[is_b ? i - A_F4 : 0];
            v.x = prelu_f(fmaf(c_, v.x, d_), alpha); v.y = prelu_f(fmaf(c_, v.y, d_), alpha);
            v.z = prelu_f(fmaf(c_, v.z, d_), alpha); v.w = prelu_f(fmaf(c_, v.w, d_), alpha);
          }
        }
        if (P.mixed) {
          // raw K-major tile: rows of 64 B, 16-byte chunk XORed with (row >> 1) & 3 -> logical K offset of this float4; the
          // bf16 tiles have rows of 32 B with the 32-byte swizzle (chunk ^= (row >> 2) & 1)
          const uint32_t r = (uint32_t)off >> 6, lc = (((uint32_t)off >> 4) & 3u) ^ ((r >> 1) & 3u);
          const uint32_t o16 = r * 32u + (((lc >> 1) ^ ((r >> 2) & 1u)) << 4) + (lc & 1u) * 8u;
          uint8_t* t16 = st + (is_b ? D2_OFF_BLO : D2_OFF_ALO);
          uint2 bh, bl;
          split_bf16_quad(v, bh, bl);
          if (PRO_B >= 1 && is_b) *reinterpret_cast<float4*>(hi_p) = v;   // the tensor core truncates to tf32 itself
          *reinterpret_cast<uint2*>(t16 + o16) = bh;
          *reinterpret_cast<uint2*>(t16 + 4096 + o16) = bl;
        } else {
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
          if (!P.skip_hi_store || (PRO_B >= 1 && is_b)) *reinterpret_cast<float4*>(hi_p) = h;
          *reinterpret_cast<float4*>(lo_p) = l;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {   // signal the LEADER's ready barrier (it gates the pair's MMAs)
        if (rank == 0) mbar_arrive(bar_ready(s));
        else mbar_arrive_cluster(bar_ready(s), 0);
      }
    }
    if constexpr (RS) {
      if (n_it > 0) flush_rs(cur_tile);
    }
  } else {
    // epilogue warps 0..3: C[o][c] += D on this CTA's 128 rows, for every segment of the cluster's range
    const int q = warp;
    uint32_t sg = 0;
    for (int64_t u = u0; u < u1; ++sg) {
      const int tile = (int)(u / KBT), kb_lo = (int)(u % KBT);
      const int kb_hi = (int)min((int64_t)KBT, (int64_t)kb_lo + (u1 - u));
      u += kb_hi - kb_lo;
      int ob, cb, row;
      decode(tile, ob, cb, row);
      const int a = sg & 1;
      const int o = ob * 256 + (int)rank * 128 + q * 32 + lane;   // M % 256 == 0: always valid
      float* C = p.C + (p.per_row ? (int64_t)row * p.M * p.ldc : 0) + (int64_t)o * p.ldc + cb * 256;
      const int ncols = min(256, p.N - cb * 256);
      mbar_wait(bar_accf(a), (sg >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        if (c0 >= ncols) break;                       // warp-uniform
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + c0), r);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < ncols) atomicAdd(C + c0 + i, __uint_as_float(r[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar_acce(a));
        else mbar_arrive_cluster(bar_acce(a), 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's smem / signalling its barriers until here
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool gemm_dw_tc2_eligible(const GemmDwP& p, int pro_b) {
  if (g_tc_flags & (2 | 512)) return false;      // debug switches: bit 1 = all GEMMs 1-CTA, bit 9 = only the dW GEMMs
  if (p.M % 256) return false;
  return gemm_dw_tc_eligible(p, pro_b);
}

int launch_gemm_dw_tc2(const GemmDwP& p, int pro_b, cudaStream_t st) {
  CUtensorMap ma, mb;
  const uint32_t box[3] = {DW_BK, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)p.M, (uint64_t)p.n};
    uint64_t strides[2] = {(uint64_t)p.lda * 4, (uint64_t)p.bsa * 4};
    if (int rc = encode_map_sw(&ma, p.A, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.T, (uint64_t)p.N, (uint64_t)p.n};
    uint64_t strides[2] = {(uint64_t)p.ldb * 4, (uint64_t)p.bsb * 4};
    if (int rc = encode_map_sw(&mb, p.B, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  DwTcParams P;
  P.g = p;
  P.n_ob = p.M / 256;
  P.n_cb = cdiv(p.N, 256);
  P.n_tiles = P.n_ob * P.n_cb * p.n;
  P.skip_hi_store = (g_tc_flags & 1) ? 0 : 1;
  P.mixed = (g_tc_flags & 1024) ? 0 : 1;   // tf32 leading term + bf16 cross terms (default)
  int n_sm = 0;
  if (int rc = sm_count(&n_sm)) return rc;     // per current device (cached per device index)
  const int max_clusters = n_sm / 2;
  const int64_t units = (int64_t)P.n_tiles * cdiv(p.T, DW_BK);
  int clusters = max_clusters;
  if (units < (int64_t)clusters * 8) clusters = (int)(units / 8 > 0 ? units / 8 : 1);
  // whole tiles per cluster when they nearly fill the machine (see launch_gemm_dw_tc); bit 8 of the flags forces stream-K
  if (!(g_tc_flags & 256) && P.n_tiles <= max_clusters && P.n_tiles * 4 >= max_clusters * 3) clusters = P.n_tiles;
  auto launch = [&](auto k) -> int {
    WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_BYTES));
    k<<<2 * clusters, DW_THREADS, D2_SMEM_BYTES, st>>>(ma, mb, P);
    return 0;
  };
  int rc;
  if (p.a_rowsum) {
    if (pro_b > 1) return fail(-2, "gemm_dw_tc2: a_rowsum needs pro_b 0 or 1");
    rc = pro_b == 0 ? launch(gemm_dw_tc2_kernel<0, true>) : launch(gemm_dw_tc2_kernel<1, true>);
  } else {
    rc = pro_b == 0 ? launch(gemm_dw_tc2_kernel<0, false>) : pro_b == 1 ? launch(gemm_dw_tc2_kernel<1, false>)
         : pro_b == 2 ? launch(gemm_dw_tc2_kernel<2, false>) : launch(gemm_dw_tc2_kernel<3, false>);
  }
  if (rc) return rc;
  WB_LAUNCH_CHECK("gemm_dw_tc2");
  return 0;
}

}  // namespace wb
