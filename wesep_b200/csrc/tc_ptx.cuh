// PTX wrappers shared by the tcgen05 kernels written after gemm_tc.cu (which keeps its own copies): mbarrier, cluster /
// DSMEM, bulk copies, tcgen05 alloc / mma / commit / ld, UMMA shared-memory descriptors.  sm_100a only.
#pragma once
#include <stdint.h>

namespace wb {
namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp: ptxas treats code predicated on elect.sync as warp-uniform, so the tcgen05 instructions
// inside are emitted straight (no per-instruction ELECT / BRA.U.ANY serialisation loop as under `if (lane == 0)`)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// one arrival + `bytes` more transaction bytes expected in the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// cluster-scope acquire: the waiter reads data other CTAs of the cluster wrote with st.shared::cluster before arriving
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAITC_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAITC_DONE;\n"
      "bra WAITC_LOOP;\n"
      "WAITC_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------- cluster / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar) : "memory");
}
// release at cluster scope: publishes this thread's earlier st.shared::cluster stores to the waiter
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// bulk copy own shared memory -> shared memory of another CTA of the cluster; completes `bytes` transaction bytes on the
// (remote) mbarrier.  The source must have been made visible to the async proxy (fence.proxy.async) by its writers.
__device__ __forceinline__ void bulk_copy_s2c(uint32_t cluster_dst, uint32_t src, uint32_t bytes, uint32_t cluster_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(cluster_dst),
               "r"(src), "r"(bytes), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// one arrival on the barrier at this offset in every CTA of `mask` once all earlier MMAs of this thread have completed
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], 16-bit operands (fp16 or bf16 per the instruction descriptor), fp32 accumulate, K = 16
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand with the 64-byte swizzle (cute::UMMA::SmemDescriptor: start[0,14)
// lbo[16,30) sbo[32,46) version[46,48)=1 layout_type[61,64)=4): rows of 64 B (32 sixteen-bit elements of K), 8-row atoms
// of 512 B -> SBO = 512 B; LBO unused.  Within an atom the 16-byte chunk index is XORed with (row >> 1) & 3
// (address bits [4,6) ^= bits [7,9)), so the tile base must be 512-byte aligned.  A K step of 16 elements (32 B)
// inside the row is addressed by adding 32 to the start address (the same rule gemm_dw_tc uses for tf32).
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// byte offset of element (row, 16-byte chunk c in [0,4)) inside a [rows x 64 B] K-major SWIZZLE_64B slab
__device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

// MN-major 16-bit operand with the 128-byte swizzle (layout_type 2): rows = K index, 64 MN elements (128 B) contiguous per
// row, atoms of 8 K rows (1024 B) -> SBO = 1024 B between the two K atoms of one K = 16 instruction; LBO (stride between
// MN atoms) is unused while the MN extent is <= 64.  Within an atom the 16-byte chunk index is XORed with (K row & 7).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of (K row, 16-byte chunk c in [0,8)) inside a [K rows x 128 B] MN-major SWIZZLE_128B slab (1024-byte aligned)
__device__ __forceinline__ uint32_t sw128_off(uint32_t krow, uint32_t chunk) { return krow * 128u + ((chunk ^ (krow & 7u)) << 4); }

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format[4,6)=1 (f32), a_format[7,10), b_format[10,13)
// (kind::f16: 0 = f16, 1 = bf16), a_major[15], b_major[16] (0 = K-major), n_dim[17,23) = N >> 3, m_dim[24,29) = M >> 4.
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int fmt /* 0 = f16, 1 = bf16 */, int b_mn_major = 0) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

}  // namespace tcx
// LDGSTS: 16-byte asynchronous global -> shared copy (L2 only); src_bytes < 16 zero-fills the rest (0: pure zero fill)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// commit / wait WITH a compiler memory barrier (the shared-memory reads that follow must not be hoisted above the wait)
__device__ __forceinline__ void cp_async_commit_mem() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_mem() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

}  // namespace wb
