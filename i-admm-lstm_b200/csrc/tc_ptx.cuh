// tcgen05 / TMA / mbarrier PTX wrappers and shared-memory descriptor helpers shared by the tensor-core kernels
// (gates_tc.cu, gemm_tc.cu).  sm_100a only.
#pragma once

#include <stdio.h>
#include <stdint.h>
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is resolved at run time)

#include "common.cuh"

namespace iadmm {

constexpr int kTcBM = 128;            // rows per CTA tile (UMMA M per CTA)
constexpr int kTcUK = 16;             // UMMA K for 16-bit inputs

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.  The try_wait carries a
// suspend-time hint, so a waiting warp sleeps in hardware until the phase completes instead of re-issuing the poll loop:
// ncu showed 30 % of the gate kernel's executed instructions to be poll-loop instructions of waiting warps, and the kernel
// is bound by the board power cap (every issued instruction costs clock).
#ifndef IADMM_MBAR_SPIN_LIMIT
#define IADMM_MBAR_SPIN_LIMIT (1u << 22)
#endif
#ifndef IADMM_MBAR_SUSPEND_NS
#define IADMM_MBAR_SUSPEND_NS 4000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t suspend_ns = IADMM_MBAR_SUSPEND_NS) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(suspend_ns)
        : "memory");
    if (!done && ++spins > IADMM_MBAR_SPIN_LIMIT) {
      printf("iadmm gates_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// plain (non-tensor) bulk copy global -> this CTA's shared memory, completion bytes credited to a local mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Relaxed: the barrier only orders the (already completed, tcgen05.wait::ld) TMEM reads of the epilogue against the
// next MMAs into that accumulator; no global-memory data is published through it, so no MEMBAR is needed
// (the .release form cost 8 % of the kernel's stall samples in fences).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrives on the barrier at this CTA-relative offset in every CTA of `cta_mask` once the issued MMAs retire
__device__ __forceinline__ void tc_commit_pair(uint32_t bar_local, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_local),
               "h"(cta_mask)
               : "memory");
}
// TMA load delivered to the same CTA-relative smem offset in every CTA of `cta_mask`; the completion bytes are
// credited, per destination CTA, to the mbarrier at `bar_local`'s offset in the LEADER of that CTA's pair
// (peer bit 24 of the rank-encoded shared address cleared)
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar_local, int c0, int c1,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(map), "r"(bar_local & 0xFEFFFFFFu), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 64 bytes, 64B swizzle (as written by TMA):
// 8-row groups are 512 B apart (stride byte offset); leading byte offset is unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units        bits [0,14)
  d |= (uint64_t)(0) << 16;                          // leading byte offset                 bits [16,30)
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset                  bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)      bits [46,48)
  d |= (uint64_t)4 << 61;                            // layout type SWIZZLE_64B             bits [61,64)
  return d;
}
// rows of 128 bytes, 128B swizzle: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
// same for rows of 32 bytes (32 fp8 elements), 32B swizzle: 8-row groups are 256 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                            // layout type SWIZZLE_32B
  return d;
}
// K-major operand tile WITHOUT swizzle (the canonical "interleaved" layout): core matrices of 8 rows x 16 bytes are
// 128 contiguous bytes; consecutive 8-row groups are 128 B apart (stride byte offset), the next 16-byte K chunk of the
// same rows is `lbo_bytes` away (leading byte offset).  This is the order [K group][row][16 B] the row-interleaved
// global layout is stored in, so a TMA box lands in it unchanged.
__device__ __forceinline__ uint64_t make_smem_desc_il(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;                            // layout type 0 = no swizzle
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n_cols
__device__ __forceinline__ uint32_t make_idesc_f16(int n_cols, int m_rows = kTcBM) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
         ((uint32_t)(m_rows >> 4) << 24);
}


// ---- host side: tensor maps ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
// 2D row-major [rows_total][cols] tensor of fp16 (elem_bytes 2) or bytes (elem_bytes 1); box = [box_rows][box_k
// elements]; swizzle follows the box row size (128/64/32 bytes); out-of-bounds elements read as zero
int make_map(CUtensorMap* map, const void* base, uint64_t rows_total, int cols, int box_rows, int elem_bytes, int box_k);
// row-interleaved operand images (no swizzle): fp16 [groups][rows_total][8] as a 3D tensor {64 elements = 8 rows x 16 B,
// rows_total/8, groups}, box {64, 16, 8} = 128 rows x 64 K; e4m3 [groups16][2][rows_total][16] as a 4D tensor
// {128 bytes, rows_total/8, 2, groups16}, box {128, 16, 2, 4}.  rows_total % 8 == 0.
int make_map_il(CUtensorMap* map, const void* base, uint64_t rows_total, int groups, bool q8, int box_k = 64);

}  // namespace iadmm
