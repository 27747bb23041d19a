// Coordinate-wise LSTM cell on the 5th-generation tensor cores (IADMM_GATES_TC_3XFP16 / _1XFP16).
//
// Reference: models/lstm.py:74-80.  The four gate products H_t @ U_{i,f,o,u} over all rows = B*(n+m)
// coordinates are ONE [rows,h] x [h,4h] GEMM with weights shared by every row: the only tensor-core
// shaped work on the path (8*rows*h^2 flop per iteration, >99% of the arithmetic at h=800).
//
// Precision.  The reference computes in fp32 and north_star demands rel <= 1e-4 after K=100, which
// rules out bf16 and makes single-pass tf32/fp16 marginal.  Both operands
// are therefore split into fp16 hi+lo pairs of power-of-two scaled values (H*2^14, U*2^s), and
//     H U  ~=  (H_hi U_hi + H_lo U_hi + H_hi U_lo) * 2^-(14+s)
// is accumulated in fp32 in tensor memory: ~22-bit operands for 3 MMAs at the fp16 rate (cheaper than
// one tf32 MMA per bit of accuracy, and 4 B/element of operand traffic like fp32).  H is kept in this
// hi/lo form between iterations (written by the epilogue), so nothing is re-split on the fly.
//
// Kernel structure (persistent, one CTA per SM, warp specialised):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D tiles (64B swizzle) of H_hi/H_lo [128 x 32] and
//                              U_hi/U_lo [256 x 32] into a 4..8-stage shared-memory ring (mbarrier tx)
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128,
//                              N=256 = 64 hidden units x 4 interleaved gates, K=16), accumulators in
//                              TMEM, double buffered (2 x 256 columns); tcgen05.commit frees smem
//                              stages and publishes finished accumulators
//   warps 2..9  epilogue       tcgen05.ld 32 columns (= 8 units x 4 gates) per step, add the rank-2
//                              input term and bias, sigmoid/tanh, C and H update, fp16 hi/lo re-split of
//                              H, partial dot with W_h -- the whole cell, fused; overlaps the next
//                              tile's MMAs
// Tiles are ordered unit-tile fastest, so CTAs running concurrently share H row tiles through L2 and H is
// read from HBM once per iteration.
#include "common.cuh"

#include <stdio.h>
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is resolved at run time)

namespace iadmm {

constexpr int kTcBM = 128;            // rows per tile (UMMA M)
constexpr int kTcBN = 256;            // gate columns per tile (UMMA N) = 64 hidden units
constexpr int kTcUnits = kTcBN / 4;
constexpr int kTcBK = 32;             // K elements per stage (64-byte swizzled rows of fp16)
constexpr int kTcUK = 16;             // UMMA K for 16-bit inputs
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr int kTcABytes = kTcBM * kTcBK * 2;     // 8 KB
constexpr int kTcBBytes = kTcBN * kTcBK * 2;     // 16 KB
constexpr int kTcChunk = 32;                     // TMEM columns per epilogue step (8 units)

int tc_gate_tiles(int h) { return 2 * cdiv(h, kTcUnits); }   // two head partials per unit tile
size_t tc_state_bytes(long rows, int h) { return (size_t)rows * h * sizeof(__half) * 4; }

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
// Bounded spin: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
#ifndef IADMM_MBAR_SPIN_LIMIT
#define IADMM_MBAR_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
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
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n_cols
__device__ __forceinline__ uint32_t make_idesc_f16(int n_cols) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
         ((uint32_t)(kTcBM >> 4) << 24);
}

// accurate-enough transcendental pieces for the epilogue (errors ~1e-7, far below the gate-GEMM split error)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
  const float* wc;        // [2][4h]
  const float* bias;      // [4h]
  const float* wh;        // [h]
  const float* scale;     // [4]: [1] = dequant
  const float* xv;        // [rows]
  const float* g;         // [rows]
  __half* hout_hi;        // [rows][h]
  __half* hout_lo;
  float*  hout_f32;       // optional
  float*  C;              // [rows][h] in place
  float*  head_part;      // [2*unit_tiles][rows]
  long rows;
  int  h, unit_tiles, k_blocks, stages, nprod;
  long num_tiles;
};

template <int NPROD>
__global__ void __launch_bounds__(kTcThreads, 1)
gates_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A_hi | A_lo | B_hi | B_lo)] | params[2][3*256+64] floats | barriers
  constexpr int kStageBytes = (NPROD == 3) ? 2 * (kTcABytes + kTcBBytes) : (kTcABytes + kTcBBytes);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = P.stages;
  float* sparam = reinterpret_cast<float*>(smem + (size_t)stages * kStageBytes);
  constexpr int kParamFloats = 3 * kTcBN + kTcUnits;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sparam + 2 * kParamFloats);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + stages;           // [stages]
  uint64_t* tfull_bar = bars + 2 * stages;       // [2]
  uint64_t* tempty_bar = bars + 2 * stages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_b_hi);
    if (NPROD == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int unit_tiles = P.unit_tiles;
  const int h4 = 4 * P.h;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int  ut = (int)(tile % unit_tiles);
        const long rt = tile / unit_tiles;
        const int row0 = (int)(rt * kTcBM);
        const int col0 = ut * kTcBN;
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, kStageBytes);
          uint8_t* sbase = smem + (size_t)stage * kStageBytes;
          const int k0 = kb * kTcBK;
          tma_load_2d(smem_u32(sbase), &map_a_hi, fb, k0, row0);
          if (NPROD == 3) {
            tma_load_2d(smem_u32(sbase + kTcABytes), &map_a_lo, fb, k0, row0);
            tma_load_2d(smem_u32(sbase + 2 * kTcABytes), &map_b_hi, fb, k0, col0);
            tma_load_2d(smem_u32(sbase + 2 * kTcABytes + kTcBBytes), &map_b_lo, fb, k0, col0);
          } else {
            tma_load_2d(smem_u32(sbase + kTcABytes), &map_b_hi, fb, k0, col0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      long it = 0;
      for (long tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
        const int ut = (int)(tile % unit_tiles);
        const int n_cols = min(kTcBN, h4 - ut * kTcBN);
        const uint32_t idesc = make_idesc_f16(n_cols);
        const int buf = (int)(it & 1);
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&tempty_bar[buf]), (use & 1) ^ 1);     // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcBN);
        uint32_t acc = 0;
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * kStageBytes);
          const int k_len = min(kTcBK, P.h - kb * kTcBK);
          const int k_steps = (k_len + kTcUK - 1) / kTcUK;
          for (int ks = 0; ks < k_steps; ++ks) {
            const uint32_t koff = (uint32_t)(ks * kTcUK * 2);      // bytes inside the 64-byte swizzled row
            if (NPROD == 3) {
              const uint64_t a_hi = make_smem_desc_sw64(sbase + koff);
              const uint64_t a_lo = make_smem_desc_sw64(sbase + kTcABytes + koff);
              const uint64_t b_hi = make_smem_desc_sw64(sbase + 2 * kTcABytes + koff);
              const uint64_t b_lo = make_smem_desc_sw64(sbase + 2 * kTcABytes + kTcBBytes + koff);
              tc_mma_f16(d_tmem, a_lo, b_hi, idesc, acc); acc = 1;     // small terms first
              tc_mma_f16(d_tmem, a_hi, b_lo, idesc, 1);
              tc_mma_f16(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              const uint64_t a_hi = make_smem_desc_sw64(sbase + koff);
              const uint64_t b_hi = make_smem_desc_sw64(sbase + kTcABytes + koff);
              tc_mma_f16(d_tmem, a_hi, b_hi, idesc, acc); acc = 1;
            }
          }
          tc_commit(smem_u32(&empty_bar[stage]));                  // frees the smem stage when the MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(smem_u32(&tfull_bar[buf]));                      // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: the LSTM cell =====================
    const int ew = warp - 2;                       // 0..7
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may touch
    const int half = (ew >= 4) ? 1 : 0;            // which half of the tile's column chunks
    const int et = threadIdx.x - 64;               // 0..255
    const float dequant = P.scale[1];
    long it = 0;
    for (long tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int  ut = (int)(tile % unit_tiles);
      const long rt = tile / unit_tiles;
      const int buf = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      const int col0 = ut * kTcBN;
      // stage this tile's W rows / bias / W_h slice (shared by all rows)
      float* sp = sparam + buf * kParamFloats;
      {
        const int c = col0 + et;
        const bool ok = c < h4;
        sp[et]             = ok ? __ldg(P.wc + c) : 0.f;
        sp[kTcBN + et]     = ok ? __ldg(P.wc + h4 + c) : 0.f;
        sp[2 * kTcBN + et] = ok ? __ldg(P.bias + c) : 0.f;
        if (et < kTcUnits) {
          const int u = ut * kTcUnits + et;
          sp[3 * kTcBN + et] = (u < P.h) ? __ldg(P.wh + u) : 0.f;
        }
      }
      const long row = rt * kTcBM + quarter * 32 + lane;
      const bool row_ok = row < P.rows;
      const float xr = row_ok ? __ldg(P.xv + row) : 0.f;
      const float gr = row_ok ? __ldg(P.g + row) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");

      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1);
      tc_fence_after();

      float hp = 0.f;
      constexpr int kChunksPerHalf = kTcBN / kTcChunk / 2;   // 4
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerHalf; ++cc) {
        const int chunk = half * kChunksPerHalf + cc;
        const int unit0 = ut * kTcUnits + chunk * 8;          // first hidden unit of this chunk
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kTcBN + chunk * kTcChunk);
        tc_ld32(taddr, v);
        tc_wait_ld();
        if (row_ok && unit0 < P.h) {
          const size_t o = (size_t)row * P.h + unit0;
          const float4 c_lo = *reinterpret_cast<const float4*>(P.C + o);
          const float4 c_hi = *reinterpret_cast<const float4*>(P.C + o + 4);
          const float cold[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
          float cnew[8], hnew[8];
          const float* w0 = sp + chunk * kTcChunk;
          const float* w1 = sp + kTcBN + chunk * kTcChunk;
          const float* bb = sp + 2 * kTcBN + chunk * kTcChunk;
          const float* wh = sp + 3 * kTcBN + chunk * 8;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float pre[4];
#pragma unroll
            for (int gte = 0; gte < 4; ++gte) {
              const int j = u * 4 + gte;
              const float iw = fmaf(gr, w1[j], __fmul_rn(xr, w0[j]));
              pre[gte] = __fadd_rn(__fadd_rn(iw, __fmul_rn(__uint_as_float(v[j]), dequant)), bb[j]);
            }
            const float gi = sigmoid_fast(pre[0]);
            const float gf = sigmoid_fast(pre[1]);
            const float go = sigmoid_fast(pre[2]);
            const float gu = tanhf(pre[3]);
            const float cn = __fadd_rn(__fmul_rn(gi, gu), __fmul_rn(gf, cold[u]));
            const float hn = __fmul_rn(go, tanhf(cn));
            cnew[u] = cn;
            hnew[u] = hn;
            hp = fmaf(hn, wh[u], hp);
          }
          *reinterpret_cast<float4*>(P.C + o)     = make_float4(cnew[0], cnew[1], cnew[2], cnew[3]);
          *reinterpret_cast<float4*>(P.C + o + 4) = make_float4(cnew[4], cnew[5], cnew[6], cnew[7]);
          // fp16 hi/lo image of H * 2^14 for the next iteration's MMAs
          __align__(16) __half hh[8];
          __align__(16) __half hl[8];
          const float hs = (float)(1 << kHShift);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float s = hnew[u] * hs;
            hh[u] = __float2half_rn(s);
            hl[u] = __float2half_rn(s - __half2float(hh[u]));
          }
          *reinterpret_cast<uint4*>(P.hout_hi + o) = *reinterpret_cast<const uint4*>(hh);
          if (NPROD == 3) *reinterpret_cast<uint4*>(P.hout_lo + o) = *reinterpret_cast<const uint4*>(hl);
          if (P.hout_f32) {
            *reinterpret_cast<float4*>(P.hout_f32 + o)     = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
            *reinterpret_cast<float4*>(P.hout_f32 + o + 4) = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
          }
        }
      }
      if (row_ok) P.head_part[((size_t)ut * 2 + half) * P.rows + row] = hp;
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D fp16 row-major [rows_total][h] tensor, box = [box_rows][32], 64B swizzle, OOB reads as zero
static int make_map(CUtensorMap* map, const void* base, uint64_t rows_total, int h, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) IADMM_FAIL(IADMM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)h, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)h * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IADMM_FAIL(IADMM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu h=%d)", (int)r,
                                    (unsigned long long)rows_total, h);
  return IADMM_OK;
}

int launch_gates_tc(const void* packed, const WeightLayout& L, const float* xv, const float* g, const __half* Hin_hi,
                    const __half* Hin_lo, __half* Hout_hi, __half* Hout_lo, float* H_out_f32, float* C,
                    float* head_part, long rows, int h, int nprod, cudaStream_t st) {
  if (h % 8 != 0) IADMM_FAIL(IADMM_EMODE, "tensor-core gate path needs hidden_dim %% 8 == 0");
  if (rows > 0x7fffffffL - kTcBM) IADMM_FAIL(IADMM_ESHAPE, "too many rows for one launch");
  const char* base = static_cast<const char*>(packed);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, Hin_hi, (uint64_t)rows, h, kTcBM))) return rc;
  if ((rc = make_map(&ma_lo, Hin_lo, (uint64_t)rows, h, kTcBM))) return rc;
  if ((rc = make_map(&mb_hi, base + L.off_uhi, (uint64_t)4 * h, h, kTcBN))) return rc;
  if ((rc = make_map(&mb_lo, base + L.off_ulo, (uint64_t)4 * h, h, kTcBN))) return rc;

  TcParams P;
  P.wc = reinterpret_cast<const float*>(base + L.off_wc);
  P.bias = reinterpret_cast<const float*>(base + L.off_bias);
  P.wh = reinterpret_cast<const float*>(base + L.off_wh);
  P.scale = reinterpret_cast<const float*>(base + L.off_scale);
  P.xv = xv; P.g = g;
  P.hout_hi = Hout_hi; P.hout_lo = Hout_lo; P.hout_f32 = H_out_f32; P.C = C; P.head_part = head_part;
  P.rows = rows; P.h = h;
  P.unit_tiles = cdiv(h, kTcUnits);
  P.k_blocks = cdiv(h, kTcBK);
  P.nprod = nprod;
  P.num_tiles = ((rows + kTcBM - 1) / kTcBM) * P.unit_tiles;
  const int stage_bytes = (nprod == 3) ? 2 * (kTcABytes + kTcBBytes) : (kTcABytes + kTcBBytes);
  P.stages = (nprod == 3) ? 4 : 8;
  const size_t smem = 1024 + (size_t)P.stages * stage_bytes + 2 * (3 * kTcBN + kTcUnits) * sizeof(float) +
                      (2 * P.stages + 4) * sizeof(uint64_t) + 16;

  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    IADMM_CUDA(cudaGetDevice(&dev));
    IADMM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long grid = P.num_tiles < num_sms ? P.num_tiles : num_sms;
  if (nprod == 3) {
    static bool attr3 = false;
    if (!attr3) {
      IADMM_CUDA(cudaFuncSetAttribute(gates_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr3 = true;
    }
    gates_tc_kernel<3><<<(unsigned)grid, kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  } else {
    static bool attr1 = false;
    if (!attr1) {
      IADMM_CUDA(cudaFuncSetAttribute(gates_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr1 = true;
    }
    gates_tc_kernel<1><<<(unsigned)grid, kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  }
  IADMM_LAUNCH_CHECK("gates_tc_kernel");
  return IADMM_OK;
}

}  // namespace iadmm
