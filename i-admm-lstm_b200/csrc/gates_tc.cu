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
#include "tc_ptx.cuh"
#include "gate_math.cuh"

#include <stdlib.h>

namespace iadmm {

constexpr int kTcBN = 256;            // gate columns per tile (UMMA N) = 64 hidden units
constexpr int kTcUnits = kTcBN / 4;
constexpr int kTcBK = 32;             // K elements per stage (64-byte swizzled rows of fp16)
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr int kTcABytes = kTcBM * kTcBK * 2;     // 8 KB
constexpr int kTcBBytes = kTcBN * kTcBK * 2;     // 16 KB
constexpr int kTcChunk = 32;                     // TMEM columns per epilogue step (8 units)

int tc_gate_tiles(int h) { return 2 * cdiv(h, kTcUnits); }   // two head partials per unit tile
size_t tc_state_bytes(long rows, int h) { return (size_t)rows * h * sizeof(__half) * 4; }

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
  __half* hout_lo;        // NPROD==2: two e4m3 arrays, [rows*h] residual then [rows*h] coarse copy
  float*  hout_f32;       // optional
  float*  gates_out;      // optional (training): gate activations [rows][4h], column 4j+g
  float*  C;              // [rows][h] in place
  float*  head_part;      // [2*unit_tiles][rows]
  long rows;
  int  h, unit_tiles, k_blocks, stages, nprod;
  long num_tiles;
  size_t q8_pitch;        // bytes per row of the packed e4m3 image (NPROD 2)
};

// The fused LSTM-cell epilogue of one tile, executed by the 8 epilogue warps of a CTA (thread = one
// accumulator lane = one coordinate row; a warp pair splits the tile's 8 column chunks of 8 hidden units).
// Split in two so the C loads are in flight while the warp waits for the accumulator.
constexpr int kChunksPerHalf = kTcBN / kTcChunk / 2;   // 4

struct EpiRow {
  long row;
  bool row_ok;
  float xr, gr;
  float c[kChunksPerHalf][8];
};

__device__ __forceinline__ void lstm_epilogue_prefetch(const TcParams& P, EpiRow& R, int quarter, int half, int lane, int ut,
                                                       long row_base) {
  R.row = row_base + quarter * 32 + lane;
  R.row_ok = R.row < P.rows;
  R.xr = R.row_ok ? __ldg(P.xv + R.row) : 0.f;
  R.gr = R.row_ok ? __ldg(P.g + R.row) : 0.f;
#pragma unroll
  for (int cc = 0; cc < kChunksPerHalf; ++cc) {
    const int unit0 = ut * kTcUnits + (half * kChunksPerHalf + cc) * 8;
    if (R.row_ok && unit0 < P.h) {
      ld_global_v8(P.C + (size_t)R.row * P.h + unit0, R.c[cc]);
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) R.c[cc][u] = 0.f;
    }
  }
}

template <int NPROD>
__device__ __forceinline__ void lstm_epilogue_tile(const TcParams& P, const EpiRow& R, const float* sp, uint32_t tmem_base, int buf,
                                                   int quarter, int half, int ut, float dequant) {
  float hp = 0.f;
  // with h % 16 == 0 two consecutive 8-unit chunks are written together: 32-byte stores of the fp16 image
  // (one sector, one request) and 16-byte stores of the e4m3 images
  const bool wide = (P.h % 16) == 0;
  uint32_t hi_st[4], lo_st[4], res_st[2], crs_st[2];
#pragma unroll
  for (int cc = 0; cc < kChunksPerHalf; ++cc) {
    const int chunk = half * kChunksPerHalf + cc;
    const int unit0 = ut * kTcUnits + chunk * 8;          // first hidden unit of this chunk
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kTcBN + chunk * kTcChunk);
    tc_ld32(taddr, v);
    tc_wait_ld();
    if (R.row_ok && unit0 < P.h) {
      const size_t o = (size_t)R.row * P.h + unit0;
      float cnew[8], hnew[8];
      const float4* w0 = reinterpret_cast<const float4*>(sp + chunk * kTcChunk);
      const float4* w1 = reinterpret_cast<const float4*>(sp + kTcBN + chunk * kTcChunk);
      const float4* bb = reinterpret_cast<const float4*>(sp + 2 * kTcBN + chunk * kTcChunk);
      const float4* wh = reinterpret_cast<const float4*>(sp + 3 * kTcBN + chunk * 8);
      const float4 wh0 = wh[0], wh1 = wh[1];
      const float whv[8] = {wh0.x, wh0.y, wh0.z, wh0.w, wh1.x, wh1.y, wh1.z, wh1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 a0 = w0[u], a1 = w1[u], ab = bb[u];
        // pre_g = xv*W0 + grad*W1 + (H@U) + b   (models/lstm.py:74-77)
        const float pi = fmaf(__uint_as_float(v[u * 4 + 0]), dequant, fmaf(R.gr, a1.x, fmaf(R.xr, a0.x, ab.x)));
        const float pf = fmaf(__uint_as_float(v[u * 4 + 1]), dequant, fmaf(R.gr, a1.y, fmaf(R.xr, a0.y, ab.y)));
        const float po = fmaf(__uint_as_float(v[u * 4 + 2]), dequant, fmaf(R.gr, a1.z, fmaf(R.xr, a0.z, ab.z)));
        const float pu = fmaf(__uint_as_float(v[u * 4 + 3]), dequant, fmaf(R.gr, a1.w, fmaf(R.xr, a0.w, ab.w)));
        const float gi = sigmoid_fast(pi);
        const float gf = sigmoid_fast(pf);
        const float go = sigmoid_fast(po);
        const float gu = tanh_fast(pu);
        const float cn = __fadd_rn(__fmul_rn(gi, gu), __fmul_rn(gf, R.c[cc][u]));  // lstm.py:78
        const float hn = __fmul_rn(go, tanh_fast(cn));                              // lstm.py:79
        cnew[u] = cn;
        hnew[u] = hn;
        hp = fmaf(hn, whv[u], hp);                                                  // lstm.py:80 (partial)
        if (P.gates_out)      // training forward: keep the activations for the hand-written backward
          *reinterpret_cast<float4*>(P.gates_out + (size_t)R.row * 4 * P.h + 4 * (size_t)(unit0 + u)) = make_float4(gi, gf, go, gu);
      }
      st_global_v8(P.C + o, cnew);
      if (P.hout_f32) st_global_v8(P.hout_f32 + o, hnew);
      uint32_t hi[4], lo[4], res[2], crs[2];
      split_hidden8<NPROD>(hnew, hi, lo, res, crs);
      uint8_t* q8row = reinterpret_cast<uint8_t*>(P.hout_lo) + (size_t)R.row * P.q8_pitch;
      if (!wide) {
        *reinterpret_cast<uint4*>(P.hout_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (NPROD == 3) *reinterpret_cast<uint4*>(P.hout_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (NPROD == 2) {
          uint8_t* q = q8row + (size_t)(unit0 >> 6) * 128 + (unit0 & 63);
          *reinterpret_cast<uint2*>(q)      = make_uint2(res[0], res[1]);
          *reinterpret_cast<uint2*>(q + 64) = make_uint2(crs[0], crs[1]);
        }
      } else if ((cc & 1) == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { hi_st[u] = hi[u]; lo_st[u] = lo[u]; }
        res_st[0] = res[0]; res_st[1] = res[1]; crs_st[0] = crs[0]; crs_st[1] = crs[1];
      } else {
        const size_t o2 = o - 8;                           // first unit of the chunk pair (multiple of 16)
        const uint32_t w8[8] = {hi_st[0], hi_st[1], hi_st[2], hi_st[3], hi[0], hi[1], hi[2], hi[3]};
        st_global_v8u(P.hout_hi + o2, w8);
        if (NPROD == 3) {
          const uint32_t l8[8] = {lo_st[0], lo_st[1], lo_st[2], lo_st[3], lo[0], lo[1], lo[2], lo[3]};
          st_global_v8u(P.hout_lo + o2, l8);
        }
        if (NPROD == 2) {
          const int u2 = unit0 - 8;
          uint8_t* q = q8row + (size_t)(u2 >> 6) * 128 + (u2 & 63);
          *reinterpret_cast<uint4*>(q)      = make_uint4(res_st[0], res_st[1], res[0], res[1]);
          *reinterpret_cast<uint4*>(q + 64) = make_uint4(crs_st[0], crs_st[1], crs[0], crs[1]);
        }
      }
    }
  }
  if (R.row_ok) P.head_part[((size_t)ut * 2 + half) * P.rows + R.row] = hp;
}

// stage this tile's W rows / bias / W_h slice (shared by all rows) into shared memory
__device__ __forceinline__ void stage_tile_params(const TcParams& P, float* sp, int et, int ut) {
  const int h4 = 4 * P.h;
  const int c = ut * kTcBN + et;
  const bool ok = c < h4;
  sp[et]             = ok ? __ldg(P.wc + c) : 0.f;
  sp[kTcBN + et]     = ok ? __ldg(P.wc + h4 + c) : 0.f;
  sp[2 * kTcBN + et] = ok ? __ldg(P.bias + c) : 0.f;
  if (et < kTcUnits) {
    const int u = ut * kTcUnits + et;
    sp[3 * kTcBN + et] = (u < P.h) ? __ldg(P.wh + u) : 0.f;
  }
}

constexpr int kParamFloats = 3 * kTcBN + kTcUnits;

// ================================================================================================
// variant A: one CTA per tile (cta_group::1), tile = 128 rows x 256 gate columns
// ================================================================================================
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
      float* sp = sparam + buf * kParamFloats;
      stage_tile_params(P, sp, et, ut);
      EpiRow R;
      lstm_epilogue_prefetch(P, R, quarter, half, lane, ut, rt * kTcBM);
      asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");
      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1);
      tc_fence_after();
      lstm_epilogue_tile<NPROD>(P, R, sp, tmem_base, buf, quarter, half, ut, dequant);
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

// ================================================================================================
// variant B: CTA pair (cta_group::2): tile = 256 rows x 256 gate columns over two SMs.  Each CTA loads its
// own 128 rows of H and HALF of the U tile; the leader issues M=256 MMAs that read both halves, so the
// shared-memory traffic per MMA drops from 12 KB to 8 KB per SM and the TMA fill from 48 to 32 KB per
// K-block -- the single-CTA form is shared-memory-bandwidth bound (ncu: tensor pipe 71 % active).
// Barrier protocol: `full` lives in the leader and counts the TMA bytes of BOTH CTAs; `empty` and
// `tmem_full` are signalled in both CTAs by multicast tcgen05.commit; `tmem_empty` lives in the leader and
// collects the epilogue warps of both CTAs (the peer's arrive remotely through the cluster window).
// ================================================================================================
// The pair kernel moves 64 K-elements per stage: fp16 operand rows are full 128-byte lines (128B swizzle), the
// e4m3 operand rows 64 bytes (64B swizzle); 3 stages of 64 KB.
constexpr int kPairBK = 64;
constexpr int kPairABytes = kTcBM * kPairBK * 2;          // 16 KB: this CTA's 128 rows of H (fp16)
constexpr int kPairBBytes = (kTcBN / 2) * kPairBK * 2;    // 16 KB: this CTA's half of the U tile (fp16)
constexpr int kPairBBoxRows = 64;                         // U tiles are fetched in 64-row TMA boxes

// CL = cluster size: 2 = one CTA pair; 4 = two pairs working on the SAME unit tile for two different row tiles,
// which lets each 64-row piece of the U tile be fetched from L2 once and multicast to both pairs (the kernel is
// L2->SM fill bound, profiles/README.md): U-tile L2 reads halve, at the price of 132 instead of 148 usable SMs.
template <int NPROD, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kTcThreads, 1)
gates_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     const TcParams P) {
  // NPROD: 3 = fp16 hi/lo split (3 MMAs), 1 = single fp16 MMA, 2 = fp16 MMA + two e4m3 correction MMAs.
  // Stage layout: A_hi16 | A_lo | B_hi16 | B_lo, 16 KB each, every row 128 bytes (128B swizzle).  NPROD 3: lo = fp16
  // residual.  NPROD 2: lo = packed e4m3 image, bytes [0,64) of a row = residual, [64,128) = coarse copy.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kStageBytes = (NPROD == 1) ? (kPairABytes + kPairBBytes) : 2 * (kPairABytes + kPairBBytes);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = P.stages;
  float* sparam = reinterpret_cast<float*>(smem + (size_t)stages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sparam + 2 * kParamFloats);
  uint64_t* full_bar = bars;                     // [stages]   (leader's copy is the live one)
  uint64_t* empty_bar = bars + stages;           // [stages]   both CTAs
  uint64_t* tfull_bar = bars + 2 * stages;       // [2]        both CTAs
  uint64_t* tempty_bar = bars + 2 * stages + 2;  // [2]        leader's copy is the live one
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank & 1u) == 0;
  const uint32_t leader_rank = rank & ~1u;                 // leader of this CTA's pair
  const uint32_t pair_in_cluster = rank >> 1;
  constexpr int kPairsPerCluster = CL / 2;
  const uint16_t all_mask = (uint16_t)((1u << CL) - 1u);
  const uint16_t pair_mask = (uint16_t)(3u << leader_rank);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_b_hi);
    if (NPROD != 1) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    // a stage is refilled only after the MMAs of EVERY pair of the cluster have consumed it (multicast writes
    // land in the sibling pair's shared memory too)
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), kPairsPerCluster); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 2 * kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int unit_tiles = P.unit_tiles;
  const int h4 = 4 * P.h;
  // work unit of a cluster = (unit tile, kPairsPerCluster consecutive 256-row tiles); `pair`/`num_pairs` count clusters
  const long pair = blockIdx.x / CL, num_pairs = gridDim.x / CL;

  if (warp == 0) {
    // ===================== TMA producer (every CTA) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      constexpr uint32_t kATotal = (NPROD == 1) ? kPairABytes : 2 * kPairABytes;          // bytes of H operands per CTA and stage
      constexpr uint32_t kBBoxTotal = ((NPROD == 1) ? 1 : 2) * kPairBBoxRows * kPairBK * 2; // bytes of one 64-row box of every U operand
      for (long tile = pair; tile < P.num_tiles; tile += num_pairs) {
        const int  ut = (int)(tile % unit_tiles);
        const long rt = tile / unit_tiles;
        const int n_cols = min(kTcBN, h4 - ut * kTcBN);
        const int row0 = (int)((rt * kPairsPerCluster + pair_in_cluster) * (2 * kTcBM)) + (int)(rank & 1u) * kTcBM;   // this CTA's 128 rows of H
        const int col0 = ut * kTcBN + (int)(rank & 1u) * (n_cols / 2);        // first row of this CTA's half of the U tile
        // U boxes are 64 rows.  Full tiles in a 4-CTA cluster: each CTA fetches ONE 64-row piece of its half and
        // multicasts it to the CTA with the same role in the sibling pair.  Otherwise the CTA loads its half itself.
        const bool mcast = (CL == 4) && (n_cols == kTcBN);
        const int  b_boxes = mcast ? 2 : (n_cols / 2 + kPairBBoxRows - 1) / kPairBBoxRows;   // boxes landing in this CTA
        const uint16_t mc_mask = (uint16_t)((1u << rank) | (1u << (rank ^ 2u)));
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb_local = smem_u32(&full_bar[stage]);
          if (leader) mbar_expect_tx(fb_local, 2u * (kATotal + (uint32_t)b_boxes * kBBoxTotal));   // both CTAs of the pair
          const uint32_t fb = map_to_cta(fb_local, leader_rank);
          uint8_t* sbase = smem + (size_t)stage * kStageBytes;
          const int k0 = kb * kPairBK;
          auto load_u = [&](const CUtensorMap* map, uint32_t region, int kc) {
            constexpr uint32_t kBoxBytes = kPairBBoxRows * 128;
            if (mcast) {
              const uint32_t q = pair_in_cluster;
              tma_load_2d_pair_mc(region + q * kBoxBytes, map, fb_local, kc, col0 + (int)q * kPairBBoxRows, mc_mask);
            } else {
              for (int bx = 0; bx < b_boxes; ++bx)
                tma_load_2d_pair(region + (uint32_t)bx * kBoxBytes, map, fb, kc, col0 + bx * kPairBBoxRows);
            }
          };
          const uint32_t sb = smem_u32(sbase);
          tma_load_2d_pair(sb, &map_a_hi, fb, k0, row0);
          if (NPROD != 1) {
            // the packed e4m3 tensors are addressed in bytes: K block kb starts at byte kb*128 of a row
            const int k0_lo = (NPROD == 2) ? kb * 128 : k0;
            tma_load_2d_pair(sb + kPairABytes, &map_a_lo, fb, k0_lo, row0);
            load_u(&map_b_hi, sb + 2 * kPairABytes, k0);
            load_u(&map_b_lo, sb + 2 * kPairABytes + kPairBBytes, k0_lo);
          } else {
            load_u(&map_b_hi, sb + kPairABytes, k0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      long it = 0;
      for (long tile = pair; tile < P.num_tiles; tile += num_pairs, ++it) {
        const int ut = (int)(tile % unit_tiles);
        const int n_cols = min(kTcBN, h4 - ut * kTcBN);
        const uint32_t idesc = make_idesc_f16(n_cols, 2 * kTcBM);
        const int buf = (int)(it & 1);
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&tempty_bar[buf]), (use & 1) ^ 1);     // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcBN);
        uint32_t acc = 0;
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * kStageBytes);
          const int k_len = min(kPairBK, P.h - kb * kPairBK);
          const int k_steps = (k_len + kTcUK - 1) / kTcUK;
          for (int ks = 0; ks < k_steps; ++ks) {
            const uint32_t koff = (uint32_t)(ks * kTcUK * 2);      // bytes inside the 128-byte swizzled row
            if (NPROD == 3) {
              const uint64_t a_hi = make_smem_desc_sw128(sbase + koff);
              const uint64_t a_lo = make_smem_desc_sw128(sbase + kPairABytes + koff);
              const uint64_t b_hi = make_smem_desc_sw128(sbase + 2 * kPairABytes + koff);
              const uint64_t b_lo = make_smem_desc_sw128(sbase + 2 * kPairABytes + kPairBBytes + koff);
              tc_mma_f16_pair(d_tmem, a_lo, b_hi, idesc, acc); acc = 1;
              tc_mma_f16_pair(d_tmem, a_hi, b_lo, idesc, 1);
              tc_mma_f16_pair(d_tmem, a_hi, b_hi, idesc, 1);
            } else if (NPROD == 2) {
              if ((ks & 1) == 0) {
                // one e4m3 MMA covers 32 K-elements = two fp16 K-steps; residual in bytes [0,64) of the row, coarse in [64,128)
                const uint32_t koff8 = (uint32_t)(ks * kTcUK);
                const uint32_t a8 = sbase + kPairABytes, b8 = sbase + 2 * kPairABytes + kPairBBytes;
                const uint64_t a_res = make_smem_desc_sw128(a8 + koff8);
                const uint64_t a_crs = make_smem_desc_sw128(a8 + 64 + koff8);
                const uint64_t b_res = make_smem_desc_sw128(b8 + koff8);
                const uint64_t b_crs = make_smem_desc_sw128(b8 + 64 + koff8);
                tc_mma_f8_pair(d_tmem, a_res, b_crs, idesc, acc); acc = 1;       // (H - fp16(H)) * U
                tc_mma_f8_pair(d_tmem, a_crs, b_res, idesc, 1);                  // H * (U - fp16(U))
              }
              const uint64_t a_hi = make_smem_desc_sw128(sbase + koff);
              const uint64_t b_hi = make_smem_desc_sw128(sbase + 2 * kPairABytes + koff);
              tc_mma_f16_pair(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              const uint64_t a_hi = make_smem_desc_sw128(sbase + koff);
              const uint64_t b_hi = make_smem_desc_sw128(sbase + kPairABytes + koff);
              tc_mma_f16_pair(d_tmem, a_hi, b_hi, idesc, acc); acc = 1;
            }
          }
          tc_commit_pair(smem_u32(&empty_bar[stage]), all_mask);   // one of the kPairsPerCluster arrivals, in every CTA
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(smem_u32(&tfull_bar[buf]), pair_mask);      // accumulators complete in both CTAs of this pair
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = (ew >= 4) ? 1 : 0;
    const int et = threadIdx.x - 64;
    const float dequant = P.scale[1];
    long it = 0;
    for (long tile = pair; tile < P.num_tiles; tile += num_pairs, ++it) {
      const int  ut = (int)(tile % unit_tiles);
      const long rt = tile / unit_tiles;
      const int buf = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      float* sp = sparam + buf * kParamFloats;
      stage_tile_params(P, sp, et, ut);
      EpiRow R;
      lstm_epilogue_prefetch(P, R, quarter, half, lane, ut,
                             (rt * kPairsPerCluster + pair_in_cluster) * (2 * kTcBM) + (long)(rank & 1u) * kTcBM);
      asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");
      mbar_wait(smem_u32(&tfull_bar[buf]), use & 1);
      tc_fence_after();
      lstm_epilogue_tile<NPROD>(P, R, sp, tmem_base, buf, quarter, half, ut, dequant);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[buf]), leader_rank));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
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

// 2D row-major [rows_total][h] tensor of fp16 (elem_bytes 2, 64B swizzle) or e4m3 bytes (elem_bytes 1, 32B
// swizzle), box = [box_rows][32 elements], OOB reads as zero
int make_map(CUtensorMap* map, const void* base, uint64_t rows_total, int h, int box_rows, int elem_bytes, int box_k) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) IADMM_FAIL(IADMM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)h, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)h * (cuuint64_t)elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  const int row_bytes = box_k * elem_bytes;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IADMM_FAIL(IADMM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu h=%d)", (int)r,
                                    (unsigned long long)rows_total, h);
  return IADMM_OK;
}

static bool use_quads() {
  static int v = -1;
  if (v < 0) {
    // 4-CTA multicast clusters are bit-identical to plain pairs but measured 3-4 % slower on B200 (the kernel is
    // bound by the per-SM request port, not by L2 reads, and only 132 SMs host 4-CTA clusters): opt-in only.
    const char* e = getenv("IADMM_TC_QUAD");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IADMM_TC_CTA_PAIR");      // development switch: 0 = single-CTA tiles
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <typename KernelT>
static int set_smem_attr(KernelT kernel, bool* done) {
  if (!*done) {
    IADMM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    *done = true;
  }
  return IADMM_OK;
}

int launch_gates_tc(const void* packed, const WeightLayout& L, const float* xv, const float* g, const __half* Hin_hi,
                    const __half* Hin_lo, __half* Hout_hi, __half* Hout_lo, float* H_out_f32, float* C,
                    float* head_part, long rows, int h, int nprod, cudaStream_t st, float* gates_out) {
  if (h % 8 != 0) IADMM_FAIL(IADMM_EMODE, "tensor-core gate path needs hidden_dim %% 8 == 0");
  if (rows > 0x7fffffffL - 2 * kTcBM) IADMM_FAIL(IADMM_ESHAPE, "too many rows for one launch");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    IADMM_CUDA(cudaGetDevice(&dev));
    IADMM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (const char* e = getenv("IADMM_TC_MAX_SMS")) {      // development switch: restrict the persistent grid
    const int v = atoi(e);
    if (v >= 2 && v < num_sms) num_sms = v;
  }
  const bool pair = use_cta_pairs() && num_sms >= 2;
  const char* base = static_cast<const char*>(packed);
  if (nprod == 2 && !pair) IADMM_FAIL(IADMM_EMODE, "the fp16+fp8 gate mode runs on CTA pairs only");
  if (nprod == 2 && h % 16 != 0) IADMM_FAIL(IADMM_EMODE, "the fp16+fp8 gate mode needs hidden_dim %% 16 == 0");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  const int b_box_rows = pair ? kPairBBoxRows : kTcBN;
  const int bk = pair ? kPairBK : kTcBK;
  if ((rc = make_map(&ma_hi, Hin_hi, (uint64_t)rows, h, kTcBM, 2, bk))) return rc;
  if ((rc = make_map(&mb_hi, base + L.off_uhi, (uint64_t)4 * h, h, b_box_rows, 2, bk))) return rc;
  if (nprod == 2) {
    // packed e4m3 images: byte tensors [rows][q8_pitch], one 128-byte box row per 64-wide K block
    const int pitch = (int)q8_pitch(h);
    if ((rc = make_map(&ma_lo, Hin_lo, (uint64_t)rows, pitch, kTcBM, 1, 128))) return rc;
    if ((rc = make_map(&mb_lo, base + L.off_uq8, (uint64_t)4 * h, pitch, b_box_rows, 1, 128))) return rc;
  } else {
    if ((rc = make_map(&ma_lo, Hin_lo, (uint64_t)rows, h, kTcBM, 2, bk))) return rc;
    if ((rc = make_map(&mb_lo, base + L.off_ulo, (uint64_t)4 * h, h, b_box_rows, 2, bk))) return rc;
  }

  TcParams P;
  P.wc = reinterpret_cast<const float*>(base + L.off_wc);
  P.bias = reinterpret_cast<const float*>(base + L.off_bias);
  P.wh = reinterpret_cast<const float*>(base + L.off_wh);
  P.scale = reinterpret_cast<const float*>(base + L.off_scale);
  P.xv = xv; P.g = g;
  P.hout_hi = Hout_hi; P.hout_lo = Hout_lo; P.hout_f32 = H_out_f32; P.C = C; P.head_part = head_part;
  P.gates_out = gates_out;
  P.rows = rows; P.h = h; P.q8_pitch = q8_pitch(h);
  P.unit_tiles = cdiv(h, kTcUnits);
  P.k_blocks = cdiv(h, bk);
  P.nprod = nprod;
  // cluster size: 4 (two pairs sharing each U tile through TMA multicast) when there is enough work, else 2
  static int max_quads = -1;
  int cl = 1;
  if (pair) {
    cl = 2;
    const long row_tiles = (rows + 2 * kTcBM - 1) / (2 * kTcBM);
    if (use_quads() && row_tiles >= 2 && num_sms >= 4) cl = 4;
  }
  const int tile_rows = (pair ? 2 * kTcBM : kTcBM) * (cl == 4 ? 2 : 1);
  P.num_tiles = ((rows + tile_rows - 1) / tile_rows) * P.unit_tiles;
  const int a_bytes = pair ? kPairABytes : kTcABytes;
  const int b_bytes = pair ? kPairBBytes : kTcBBytes;
  const int stage_bytes = (nprod == 1) ? (a_bytes + b_bytes) : 2 * (a_bytes + b_bytes);
  P.stages = (192 * 1024) / stage_bytes;                 // 4 / 8 (single CTA, 32-wide K), 3 / 6 (pair, 64-wide K)
  const size_t smem = 1024 + (size_t)P.stages * stage_bytes + 2 * kParamFloats * sizeof(float) +
                      (2 * P.stages + 4) * sizeof(uint64_t) + 16;

  if (pair) {
    auto launch = [&](auto kernel, bool* attr_done, int cluster) -> int {
      int rc2;
      if ((rc2 = set_smem_attr(kernel, attr_done))) return rc2;
      long clusters = num_sms / cluster;
      if (cluster == 4) {
        if (max_quads < 0) {      // co-resident 4-CTA clusters (GPC granularity: 33 on a 148-SM B200)
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3((unsigned)(num_sms / 4 * 4)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          int n = 0;
          if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) n = num_sms / 4 - 4;
          max_quads = n;
        }
        clusters = max_quads < num_sms / 4 ? max_quads : num_sms / 4;
      }
      if (P.num_tiles < clusters) clusters = P.num_tiles;
      kernel<<<(unsigned)(cluster * clusters), kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
      return IADMM_OK;
    };
    static bool a34 = false, a24 = false, a14 = false, a32 = false, a22 = false, a12 = false;
    if (cl == 4) {
      if (nprod == 3)      rc = launch(gates_tc_pair_kernel<3, 4>, &a34, 4);
      else if (nprod == 2) rc = launch(gates_tc_pair_kernel<2, 4>, &a24, 4);
      else                 rc = launch(gates_tc_pair_kernel<1, 4>, &a14, 4);
    } else {
      if (nprod == 3)      rc = launch(gates_tc_pair_kernel<3, 2>, &a32, 2);
      else if (nprod == 2) rc = launch(gates_tc_pair_kernel<2, 2>, &a22, 2);
      else                 rc = launch(gates_tc_pair_kernel<1, 2>, &a12, 2);
    }
    if (rc) return rc;
    IADMM_LAUNCH_CHECK("gates_tc_pair_kernel");
    return IADMM_OK;
  }
  const long grid = P.num_tiles < num_sms ? P.num_tiles : num_sms;
  static bool a3 = false, a1 = false;
  if (nprod == 3) {
    if ((rc = set_smem_attr(gates_tc_kernel<3>, &a3))) return rc;
    gates_tc_kernel<3><<<(unsigned)grid, kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  } else {
    if ((rc = set_smem_attr(gates_tc_kernel<1>, &a1))) return rc;
    gates_tc_kernel<1><<<(unsigned)grid, kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  }
  IADMM_LAUNCH_CHECK("gates_tc_kernel");
  return IADMM_OK;
}

}  // namespace iadmm
