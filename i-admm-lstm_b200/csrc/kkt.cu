// KKT passes of one I-ADMM-LSTM iteration, never forming the KKT matrix.
//
// Reference: models/lstm.py:67-72 builds K = [[Q+sigma I, A0^T],[A0, -diag(1/rho)]] (16 MB per instance
// at n=m=1000, every iteration) and evaluates g = K^T (K xv - rhs) with two dense bmm's.  Here the same
// expression is evaluated block-wise from Q and A0 directly:
//     pass 1   w1 = Q x~ + sigma x~ + A0^T v - (sigma x - p)        w2 = A0 x~ - v/rho - (z - y/rho)
//     pass 2   g1 = Q^T w1 + sigma w1 + A0^T w2                     g2 = A0 w1 - w2/rho
// Each pass streams Q and A0 from HBM exactly once (the two passes are data dependent, so twice per
// iteration is the algorithmic minimum: 8(n^2+mn) bytes).  Pass 1 carries the previous iterate (x, y)
// as extra right-hand sides, so Q x, A0 x and A0^T y for primal_dual_loss (utils.py:68-71) come out of
// the same read and the reference's third pass per iteration disappears.
//
// Work decomposition: one CTA streams R rows of one matrix of one instance.  Its 8 warps own disjoint
// 128-column slabs (lane = one 128-bit column group), so "column" products (M^T s, accumulated over the
// CTA's rows) stay in registers with no cross-warp traffic, and "row" products (M r) are reduced with a
// transpose-reduce over the warp (2 shuffles per 128-bit load) and a small shared-memory sum over warps.
// Column partials of different row chunks are summed in a fixed order by the combine kernels, so results
// are bit-reproducible run to run.  HBM-bound by design: 8 independent 128-bit loads in flight per lane.
#include "common.cuh"
#include "tc_ptx.cuh"      // mbarrier / cp.async.bulk wrappers (the bulk-copy-staged variant of the dense pass)

#include <string.h>

namespace iadmm {

constexpr int kKktThreads = 256;
constexpr int kKktWarps   = kKktThreads / 32;
constexpr int kRowUnroll  = 8;
constexpr int kSlabCols   = 128;                    // columns per warp per column chunk
constexpr int kChunkCols  = kSlabCols * kKktWarps;  // 1024

static KktDims kkt_dims_with_rows(int B, int n, int m, int num_ineq, int R) {
  KktDims d;
  d.B = B; d.n = n; d.m = m; d.num_ineq = num_ineq;
  d.rows_per_chunk = R;
  d.chunks_q = cdiv(n, R);
  d.chunks_a = cdiv(m, R);
  d.sum_q = d.chunks_q; d.sum_a = d.chunks_a;
  return d;
}

KktDims make_kkt_dims(int B, int n, int m, int num_ineq) {
  // Rows per CTA depend on the problem size only, never on the batch: the grouping of the column partial sums
  // (and with it every rounding) is then identical however a batch is sharded over calls or GPUs, so
  // "concatenation of shards == whole batch" holds bit for bit.  64 rows keep the column-partial traffic
  // (4n bytes per chunk and product) at ~3 % of the 4nR bytes streamed and give B*(n+m)/64 CTAs.
  return kkt_dims_with_rows(B, n, m, num_ineq, ((n > m ? n : m) >= 64) ? 64 : 32);
}

KktDims make_kkt_dims_train(int B, int n, int m, int num_ineq) {
  // Training runs at a few instances per GPU (BASELINE config 3: batch 2): 64-row chunks would give B*(n+m)/64 = 62
  // CTAs for 148 SMs.  The training entry points therefore shrink the chunk until one wave (148 SMs x 4 CTAs) is
  // filled -- at the price of more column-partial traffic, and of results that depend on the batch size in the last
  // bits (gradients are averaged over ranks anyway).
  const int base = make_kkt_dims(B, n, m, num_ineq).rows_per_chunk;
  int R = base;
  while (R > 8 && (long)(cdiv(n, R) + cdiv(m, R)) * B < 592) R >>= 1;
  return kkt_dims_with_rows(B, n, m, num_ineq, R);
}

size_t kkt_scratch_floats(const KktDims& d) {
  const size_t B = d.B, n = d.n, m = d.m;
  return 3 * B * n + 5 * B * m + B * (size_t)d.chunks_a * 2 * n + B * (size_t)d.chunks_q * n + 2 * B * (n + m) + 64;
}

void kkt_scratch_carve(const KktDims& d, float* base, KktScratch* s) {
  const size_t B = d.B, n = d.n, m = d.m;
  auto take = [&](size_t cnt) { float* p = base; base += (cnt + 3) / 4 * 4; return p; };
  s->qxt = take(B * n);  s->qx = take(B * n);
  s->axt = take(B * m);  s->ax = take(B * m);  s->aw1 = take(B * m);
  s->part_a = take(B * d.chunks_a * 2 * n);
  s->part_q = take(B * d.chunks_q * n);
  s->w = take(B * (n + m));
  s->g = take(B * (n + m));
  s->x_old = take(B * n); s->y_old = take(B * m); s->z_old = take(B * m);
}

// ------------------------------------------------------------------------------------------------
// the streaming tile pass
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__device__ __forceinline__ float4 load_cols4(const float* __restrict__ rowp, int col, int n) {
  if (VEC) return ldg_stream4(rowp + col);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col + 0 < n) r.x = ldg_stream1(rowp + col + 0);
  if (col + 1 < n) r.y = ldg_stream1(rowp + col + 1);
  if (col + 2 < n) r.z = ldg_stream1(rowp + col + 2);
  if (col + 3 < n) r.w = ldg_stream1(rowp + col + 3);
  return r;
}
template <bool VEC>
__device__ __forceinline__ float4 load_vec4(const float* __restrict__ v, int col, int n) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  // the per-instance vectors start at b*(n+m) floats, which is 16-byte aligned only when (n+m) % 4 == 0
  if (VEC && (reinterpret_cast<uintptr_t>(v) & 15u) == 0) {
    if (col < n) r = __ldg(reinterpret_cast<const float4*>(v + col));
    return r;
  }
  if (col + 0 < n) r.x = __ldg(v + col + 0);
  if (col + 1 < n) r.y = __ldg(v + col + 1);
  if (col + 2 < n) r.z = __ldg(v + col + 2);
  if (col + 3 < n) r.w = __ldg(v + col + 3);
  return r;
}
template <bool VEC>
__device__ __forceinline__ void store_cols4(float* __restrict__ v, int col, int n, float4 a) {
  if (VEC) {
    if (col < n) *reinterpret_cast<float4*>(v + col) = a;
    return;
  }
  if (col + 0 < n) v[col + 0] = a.x;
  if (col + 1 < n) v[col + 1] = a.y;
  if (col + 2 < n) v[col + 2] = a.z;
  if (col + 3 < n) v[col + 3] = a.w;
}

// Sparse operand ("bitmap slabs", sparse.cu): per row and 128-column slab a 128-bit occupancy mask (bit b of word w <-> column
// 128*slab + 32*w + b), the offset of the slab's first stored value, and the non-zero values of the instance in row-major
// order.  A lane of the streaming pass owns 4 consecutive columns = one nibble of the mask, so the dense kernel's work
// decomposition, accumulation order and therefore every rounding carry over unchanged: the sparse loader returns exactly the
// values the dense loader would (zeros where the mask is clear) while reading 4*nnz + 20*rows*ceil(n/128) bytes.
__device__ __forceinline__ float4 expand_cols4_sparse(const float* __restrict__ vals, const uint4 mk, uint32_t base, int lane) {
  const int w = lane >> 3, sh = (lane & 7) * 4;
  const uint32_t mine = (w == 0) ? mk.x : (w == 1) ? mk.y : (w == 2) ? mk.z : mk.w;
  const uint32_t before = ((w > 0) ? __popc(mk.x) : 0) + ((w > 1) ? __popc(mk.y) : 0) + ((w > 2) ? __popc(mk.z) : 0) +
                          __popc(mine & ((1u << sh) - 1u));
  const uint32_t nib = (mine >> sh) & 15u;
  const float* p = vals + base + before;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nib & 1u) r.x = __ldg(p);
  p += (nib & 1u);
  if (nib & 2u) r.y = __ldg(p);
  p += ((nib >> 1) & 1u);
  if (nib & 4u) r.z = __ldg(p);
  p += ((nib >> 2) & 1u);
  if (nib & 8u) r.w = __ldg(p);
  return r;
}

struct TileArgs {
  const float* mat;        // first element of the instance's matrix [rows_total, n]
  SpMat sp;                // SP: the matrix in bitmap-slab form instead
  size_t inst;             // SP: instance index
  const unsigned long long* blk;   // block-skip form: this instance's block-occupancy words [ceil(rows/8)]
  int rows_total, n, r0, R;
  const float* rrhs[2];    // NR vectors of length n      (row products  M r)
  float*       rout[2];    // NR outputs of length rows_total
  const float* crhs[2];    // NC vectors of length rows_total (column products M^T s)
  float*       cpart[2];   // NC outputs of length n: this chunk's partial of M^T s
};

// smem: rowscal[2][R] | rowacc[R*2] | rowpart[warps][R*2]
template <int NR, int NC, bool VEC>
__device__ __forceinline__ void tile_pass(const TileArgs& a, float* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R, n = a.n;
  float* rowscal = smem;
  float* rowacc  = smem + 2 * R;
  float* rowpart = smem + 4 * R;

#pragma unroll
  for (int c = 0; c < NC; ++c)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      rowscal[c * R + r] = (row < a.rows_total) ? __ldg(a.crhs[c] + row) : 0.f;
    }
  for (int i = tid; i < R * NR; i += kKktThreads) rowacc[i] = 0.f;
  __syncthreads();

  const int nchunk = (n + kChunkCols - 1) / kChunkCols;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kChunkCols + warp * kSlabCols + lane * 4;
    const bool active = col < n;
    float4 rv[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < NR; ++k) rv[k] = load_vec4<VEC>(a.rrhs[k], col, n);
    float4 cacc[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) cacc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int rg = 0; rg < R; rg += kRowUnroll) {
      float4 v[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) {
        const int row = a.r0 + rg + u;
        v[u] = (active && row < a.rows_total) ? load_cols4<VEC>(a.mat + (size_t)row * n, col, n)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (NC > 0) {
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float s = rowscal[c * R + rg + u];
            cacc[c].x = fmaf(v[u].x, s, cacc[c].x);
            cacc[c].y = fmaf(v[u].y, s, cacc[c].y);
            cacc[c].z = fmaf(v[u].z, s, cacc[c].z);
            cacc[c].w = fmaf(v[u].w, s, cacc[c].w);
          }
        }
      }
      if (NR > 0) {
        float rp[kRowUnroll * (NR > 0 ? NR : 1)];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int k = 0; k < NR; ++k) {
            float t = v[u].x * rv[k].x;
            t = fmaf(v[u].y, rv[k].y, t);
            t = fmaf(v[u].z, rv[k].z, t);
            t = fmaf(v[u].w, rv[k].w, t);
            rp[u * NR + k] = t;
          }
        }
        constexpr int NV = kRowUnroll * (NR > 0 ? NR : 1);
        const float tot = warp_transpose_reduce<NV>(rp, lane);
        constexpr int kGroup = 32 / NV;              // lanes holding the same value
        if ((lane % kGroup) == 0) {
          const int idx = lane / kGroup;             // = u*NR + k
          rowpart[warp * (R * NR) + rg * NR + idx] = tot;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) store_cols4<VEC>(a.cpart[c], col, n, cacc[c]);

    if (NR > 0) {
      __syncthreads();
      for (int i = tid; i < R * NR; i += kKktThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kKktWarps; ++w) s += rowpart[w * (R * NR) + i];
        rowacc[i] += s;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < NR; ++k)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      if (row < a.rows_total) a.rout[k][row] = rowacc[r * NR + k];
    }
}

// The same dense pass with the matrix bytes moved by the copy engine instead of by load instructions (north_star: "stream from
// HBM once per iteration, through TMA into shared memory, with warp-shuffle reductions").  A ninth warp issues, per 8-row group and
// 1024-column chunk, eight `cp.async.bulk` copies of one row segment each (<= 4 KB, whole 16-byte lines) into a ring of
// kTmaStages x 32 KB stages guarded by full / empty mbarriers; the eight consumer warps read their 128-column slab of a stage
// into the SAME registers the load-instruction version fills, release the stage and do the same arithmetic in the same order:
// results are bit-identical (all-zero padding row groups of a ragged last chunk are skipped; they add exact zeros).  Row-major
// matrices with n % 4 == 0 only (16-byte aligned segments).  smem: tile_smem_bytes(R) | stages | barriers.
constexpr int kTmaStages     = 3;
constexpr int kTmaStageBytes = kRowUnroll * kChunkCols * 4;      // 32 KB
constexpr int kKktTmaThreads = kKktThreads + 32;

__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kKktThreads) : "memory"); }

static size_t tile_smem_bytes_tma(int R) {
  return align_up((size_t)(4 * R + kKktWarps * R * 2) * sizeof(float), 128) + 128 + (size_t)kTmaStages * kTmaStageBytes +
         2 * kTmaStages * sizeof(uint64_t);
}

template <int NR, int NC>
__device__ __forceinline__ void tile_pass_tma(const TileArgs& a, float* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R, n = a.n;
  float* rowscal = smem;
  float* rowacc  = smem + 2 * R;
  float* rowpart = smem + 4 * R;
  // the stages start at the next 128-byte boundary behind the row arrays (the launch reserves the slack)
  const uint32_t nominal = smem_u32(smem) + (uint32_t)(4 * R + kKktWarps * R * 2) * 4u;
  unsigned char* stages = reinterpret_cast<unsigned char*>(smem) + (((nominal + 127u) & ~127u) - smem_u32(smem));
  const uint32_t stage0 = smem_u32(stages);
  const uint32_t full0  = smem_u32(stages + (size_t)kTmaStages * kTmaStageBytes);
  const uint32_t empty0 = full0 + 8 * kTmaStages;
  const int nchunk = (n + kChunkCols - 1) / kChunkCols;
  const int rows_valid = (a.rows_total - a.r0 < R) ? a.rows_total - a.r0 : R;

  pdl_trigger();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kKktWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp < kKktWarps) {
    // PDL: the vectors come from the kernel before this one; the producer warp below streams the (constant) matrix meanwhile
    pdl_wait();
#pragma unroll
    for (int c = 0; c < NC; ++c)
      for (int r = tid; r < R; r += kKktThreads) {
        const int row = a.r0 + r;
        rowscal[c * R + r] = (row < a.rows_total) ? __ldg(a.crhs[c] + row) : 0.f;
      }
    for (int i = tid; i < R * NR; i += kKktThreads) rowacc[i] = 0.f;
    // (row groups beyond the valid rows are never computed: their row partials must still be finite zeros for the chunk sums)
    for (int i = tid; i < kKktWarps * R * (NR > 0 ? NR : 1); i += kKktThreads) rowpart[i] = 0.f;
  }
  __syncthreads();

  if (warp == kKktWarps) {
    // ---- producer: one thread feeds the ring ----------------------------------------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int cc = 0; cc < nchunk; ++cc) {
        const int col0 = cc * kChunkCols;
        const uint32_t seg = (uint32_t)(((n - col0 < kChunkCols) ? n - col0 : kChunkCols) * 4);
        for (int rg = 0; rg < rows_valid; rg += kRowUnroll, ++it) {
          const int st = it % kTmaStages;
          if (it >= kTmaStages) mbar_wait(empty0 + 8 * st, ((it / kTmaStages) & 1) ^ 1);
          const int nrow = (rows_valid - rg < kRowUnroll) ? rows_valid - rg : kRowUnroll;
          mbar_expect_tx(full0 + 8 * st, seg * (uint32_t)nrow);
          for (int u = 0; u < nrow; ++u)
            bulk_load(stage0 + (uint32_t)st * kTmaStageBytes + (uint32_t)u * (kChunkCols * 4),
                      a.mat + (size_t)(a.r0 + rg + u) * n + col0, seg, full0 + 8 * st);
        }
      }
    }
    return;
  }

  // ---- consumers: tile_pass with the loads taken from the ring ------------------------------------------------------
  int it = 0;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kChunkCols + warp * kSlabCols + lane * 4;
    const bool active = col < n;
    float4 rv[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < NR; ++k) rv[k] = load_vec4<true>(a.rrhs[k], col, n);
    float4 cacc[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) cacc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int rg = 0; rg < rows_valid; rg += kRowUnroll, ++it) {
      const int st = it % kTmaStages;
      mbar_wait(full0 + 8 * st, (it / kTmaStages) & 1);
      const float4* src = reinterpret_cast<const float4*>(stages + (size_t)st * kTmaStageBytes) + warp * (kSlabCols / 4) + lane;
      float4 v[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u)
        v[u] = (active && rg + u < rows_valid) ? src[u * (kChunkCols / 4)] : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * st);          // this warp has read its slab of the stage
      if (NC > 0) {
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float sc = rowscal[c * R + rg + u];
            cacc[c].x = fmaf(v[u].x, sc, cacc[c].x);
            cacc[c].y = fmaf(v[u].y, sc, cacc[c].y);
            cacc[c].z = fmaf(v[u].z, sc, cacc[c].z);
            cacc[c].w = fmaf(v[u].w, sc, cacc[c].w);
          }
        }
      }
      if (NR > 0) {
        float rp[kRowUnroll * (NR > 0 ? NR : 1)];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int k = 0; k < NR; ++k) {
            float t = v[u].x * rv[k].x;
            t = fmaf(v[u].y, rv[k].y, t);
            t = fmaf(v[u].z, rv[k].z, t);
            t = fmaf(v[u].w, rv[k].w, t);
            rp[u * NR + k] = t;
          }
        }
        constexpr int NV = kRowUnroll * (NR > 0 ? NR : 1);
        const float tot = warp_transpose_reduce<NV>(rp, lane);
        constexpr int kGroup = 32 / NV;
        if ((lane % kGroup) == 0) {
          const int idx = lane / kGroup;
          rowpart[warp * (R * NR) + rg * NR + idx] = tot;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) store_cols4<true>(a.cpart[c], col, n, cacc[c]);

    if (NR > 0) {
      consumers_sync();
      for (int i = tid; i < R * NR; i += kKktThreads) {
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < kKktWarps; ++w) sacc += rowpart[w * (R * NR) + i];
        rowacc[i] += sacc;
      }
      consumers_sync();
    }
  }
#pragma unroll
  for (int k = 0; k < NR; ++k)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      if (row < a.rows_total) a.rout[k][row] = rowacc[r * NR + k];
    }
}

// The dense pass with block skipping (structured sparsity): `a.blk` holds one bit per 8-row x 128-column block of the
// instance's matrix (bit s of word g <-> rows 8g..8g+7, columns 128s..128s+127 contain a non-zero).  A warp's unit of work IS
// such a block, so a clear bit skips its 4 KB of loads and its products with a warp-uniform branch and nothing else changes:
// no decode cost, bit-identical results (the skipped entries are exact zeros), HBM bytes = the non-empty blocks.  Diagonal Q
// (QP family), identity blocks (SVM) and banded / block-structured QPLIB matrices are the cases this is for.
template <int NR, int NC, bool VEC>
__device__ __forceinline__ void tile_pass_blocks(const TileArgs& a, float* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R, n = a.n;
  float* rowscal = smem;
  float* rowacc  = smem + 2 * R;
  float* rowpart = smem + 4 * R;

#pragma unroll
  for (int c = 0; c < NC; ++c)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      rowscal[c * R + r] = (row < a.rows_total) ? __ldg(a.crhs[c] + row) : 0.f;
    }
  for (int i = tid; i < R * NR; i += kKktThreads) rowacc[i] = 0.f;
  __syncthreads();

  const int nchunk = (n + kChunkCols - 1) / kChunkCols;
  unsigned long long blkword = 0ull;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kChunkCols + warp * kSlabCols + lane * 4;
    const bool active = col < n;
    float4 rv[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < NR; ++k) rv[k] = load_vec4<VEC>(a.rrhs[k], col, n);
    float4 cacc[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) cacc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the occupancy words of the CTA's row groups (R <= 64 -> at most 8), one load per CTA and warp instead of a dependent load
    // per group: lane g holds word g; bit g of `mygroups` = this warp's slab is non-empty in row group g
    if (cc == 0) {
      const int g0 = a.r0 >> 3;                               // R and r0 are multiples of 8 (= kRowUnroll)
      blkword = (lane < (R >> 3) && a.r0 + lane * 8 < a.rows_total) ? __ldg(a.blk + g0 + lane) : 0ull;
    }
    const unsigned mygroups = __ballot_sync(kFullMask, (blkword >> (cc * kKktWarps + warp)) & 1ull);

    for (int rg = 0; rg < R; rg += kRowUnroll) {
      float4 v[kRowUnroll];
      {
        if (!((mygroups >> (rg >> 3)) & 1u)) {                 // this warp's slab holds no non-zero in rows rg..rg+7
          if (NR > 0) {
            constexpr int NV0 = kRowUnroll * (NR > 0 ? NR : 1);
            constexpr int kGroup0 = 32 / NV0;
            if ((lane % kGroup0) == 0) rowpart[warp * (R * NR) + rg * NR + lane / kGroup0] = 0.f;
          }
          continue;
        }
      }
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) {
        const int row = a.r0 + rg + u;
        v[u] = (active && row < a.rows_total) ? load_cols4<VEC>(a.mat + (size_t)row * n, col, n)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (NC > 0) {
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float s = rowscal[c * R + rg + u];
            cacc[c].x = fmaf(v[u].x, s, cacc[c].x);
            cacc[c].y = fmaf(v[u].y, s, cacc[c].y);
            cacc[c].z = fmaf(v[u].z, s, cacc[c].z);
            cacc[c].w = fmaf(v[u].w, s, cacc[c].w);
          }
        }
      }
      if (NR > 0) {
        float rp[kRowUnroll * (NR > 0 ? NR : 1)];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int k = 0; k < NR; ++k) {
            float t = v[u].x * rv[k].x;
            t = fmaf(v[u].y, rv[k].y, t);
            t = fmaf(v[u].z, rv[k].z, t);
            t = fmaf(v[u].w, rv[k].w, t);
            rp[u * NR + k] = t;
          }
        }
        constexpr int NV = kRowUnroll * (NR > 0 ? NR : 1);
        const float tot = warp_transpose_reduce<NV>(rp, lane);
        constexpr int kGroup = 32 / NV;              // lanes holding the same value
        if ((lane % kGroup) == 0) {
          const int idx = lane / kGroup;             // = u*NR + k
          rowpart[warp * (R * NR) + rg * NR + idx] = tot;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) store_cols4<VEC>(a.cpart[c], col, n, cacc[c]);

    if (NR > 0) {
      __syncthreads();
      for (int i = tid; i < R * NR; i += kKktThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kKktWarps; ++w) s += rowpart[w * (R * NR) + i];
        rowacc[i] += s;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < NR; ++k)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      if (row < a.rows_total) a.rout[k][row] = rowacc[r * NR + k];
    }
}

// The same pass over a matrix in bitmap-slab form.  Kept as a separate function so that the dense pass above compiles exactly
// as before (sharing one body cost the dense kernels 2x the registers).
// smem: rowscal[2][R] | rowacc[R*2] | rowpart[warps][R*2] | slab masks [warps][R] uint4 | slab offsets [warps][R] u32
template <int NR, int NC, bool VEC>
__device__ __forceinline__ void tile_pass_sparse(const TileArgs& a, float* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R, n = a.n;
  constexpr bool SP = true;
  float* rowscal = smem;
  float* rowacc  = smem + 2 * R;
  float* rowpart = smem + 4 * R;

#pragma unroll
  for (int c = 0; c < NC; ++c)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      rowscal[c * R + r] = (row < a.rows_total) ? __ldg(a.crhs[c] + row) : 0.f;
    }
  for (int i = tid; i < R * NR; i += kKktThreads) rowacc[i] = 0.f;
  __syncthreads();

  const int nchunk = (n + kChunkCols - 1) / kChunkCols;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kChunkCols + warp * kSlabCols + lane * 4;
    const bool active = col < n;
    float4 rv[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < NR; ++k) rv[k] = load_vec4<VEC>(a.rrhs[k], col, n);
    float4 cacc[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c) cacc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    // SP: stage the masks and value offsets of this warp's 128-column slab for all R rows of the CTA in shared memory with
    // one round of independent loads (a dependent load per row group would make the pass latency bound), and keep a bit per
    // row "slab not empty" (R <= 64)
    uint4* smk = nullptr;
    uint32_t* sof = nullptr;
    unsigned long long occ = 0ull;
    if (SP) {
      smk = reinterpret_cast<uint4*>(smem + 4 * R + kKktWarps * R * 2) + warp * R;
      sof = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(smem + 4 * R + kKktWarps * R * 2) + kKktWarps * R) + warp * R;
      const int slab = cc * kKktWarps + warp;
      __syncwarp();                                   // the previous column chunk's readers are done
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = i * 32 + lane;
        uint4 mk = make_uint4(0u, 0u, 0u, 0u);
        uint32_t of = 0u;
        if (r < R && slab < a.sp.S && a.r0 + r < a.rows_total) {
          const size_t rs = a.inst * a.sp.mask_stride + (size_t)(a.r0 + r) * a.sp.S + slab;
          mk = __ldg(a.sp.mask + rs);
          of = __ldg(a.sp.off + rs);
        }
        if (r < R) { smk[r] = mk; sof[r] = of; }
        occ |= (unsigned long long)__ballot_sync(kFullMask, (mk.x | mk.y | mk.z | mk.w) != 0u) << (32 * i);
      }
      __syncwarp();
    }

    // SP: only rows whose slab is not empty are visited, 8 at a time in increasing row order (`todo` = their bit set).  A
    // skipped row contributes exact zeros to every sum, so the results are those of the dense pass bit for bit, and a pass
    // over identity blocks / truly sparse rows costs the mask bytes plus work proportional to the non-empty (row, slab) pairs.
    unsigned long long todo = occ;
    if (SP && NR > 0) {
      for (int i = lane; i < R * NR; i += 32) rowpart[warp * (R * NR) + i] = 0.f;
      __syncwarp();
    }
    for (int rg = 0; SP ? (todo != 0ull) : (rg < R); rg += kRowUnroll) {
      float4 v[kRowUnroll];
      int rsel[kRowUnroll];                      // SP: the CTA-local rows of this group (-1 = none)
      if (SP) {
        const float* vals = a.sp.vals + a.inst * a.sp.vals_stride;
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
          rsel[u] = -1;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (todo != 0ull) {
            const int r = __ffsll((long long)todo) - 1;
            todo &= todo - 1ull;
            rsel[u] = r;
            v[u] = expand_cols4_sparse(vals, smk[r], sof[r], lane);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
          const int row = a.r0 + rg + u;
          v[u] = (active && row < a.rows_total) ? load_cols4<VEC>(a.mat + (size_t)row * n, col, n)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (NC > 0) {
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float s = SP ? ((rsel[u] >= 0) ? rowscal[c * R + rsel[u]] : 0.f) : rowscal[c * R + rg + u];
            cacc[c].x = fmaf(v[u].x, s, cacc[c].x);
            cacc[c].y = fmaf(v[u].y, s, cacc[c].y);
            cacc[c].z = fmaf(v[u].z, s, cacc[c].z);
            cacc[c].w = fmaf(v[u].w, s, cacc[c].w);
          }
        }
      }
      if (NR > 0) {
        float rp[kRowUnroll * (NR > 0 ? NR : 1)];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
#pragma unroll
          for (int k = 0; k < NR; ++k) {
            float t = v[u].x * rv[k].x;
            t = fmaf(v[u].y, rv[k].y, t);
            t = fmaf(v[u].z, rv[k].z, t);
            t = fmaf(v[u].w, rv[k].w, t);
            rp[u * NR + k] = t;
          }
        }
        constexpr int NV = kRowUnroll * (NR > 0 ? NR : 1);
        const float tot = warp_transpose_reduce<NV>(rp, lane);
        constexpr int kGroup = 32 / NV;              // lanes holding the same value
        if ((lane % kGroup) == 0) {
          const int idx = lane / kGroup;             // = u*NR + k
          if (SP) {
            int rr = -1;                             // row of slot u = idx / NR (static unroll: rsel stays in registers)
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u) if (u == idx / (NR > 0 ? NR : 1)) rr = rsel[u];
            if (rr >= 0) rowpart[warp * (R * NR) + rr * NR + idx % (NR > 0 ? NR : 1)] = tot;
          } else {
            rowpart[warp * (R * NR) + rg * NR + idx] = tot;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) store_cols4<VEC>(a.cpart[c], col, n, cacc[c]);

    if (NR > 0) {
      __syncthreads();
      for (int i = tid; i < R * NR; i += kKktThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kKktWarps; ++w) s += rowpart[w * (R * NR) + i];
        rowacc[i] += s;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < NR; ++k)
    for (int r = tid; r < R; r += kKktThreads) {
      const int row = a.r0 + r;
      if (row < a.rows_total) a.rout[k][row] = rowacc[r * NR + k];
    }
}

static size_t tile_smem_bytes(int R, bool sparse) {
  return (size_t)(4 * R + kKktWarps * R * 2) * sizeof(float) + (sparse ? (size_t)kKktWarps * R * 20 : 0);
}

// ------------------------------------------------------------------------------------------------
// pass 1 / pass 2 kernels.  grid = (chunks_q + chunks_a, B)
// ------------------------------------------------------------------------------------------------
struct Pass1Args {
  KktDims d;
  const float *Q, *A0;
  SpMat spq, spa;                    // bitmap-slab forms (used by the SPQ / SPA instantiations)
  const float *xt; long xt_stride;   // x~  (first n entries of xv)
  const float *v;  long v_stride;    // v   (last m entries of xv)
  const float *x, *y;                // previous iterate [B,n], [B,m]
  KktScratch s;
};

// MQ / MA: form of Q / A0 -- 0 dense, 1 bitmap slabs, 2 dense with block skipping
template <bool VEC, int MQ = 0, int MA = 0>
__global__ void __launch_bounds__(kKktThreads, (MQ == 1 || MA == 1) ? 3 : 0) kkt_pass1_kernel(const Pass1Args P) {
  constexpr bool SPQ = (MQ == 1), SPA = (MA == 1);
  extern __shared__ float smem[];
  const KktDims& d = P.d;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const size_t n = d.n, m = d.m;
  TileArgs a;
  a.n = d.n; a.R = d.rows_per_chunk;
  a.rrhs[0] = P.xt + (size_t)b * P.xt_stride;
  a.rrhs[1] = P.x + b * n;
  if (chunk < d.chunks_q) {
    a.mat = SPQ ? nullptr : P.Q + b * n * n; a.sp = P.spq; a.inst = b; a.rows_total = d.n; a.r0 = chunk * a.R;
    a.blk = (MQ == 2) ? P.spq.blk + b * P.spq.blk_stride : nullptr;
    a.rout[0] = P.s.qxt + b * n; a.rout[1] = P.s.qx + b * n;
    a.crhs[0] = a.crhs[1] = nullptr; a.cpart[0] = a.cpart[1] = nullptr;
    if constexpr (MQ == 1) tile_pass_sparse<2, 0, VEC>(a, smem);
    else if constexpr (MQ == 2) tile_pass_blocks<2, 0, VEC>(a, smem);
    else tile_pass<2, 0, VEC>(a, smem);
  } else {
    const int ca = chunk - d.chunks_q;
    a.mat = SPA ? nullptr : P.A0 + b * m * n; a.sp = P.spa; a.inst = b; a.rows_total = d.m; a.r0 = ca * a.R;
    a.blk = (MA == 2) ? P.spa.blk + b * P.spa.blk_stride : nullptr;
    a.rout[0] = P.s.axt + b * m; a.rout[1] = P.s.ax + b * m;
    a.crhs[0] = P.v + (size_t)b * P.v_stride; a.crhs[1] = P.y + b * m;
    float* part = P.s.part_a + ((size_t)b * d.chunks_a + ca) * 2 * n;
    a.cpart[0] = part; a.cpart[1] = part + n;
    if constexpr (MA == 1) tile_pass_sparse<2, 2, VEC>(a, smem);
    else if constexpr (MA == 2) tile_pass_blocks<2, 2, VEC>(a, smem);
    else tile_pass<2, 2, VEC>(a, smem);
  }
}

template <bool VEC, int MQ = 0, int MA = 0>
__global__ void __launch_bounds__(kKktThreads, (MQ == 1 || MA == 1) ? 3 : 0) kkt_pass2_kernel(const KktDims d, const float* __restrict__ Q,
                                                                const float* __restrict__ A0, const KktScratch s,
                                                                const SpMat spq, const SpMat spa) {
  constexpr bool SPQ = (MQ == 1), SPA = (MA == 1);
  extern __shared__ float smem[];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const size_t n = d.n, m = d.m, N = n + m;
  TileArgs a;
  a.n = d.n; a.R = d.rows_per_chunk;
  const float* w1 = s.w + b * N;
  const float* w2 = w1 + n;
  a.rrhs[1] = nullptr; a.rout[1] = nullptr; a.crhs[1] = nullptr; a.cpart[1] = nullptr;
  if (chunk < d.chunks_q) {
    a.mat = SPQ ? nullptr : Q + b * n * n; a.sp = spq; a.inst = b; a.rows_total = d.n; a.r0 = chunk * a.R;
    a.blk = (MQ == 2) ? spq.blk + b * spq.blk_stride : nullptr;
    a.rrhs[0] = nullptr; a.rout[0] = nullptr;
    a.crhs[0] = w1; a.cpart[0] = s.part_q + ((size_t)b * d.chunks_q + chunk) * n;
    if constexpr (MQ == 1) tile_pass_sparse<0, 1, VEC>(a, smem);
    else if constexpr (MQ == 2) tile_pass_blocks<0, 1, VEC>(a, smem);
    else tile_pass<0, 1, VEC>(a, smem);
  } else {
    const int ca = chunk - d.chunks_q;
    a.mat = SPA ? nullptr : A0 + b * m * n; a.sp = spa; a.inst = b; a.rows_total = d.m; a.r0 = ca * a.R;
    a.blk = (MA == 2) ? spa.blk + b * spa.blk_stride : nullptr;
    a.rrhs[0] = w1; a.rout[0] = s.aw1 + b * m;
    a.crhs[0] = w2; a.cpart[0] = s.part_a + ((size_t)b * d.chunks_a + ca) * 2 * n;
    if constexpr (MA == 1) tile_pass_sparse<1, 1, VEC>(a, smem);
    else if constexpr (MA == 2) tile_pass_blocks<1, 1, VEC>(a, smem);
    else tile_pass<1, 1, VEC>(a, smem);
  }
}

// bulk-copy-staged forms of the two dense passes (tile_pass_tma): 8 consumer warps + 1 producer warp, 2 CTAs per SM
__global__ void __launch_bounds__(kKktTmaThreads, 2) kkt_pass1_tma_kernel(const Pass1Args P) {
  extern __shared__ float smem[];
  const KktDims& d = P.d;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const size_t n = d.n, m = d.m;
  TileArgs a;
  a.n = d.n; a.R = d.rows_per_chunk; a.inst = b; a.blk = nullptr;
  a.rrhs[0] = P.xt + (size_t)b * P.xt_stride;
  a.rrhs[1] = P.x + b * n;
  if (chunk < d.chunks_q) {
    a.mat = P.Q + b * n * n; a.rows_total = d.n; a.r0 = chunk * a.R;
    a.rout[0] = P.s.qxt + b * n; a.rout[1] = P.s.qx + b * n;
    a.crhs[0] = a.crhs[1] = nullptr; a.cpart[0] = a.cpart[1] = nullptr;
    tile_pass_tma<2, 0>(a, smem);
  } else {
    const int ca = chunk - d.chunks_q;
    a.mat = P.A0 + b * m * n; a.rows_total = d.m; a.r0 = ca * a.R;
    a.rout[0] = P.s.axt + b * m; a.rout[1] = P.s.ax + b * m;
    a.crhs[0] = P.v + (size_t)b * P.v_stride; a.crhs[1] = P.y + b * m;
    float* part = P.s.part_a + ((size_t)b * d.chunks_a + ca) * 2 * n;
    a.cpart[0] = part; a.cpart[1] = part + n;
    tile_pass_tma<2, 2>(a, smem);
  }
}

__global__ void __launch_bounds__(kKktTmaThreads, 2) kkt_pass2_tma_kernel(const KktDims d, const float* __restrict__ Q,
                                                                          const float* __restrict__ A0, const KktScratch s) {
  extern __shared__ float smem[];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const size_t n = d.n, m = d.m, N = n + m;
  TileArgs a;
  a.n = d.n; a.R = d.rows_per_chunk; a.inst = b; a.blk = nullptr;
  const float* w1 = s.w + b * N;
  const float* w2 = w1 + n;
  a.rrhs[1] = nullptr; a.rout[1] = nullptr; a.crhs[1] = nullptr; a.cpart[1] = nullptr;
  if (chunk < d.chunks_q) {
    a.mat = Q + b * n * n; a.rows_total = d.n; a.r0 = chunk * a.R;
    a.rrhs[0] = nullptr; a.rout[0] = nullptr;
    a.crhs[0] = w1; a.cpart[0] = s.part_q + ((size_t)b * d.chunks_q + chunk) * n;
    tile_pass_tma<0, 1>(a, smem);
  } else {
    const int ca = chunk - d.chunks_q;
    a.mat = A0 + b * m * n; a.rows_total = d.m; a.r0 = ca * a.R;
    a.rrhs[0] = w1; a.rout[0] = s.aw1 + b * m;
    a.crhs[0] = w2; a.cpart[0] = s.part_a + ((size_t)b * d.chunks_a + ca) * 2 * n;
    tile_pass_tma<1, 1>(a, smem);
  }
}

// Which calls take the bulk-copy-staged kernels: dense row-major matrices with 16-byte aligned rows in the solve's 32- / 64-row
// chunks (training's small chunks and the sparse forms keep the load-instruction kernels).  Same-box A/B, bit-identical results
// (tools/kkt_tma_ab.py, profiles/r02_kkt_tma_ab.jsonl): KKT phase 0.660 -> 0.646 ms at the headline size, 1.660 -> 1.522 ms
// at n = 5000 (5.78 -> 6.31 TB/s).
constexpr int kMaxRowsPerChunk = 64;          // make_kkt_dims / make_kkt_dims_train never exceed it
static bool use_tma_pass(const KktDims& d, bool vec, int mq, int ma) {
  if (!vec || mq || ma || d.rows_per_chunk < 32 || d.rows_per_chunk > kMaxRowsPerChunk || d.rows_per_chunk % kRowUnroll) return false;
  const char* e = dev_env("IADMM_KKT_TMA");              // development switch: 0 = load-instruction kernels
  return !(e && e[0] == '0');
}

static int sp_mode(const SpMat* m) { return !m ? 0 : (m->vals ? 1 : (m->blk ? 2 : 0)); }

static bool can_vectorise(const KktDims& d, const void* Q, const void* A0, const KktSparse* sp) {
  // the 128-bit path needs 16-byte aligned matrix rows (n % 4 == 0); a matrix given in bitmap-slab form has no such constraint
  return (d.n % 4 == 0) && ((sp && sp->q.vals) || aligned16(Q)) && ((sp && sp->a.vals) || aligned16(A0));
}

template <bool VEC, int MQ>
static void launch_pass1_a(const Pass1Args& P, int ma, dim3 grid, size_t smem, cudaStream_t st) {
  if (ma == 1)      kkt_pass1_kernel<VEC, MQ, 1><<<grid, kKktThreads, smem, st>>>(P);
  else if (ma == 2) kkt_pass1_kernel<VEC, MQ, 2><<<grid, kKktThreads, smem, st>>>(P);
  else              kkt_pass1_kernel<VEC, MQ, 0><<<grid, kKktThreads, smem, st>>>(P);
}
template <bool VEC>
static void launch_pass1_variant(const Pass1Args& P, int mq, int ma, dim3 grid, size_t smem, cudaStream_t st) {
  if (mq == 1)      launch_pass1_a<VEC, 1>(P, ma, grid, smem, st);
  else if (mq == 2) launch_pass1_a<VEC, 2>(P, ma, grid, smem, st);
  else              launch_pass1_a<VEC, 0>(P, ma, grid, smem, st);
}

static int launch_pass1_common(Pass1Args& P, const KktSparse* sp, cudaStream_t st, bool pdl = false) {
  const KktDims& d = P.d;
  const int mq = sp_mode(sp ? &sp->q : nullptr), ma = d.m > 0 ? sp_mode(sp ? &sp->a : nullptr) : 0;
  if (mq) P.spq = sp->q;
  if (ma) P.spa = sp->a;
  const dim3 grid(d.chunks_q + d.chunks_a, d.B);
  if ((mq || ma) && (d.rows_per_chunk > 64 || d.rows_per_chunk % 8)) IADMM_FAIL(IADMM_EMODE, "sparse KKT pass: row chunks must be 8..64 rows");
  if ((mq == 2 || ma == 2) && d.n > 64 * 128) IADMM_FAIL(IADMM_EMODE, "block-skip KKT pass: at most 8192 columns");
  const size_t smem = tile_smem_bytes(d.rows_per_chunk, mq == 1 || ma == 1);
  const bool vec = can_vectorise(d, P.Q, P.A0, sp);
  if (use_tma_pass(d, vec, mq, ma)) {
    static PerDeviceOnce once;
    int rc;
    // (the opt-in is made once per device: ask for the largest row chunk there is, not for this call's)
    if ((rc = ensure_dyn_smem(kkt_pass1_tma_kernel, (int)tile_smem_bytes_tma(kMaxRowsPerChunk), &once))) return rc;
    launch_kernel_pdl(pdl, kkt_pass1_tma_kernel, grid, dim3(kKktTmaThreads), tile_smem_bytes_tma(d.rows_per_chunk), st, P);
  } else if (vec) launch_pass1_variant<true>(P, mq, ma, grid, smem, st);
  else            launch_pass1_variant<false>(P, mq, ma, grid, smem, st);
  IADMM_LAUNCH_CHECK("kkt_pass1_kernel");
  return IADMM_OK;
}

int launch_kkt_pass1(const KktDims& d, const float* Q, const float* A0, const float* xv, const float* x,
                     const float* y, const KktScratch& s, cudaStream_t st, const KktSparse* sp, bool pdl) {
  Pass1Args P;
  memset(&P.spq, 0, sizeof(SpMat)); memset(&P.spa, 0, sizeof(SpMat));
  P.d = d; P.Q = Q; P.A0 = A0;
  P.xt = xv; P.xt_stride = d.n + d.m;
  P.v = xv + d.n; P.v_stride = d.n + d.m;
  P.x = x; P.y = y; P.s = s;
  return launch_pass1_common(P, sp, st, pdl);
}

// primal_dual_loss on its own (utils.py:68-71): x plays x~ and y plays v, the second product pair is unused
int launch_kkt_pass1_plain(const KktDims& d, const float* Q, const float* A0, const float* x, const float* y,
                           const KktScratch& s, cudaStream_t st) {
  Pass1Args P;
  memset(&P.spq, 0, sizeof(SpMat)); memset(&P.spa, 0, sizeof(SpMat));
  P.d = d; P.Q = Q; P.A0 = A0;
  P.xt = x; P.xt_stride = d.n;
  P.v = y; P.v_stride = d.m;
  P.x = x; P.y = y; P.s = s;
  return launch_pass1_common(P, nullptr, st);
}

template <bool VEC, int MQ>
static void launch_pass2_a(const KktDims& d, const float* Q, const float* A0, const KktScratch& s, const SpMat& q, const SpMat& a,
                           int ma, dim3 grid, size_t smem, cudaStream_t st) {
  if (ma == 1)      kkt_pass2_kernel<VEC, MQ, 1><<<grid, kKktThreads, smem, st>>>(d, Q, A0, s, q, a);
  else if (ma == 2) kkt_pass2_kernel<VEC, MQ, 2><<<grid, kKktThreads, smem, st>>>(d, Q, A0, s, q, a);
  else              kkt_pass2_kernel<VEC, MQ, 0><<<grid, kKktThreads, smem, st>>>(d, Q, A0, s, q, a);
}
template <bool VEC>
static void launch_pass2_variant(const KktDims& d, const float* Q, const float* A0, const KktScratch& s, const SpMat& q, const SpMat& a,
                                 int mq, int ma, dim3 grid, size_t smem, cudaStream_t st) {
  if (mq == 1)      launch_pass2_a<VEC, 1>(d, Q, A0, s, q, a, ma, grid, smem, st);
  else if (mq == 2) launch_pass2_a<VEC, 2>(d, Q, A0, s, q, a, ma, grid, smem, st);
  else              launch_pass2_a<VEC, 0>(d, Q, A0, s, q, a, ma, grid, smem, st);
}

int launch_kkt_pass2(const KktDims& d, const float* Q, const float* A0, const KktScratch& s, cudaStream_t st, const KktSparse* sp,
                     bool pdl) {
  const dim3 grid(d.chunks_q + d.chunks_a, d.B);
  const int mq = sp_mode(sp ? &sp->q : nullptr), ma = d.m > 0 ? sp_mode(sp ? &sp->a : nullptr) : 0;
  if ((mq || ma) && (d.rows_per_chunk > 64 || d.rows_per_chunk % 8)) IADMM_FAIL(IADMM_EMODE, "sparse KKT pass: row chunks must be 8..64 rows");
  if ((mq == 2 || ma == 2) && d.n > 64 * 128) IADMM_FAIL(IADMM_EMODE, "block-skip KKT pass: at most 8192 columns");
  const size_t smem = tile_smem_bytes(d.rows_per_chunk, mq == 1 || ma == 1);
  SpMat q, a;
  memset(&q, 0, sizeof(q)); memset(&a, 0, sizeof(a));
  if (mq) q = sp->q;
  if (ma) a = sp->a;
  const bool vec = can_vectorise(d, Q, A0, sp);
  if (use_tma_pass(d, vec, mq, ma)) {
    static PerDeviceOnce once;
    int rc;
    if ((rc = ensure_dyn_smem(kkt_pass2_tma_kernel, (int)tile_smem_bytes_tma(kMaxRowsPerChunk), &once))) return rc;
    launch_kernel_pdl(pdl, kkt_pass2_tma_kernel, grid, dim3(kKktTmaThreads), tile_smem_bytes_tma(d.rows_per_chunk), st, d, Q, A0, s);
  } else if (vec) launch_pass2_variant<true>(d, Q, A0, s, q, a, mq, ma, grid, smem, st);
  else            launch_pass2_variant<false>(d, Q, A0, s, q, a, mq, ma, grid, smem, st);
  IADMM_LAUNCH_CHECK("kkt_pass2_kernel");
  return IADMM_OK;
}

// ------------------------------------------------------------------------------------------------
// combine 1: w = K xv - rhs, and the residual norms of the previous iterate.  One CTA per instance.
// Element order of every sum follows the reference expression (models/lstm.py:69,72; utils.py:69-70);
// products and sums are rounded separately where the reference rounds them separately.
// ------------------------------------------------------------------------------------------------
constexpr int kCombThreads = 512;

struct Combine1Args {
  KktDims d;
  const float *p, *xt, *v, *x, *y, *z;
  long xt_stride, v_stride;
  const Sched* sched;           // device pointer to this iteration's schedule row (NULL when residual_only)
  float sigma;
  KktScratch s;
  float *pri, *dual, *pri_u, *dual_u;   // trace rows (already offset to the row), may be NULL
  float *metrics;                       // [5][B] row block: objective, ineq max/mean, eq max/mean; may be NULL
  const float *zu;                      // upper bounds (c for inequality rows, b for equality rows)
  const Sched* sched_prev;              // schedule row of the iteration that produced (x,y,z): linear-system residual; may be NULL
  const float *sd, *se, *sc;            // Ruiz diagonals, may be NULL
  int residual_only;
};

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kCombThreads / 32; ++w) t += sh[w];
  return t;
}

__device__ __forceinline__ double block_max_double(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = sh[0];
  for (int w = 1; w < kCombThreads / 32; ++w) t = fmax(t, sh[w]);
  return t;
}

__global__ void __launch_bounds__(kCombThreads) kkt_combine1_kernel(const Combine1Args A) {
  __shared__ double sh[kCombThreads / 32];
  pdl_trigger();
  pdl_wait();
  const KktDims& d = A.d;
  const int b = blockIdx.x;
  const size_t n = d.n, m = d.m, N = n + m;
  const float* part = A.s.part_a + (size_t)b * d.chunks_a * 2 * n;
  const bool want_met = A.metrics != nullptr;
  const bool want_res = (A.pri != nullptr) || (A.dual != nullptr) || (A.pri_u != nullptr) || (A.dual_u != nullptr) || want_met;
  const bool unscaled = (A.sd != nullptr) && ((A.pri_u != nullptr) || (A.dual_u != nullptr) || want_met);
  float inv_ineq = 0.f, inv_eq = 0.f;
  if (!A.residual_only) { inv_ineq = A.sched->inv_rho_ineq; inv_eq = A.sched->inv_rho_eq; }
  const float cscale = unscaled ? A.sc[b] : 1.f;

  double dual2 = 0.0, dual2u = 0.0, pri2 = 0.0, pri2u = 0.0;
  // main.py:949-968 metrics of the iterate (x,y,z): objective 0.5 x^T Q x + p^T x and the constraint violations
  // relu(G x - c), |b - A x|; on the original data they follow from the scaled products through the Ruiz
  // diagonals: x_u^T Q_0 x_u = x^T Q x / c, p_0^T x_u = p^T x / c, (A0_0 x_u)_i = (A0 x)_i / e_i
  double obj = 0.0, isum = 0.0, esum = 0.0;
  float imax = 0.f, emax = 0.f;
  // main.py:952: || K xv - rhs || with the K and rhs of the iteration that produced this iterate, i.e. built from the
  // iterate BEFORE its tail update (x_old, y_old, z_old) and that iteration's penalties, and the xv it produced
  const bool want_ls = want_met && A.sched_prev != nullptr && A.xt != nullptr;
  double ls2 = 0.0;
  float pinv_ineq = 0.f, pinv_eq = 0.f;
  if (want_ls) { pinv_ineq = A.sched_prev->inv_rho_ineq; pinv_eq = A.sched_prev->inv_rho_eq; }
  for (int j = threadIdx.x; j < d.n; j += kCombThreads) {
    float atv = 0.f, aty = 0.f;
    for (int c = 0; c < d.sum_a; ++c) {
      atv += part[((size_t)c * 2 + 0) * n + j];
      aty += part[((size_t)c * 2 + 1) * n + j];
    }
    const float pj = A.p[b * n + j];
    if (!A.residual_only) {
      const float xt  = A.xt[(size_t)b * A.xt_stride + j];
      const float kxv = __fadd_rn(__fadd_rn(A.s.qxt[b * n + j], __fmul_rn(A.sigma, xt)), atv);
      const float rhs = __fsub_rn(__fmul_rn(A.sigma, A.x[b * n + j]), pj);
      A.s.w[b * N + j] = __fsub_rn(kxv, rhs);
    }
    if (want_met) {
      const float xj = A.x[b * n + j];
      obj += (double)xj * (0.5 * (double)A.s.qx[b * n + j] + (double)pj);
    }
    if (want_ls) {
      const float xt  = A.xt[(size_t)b * A.xt_stride + j];
      const float kxv = __fadd_rn(__fadd_rn(A.s.qxt[b * n + j], __fmul_rn(A.sigma, xt)), atv);
      const float r = __fsub_rn(kxv, __fsub_rn(__fmul_rn(A.sigma, A.s.x_old[b * n + j]), pj));
      ls2 += (double)r * (double)r;
    }
    if (want_res) {
      const float r = __fadd_rn(__fadd_rn(A.s.qx[b * n + j], pj), aty);
      dual2 += (double)r * (double)r;
      if (unscaled) {
        const float ru = r / (cscale * A.sd[b * n + j]);
        dual2u += (double)ru * (double)ru;
      }
    }
  }
  for (int i = threadIdx.x; i < d.m; i += kCombThreads) {
    const float zi = A.z[b * m + i];
    if (!A.residual_only) {
      const float inv = (i < d.num_ineq) ? inv_ineq : inv_eq;
      const float vi  = A.v[(size_t)b * A.v_stride + i];
      const float kxv = __fsub_rn(A.s.axt[b * m + i], __fmul_rn(inv, vi));
      const float rhs = __fsub_rn(zi, __fmul_rn(inv, A.y[b * m + i]));
      A.s.w[b * N + n + i] = __fsub_rn(kxv, rhs);
    }
    if (want_met) {
      const float einv = unscaled ? 1.0f / A.se[b * m + i] : 1.0f;
      const float dv = (A.s.ax[b * m + i] - A.zu[b * m + i]) * einv;
      if (i < d.num_ineq) { const float v = fmaxf(dv, 0.f); imax = fmaxf(imax, v); isum += (double)v; }
      else                { const float v = fabsf(dv);      emax = fmaxf(emax, v); esum += (double)v; }
    }
    if (want_ls) {
      const float inv = (i < d.num_ineq) ? pinv_ineq : pinv_eq;
      const float vi  = A.v[(size_t)b * A.v_stride + i];
      const float kxv = __fsub_rn(A.s.axt[b * m + i], __fmul_rn(inv, vi));
      const float r = __fsub_rn(kxv, __fsub_rn(A.s.z_old[b * m + i], __fmul_rn(inv, A.s.y_old[b * m + i])));
      ls2 += (double)r * (double)r;
    }
    if (want_res) {
      const float r = __fsub_rn(A.s.ax[b * m + i], zi);
      pri2 += (double)r * (double)r;
      if (unscaled) {
        const float ru = r / A.se[b * m + i];
        pri2u += (double)ru * (double)ru;
      }
    }
  }
  if (want_res) {
    pri2  = block_sum_double(pri2, sh);
    dual2 = block_sum_double(dual2, sh);
    if (unscaled) { pri2u = block_sum_double(pri2u, sh); dual2u = block_sum_double(dual2u, sh); }
    if (want_met) {
      obj = block_sum_double(obj, sh); isum = block_sum_double(isum, sh); esum = block_sum_double(esum, sh);
      imax = (float)block_max_double((double)imax, sh); emax = (float)block_max_double((double)emax, sh);
      ls2 = block_sum_double(ls2, sh);
    }
    if (threadIdx.x == 0) {
      if (A.pri)  A.pri[b]  = (float)sqrt(pri2);
      if (A.dual) A.dual[b] = (float)sqrt(dual2);
      if (unscaled && A.pri_u)  A.pri_u[b]  = (float)sqrt(pri2u);
      if (unscaled && A.dual_u) A.dual_u[b] = (float)sqrt(dual2u);
      if (want_met) {
        const size_t B = d.B;
        const int me = d.m - d.num_ineq;
        A.metrics[0 * B + b] = (float)(unscaled ? obj / (double)cscale : obj);
        A.metrics[1 * B + b] = imax;
        A.metrics[2 * B + b] = d.num_ineq > 0 ? (float)(isum / d.num_ineq) : 0.f;
        A.metrics[3 * B + b] = emax;
        A.metrics[4 * B + b] = me > 0 ? (float)(esum / me) : 0.f;
        A.metrics[5 * B + b] = want_ls ? (float)sqrt(ls2) : 0.f;
      }
    }
  }
}

// With many small chunks (training at a few instances per GPU) the per-column loops over the chunk partials in the
// combine kernels become a serial chain in too few CTAs.  fold_partials sums them with 4 x 64 threads per 64 columns
// into chunk 0 (fixed order: four interleaved groups, then ((g0+g1)+g2)+g3) and the combine kernels read one partial.
constexpr int kFoldThreshold = 32;
__global__ void __launch_bounds__(256) fold_partials_kernel(float* __restrict__ part, int chunks, int slots, int n) {
  __shared__ float sh[4][64];
  const int t = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + t;
  float* base = part + ((size_t)blockIdx.z * chunks * slots + blockIdx.y) * n;
  float s = 0.f;
  if (j < n) {
#pragma unroll 8
    for (int c = g; c < chunks; c += 4) s += base[(size_t)c * slots * n + j];
  }
  sh[g][t] = s;
  __syncthreads();
  if (g == 0 && j < n) base[j] = ((sh[0][t] + sh[1][t]) + sh[2][t]) + sh[3][t];
}
static int fold_partials(float* part, int B, int chunks, int slots, int nfold, int n, cudaStream_t st) {
  fold_partials_kernel<<<dim3(cdiv(n, 64), nfold, B), 256, 0, st>>>(part, chunks, slots, n);
  IADMM_LAUNCH_CHECK("fold_partials_kernel");
  return IADMM_OK;
}
int kkt_sum_chunks(int chunks) { return chunks >= kFoldThreshold ? 1 : chunks; }

int launch_kkt_combine1(const KktDims& d, const float* p, const float* xv, const float* x, const float* y,
                        const float* z, const Sched* sched_t, float sigma, const KktScratch& s,
                        float* pri_trace, float* dual_trace, float* pri_trace_u, float* dual_trace_u,
                        const float* sd, const float* se, const float* sc, int trace_row, int residual_only,
                        cudaStream_t st, float* metric_trace, const float* zu, const Sched* sched_prev, bool pdl) {
  Combine1Args A;
  A.d = d; A.p = p; A.x = x; A.y = y; A.z = z;
  A.xt = xv; A.v = xv ? xv + d.n : nullptr; A.xt_stride = A.v_stride = d.n + d.m;
  A.sched = sched_t; A.sigma = sigma; A.s = s;
  const size_t off = trace_row >= 0 ? (size_t)trace_row * d.B : 0;
  A.pri    = (trace_row >= 0 && pri_trace)    ? pri_trace + off    : nullptr;
  A.dual   = (trace_row >= 0 && dual_trace)   ? dual_trace + off   : nullptr;
  A.pri_u  = (trace_row >= 0 && pri_trace_u)  ? pri_trace_u + off  : nullptr;
  A.dual_u = (trace_row >= 0 && dual_trace_u) ? dual_trace_u + off : nullptr;
  A.metrics = (trace_row >= 0 && metric_trace && (zu || d.m == 0)) ? metric_trace + (size_t)trace_row * kMetricRows * d.B : nullptr;
  A.zu = zu; A.sched_prev = sched_prev;
  A.sd = sd; A.se = se; A.sc = sc;
  A.residual_only = residual_only;
  if (d.m > 0 && d.chunks_a >= kFoldThreshold) {
    int rc = fold_partials(s.part_a, d.B, d.chunks_a, 2, 2, d.n, st);
    if (rc) return rc;
    A.d.sum_a = 1;
  }
  launch_kernel_pdl(pdl, kkt_combine1_kernel, dim3(d.B), dim3(kCombThreads), 0, st, A);
  IADMM_LAUNCH_CHECK("kkt_combine1_kernel");
  return IADMM_OK;
}

// ------------------------------------------------------------------------------------------------
// combine 2: g = K^T w   (element-wise over B*(n+m))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kkt_combine2_kernel(const KktDims d, const Sched* __restrict__ sched,
                                                           float sigma, const KktScratch s) {
  pdl_trigger();
  pdl_wait();
  const size_t n = d.n, m = d.m, N = n + m;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d.B * N) return;
  const size_t b = idx / N;
  const int    r = (int)(idx - b * N);
  if (r < d.n) {
    float qtw = 0.f, atw = 0.f;
    const float* pq = s.part_q + b * d.chunks_q * n + r;
    for (int c = 0; c < d.sum_q; ++c) qtw += pq[(size_t)c * n];
    const float* pa = s.part_a + b * d.chunks_a * 2 * n + r;
    for (int c = 0; c < d.sum_a; ++c) atw += pa[(size_t)c * 2 * n];
    const float w1 = s.w[b * N + r];
    s.g[idx] = __fadd_rn(__fadd_rn(qtw, __fmul_rn(sigma, w1)), atw);
  } else {
    const int   i   = r - d.n;
    const float inv = (i < d.num_ineq) ? sched->inv_rho_ineq : sched->inv_rho_eq;
    s.g[idx] = __fsub_rn(s.aw1[b * m + i], __fmul_rn(inv, s.w[idx]));
  }
}

int launch_kkt_combine2(const KktDims& d, const Sched* sched_t, float sigma, const KktScratch& s, cudaStream_t st, bool pdl) {
  const size_t total = (size_t)d.B * (d.n + d.m);
  KktDims d2 = d;
  int rc;
  if (d.chunks_q >= kFoldThreshold) { if ((rc = fold_partials(s.part_q, d.B, d.chunks_q, 1, 1, d.n, st))) return rc; d2.sum_q = 1; }
  if (d.m > 0 && d.chunks_a >= kFoldThreshold) {     // slot 0 only is live after pass 2
    if ((rc = fold_partials(s.part_a, d.B, d.chunks_a, 2, 1, d.n, st))) return rc;
    d2.sum_a = 1;
  }
  launch_kernel_pdl(pdl, kkt_combine2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, d2, sched_t, sigma, s);
  IADMM_LAUNCH_CHECK("kkt_combine2_kernel");
  return IADMM_OK;
}

// ------------------------------------------------------------------------------------------------
// O(N) tail of the iteration (models/lstm.py:80-94): head bias, xv step, x relaxation, z projection,
// dual update.  Every product/sum is rounded separately, in the reference's order: equality rows carry
// rho ~ 500 and y + rho*(z~ - z) cancels catastrophically, so a contracted FMA here would show up as a
// 1e-4-level difference in y.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tail_kernel(const KktDims d, const float* __restrict__ head_part, int tiles,
                                                   const float* __restrict__ b_h, const Sched* __restrict__ sched,
                                                   const float* __restrict__ zl, const float* __restrict__ zu,
                                                   float* __restrict__ x, float* __restrict__ y, float* __restrict__ z,
                                                   float* __restrict__ xv, float* __restrict__ x_old,
                                                   float* __restrict__ y_old, float* __restrict__ z_old) {
  pdl_trigger();            // (the next iteration's pass 1 may be a programmatic dependent of this kernel)
  const size_t n = d.n, m = d.m, N = n + m;
  const size_t rows = (size_t)d.B * N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows) return;
  float head = 0.f;
  for (int t = 0; t < tiles; ++t) head += head_part[(size_t)t * rows + idx];
  head = __fadd_rn(head, b_h[0]);
  const float xvn = __fsub_rn(xv[idx], head);
  xv[idx] = xvn;
  const size_t b = idx / N;
  const int    r = (int)(idx - b * N);
  if (r < d.n) {
    const float a = sched->alpha, oma = sched->one_minus_alpha;
    const size_t j = b * n + r;
    const float xo = x[j];
    if (x_old) x_old[j] = xo;
    x[j] = __fadd_rn(__fmul_rn(a, xvn), __fmul_rn(oma, xo));
  } else {
    const int    i   = r - d.n;
    const size_t k   = b * m + i;
    const bool   eq  = i >= d.num_ineq;
    const float  rho = eq ? sched->rho_eq : sched->rho_ineq;
    const float  inv = eq ? sched->inv_rho_eq : sched->inv_rho_ineq;
    const float  yo = y[k], zo = z[k];
    if (y_old) { y_old[k] = yo; z_old[k] = zo; }
    const float  zmid = __fadd_rn(zo, __fmul_rn(inv, __fsub_rn(xvn, yo)));
    const float  zc   = fmaxf(fminf(__fadd_rn(zmid, __fmul_rn(inv, yo)), zu[k]), zl[k]);
    z[k] = zc;
    y[k] = __fadd_rn(yo, __fmul_rn(rho, __fsub_rn(zmid, zc)));
  }
}

int launch_tail(const KktDims& d, const float* head_part, int tiles, const float* b_h, const Sched* sched_t,
                const float* zl, const float* zu, float* x, float* y, float* z, float* xv, cudaStream_t st,
                const KktScratch* keep_old) {
  const size_t rows = (size_t)d.B * (d.n + d.m);
  tail_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(d, head_part, tiles, b_h, sched_t, zl, zu, x, y, z, xv,
                                                              keep_old ? keep_old->x_old : nullptr,
                                                              keep_old ? keep_old->y_old : nullptr,
                                                              keep_old ? keep_old->z_old : nullptr);
  IADMM_LAUNCH_CHECK("tail_kernel");
  return IADMM_OK;
}

// ------------------------------------------------------------------------------------------------
// dense K / rhs / rho_vec materialisation for API compatibility (models/lstm.py:61-62,67-69)
// ------------------------------------------------------------------------------------------------
// 64 x 64 tiles that never straddle a block boundary of K = [[Q + sigma I, A0^T], [A0, -diag(1/rho)]]: blockIdx.x / .y run over
// the column / row tiles of the first n and then of the last m indices.  Every global access is a whole 128-byte line per warp;
// the A0^T block goes through a shared-memory transpose (the first version read it with stride n: 1.8 TB/s).
constexpr int kBkT = 64;
template <bool VEC>
__global__ void __launch_bounds__(256) build_kkt_kernel(int B, int n, int m, int num_ineq, const float* __restrict__ Q,
                                                        const float* __restrict__ p, const float* __restrict__ A0,
                                                        const float* __restrict__ x, const float* __restrict__ y,
                                                        const float* __restrict__ z, const Sched* __restrict__ sched,
                                                        float sigma, float* __restrict__ K, float* __restrict__ rhs,
                                                        float* __restrict__ rho_vec) {
  __shared__ float tile[kBkT][kBkT + 1];
  const size_t N = (size_t)n + m;
  const size_t b = blockIdx.z;
  const int tn = (n + kBkT - 1) / kBkT;                       // tiles over the first n indices
  const bool low = (int)blockIdx.y >= tn, right = (int)blockIdx.x >= tn;
  const int r0 = low ? ((int)blockIdx.y - tn) * kBkT : (int)blockIdx.y * kBkT;      // origin inside the block
  const int c0 = right ? ((int)blockIdx.x - tn) * kBkT : (int)blockIdx.x * kBkT;
  const int rlim = low ? m : n, clim = right ? m : n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (K != nullptr) {
    float* Kb = K + b * N * N;
    if (!low && right) {
      // A0^T block: K[r][n + j] = A0[j][r]
      const float* Ab = A0 + b * (size_t)m * n;
#pragma unroll
      for (int i = 0; i < kBkT; i += 8) {
        const int j = c0 + ty + i;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int r = r0 + tx + 32 * h2;
          tile[ty + i][tx + 32 * h2] = (j < m && r < n) ? Ab[(size_t)j * n + r] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kBkT; i += 8) {
        const int r = r0 + ty + i;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int j = c0 + tx + 32 * h2;
          if (r < n && j < m) Kb[(size_t)r * N + n + j] = tile[tx + 32 * h2][ty + i];
        }
      }
    } else if (VEC) {
      // n % 4 == 0 and m % 4 == 0: 128-bit accesses, 16 threads per 64-column row, all loads of a thread issued before its stores
      const float* src = low ? A0 + b * (size_t)m * n : Q + b * (size_t)n * n;
      const int c = c0 + 4 * (int)(threadIdx.x & 15), rr = r0 + (int)(threadIdx.x >> 4);
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr + 16 * i;
        v[i] = make_float4(-0.0f, -0.0f, -0.0f, -0.0f);          // -(1/rho) * 0 in the reference is -0.0
        if (!right && r < rlim && c < clim) v[i] = *reinterpret_cast<const float4*>(src + (size_t)r * n + c);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr + 16 * i;
        if (r >= rlim || c >= clim) continue;
        if (low == right && r >= c && r < c + 4) {                // the diagonal crosses this quad
          float* e = reinterpret_cast<float*>(&v[i]) + (r - c);
          *e = low ? -((r < num_ineq) ? sched->inv_rho_ineq : sched->inv_rho_eq) : __fadd_rn(*e, sigma);
        }
        *reinterpret_cast<float4*>(Kb + ((size_t)(low ? n : 0) + r) * N + (right ? n : 0) + c) = v[i];
      }
    } else {
      const float* src = low ? A0 + b * (size_t)m * n : Q + b * (size_t)n * n;
#pragma unroll
      for (int i = 0; i < kBkT; i += 8) {
        const int r = r0 + ty + i;
        if (r >= rlim) break;
        const float inv_r = (low && right) ? ((r < num_ineq) ? sched->inv_rho_ineq : sched->inv_rho_eq) : 0.f;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int c = c0 + tx + 32 * h2;
          if (c >= clim) continue;
          float v;
          if (!right) {
            v = src[(size_t)r * n + c];
            if (!low && c == r) v = __fadd_rn(v, sigma);
          } else {
            v = (c == r) ? -inv_r : -0.0f;        // -(1/rho) * 0 in the reference is -0.0
          }
          Kb[((size_t)(low ? n : 0) + r) * N + (right ? n : 0) + c] = v;
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < kBkT) {
    const int r = r0 + (int)threadIdx.x;
    if (!low && r < n) rhs[b * N + r] = __fsub_rn(__fmul_rn(sigma, x[b * n + r]), p[b * n + r]);
    if (low && r < m) {
      const bool in = r < num_ineq;
      const float inv_r = in ? sched->inv_rho_ineq : sched->inv_rho_eq;
      rhs[b * N + n + r] = __fsub_rn(z[b * m + r], __fmul_rn(inv_r, y[b * m + r]));
      rho_vec[b * m + r] = in ? sched->rho_ineq : sched->rho_eq;
    }
  }
}

int launch_build_kkt(int B, int n, int m, int num_ineq, const float* Q, const float* p, const float* A0,
                     const float* x, const float* y, const float* z, const Sched* sched_t, float sigma,
                     float* K, float* rhs, float* rho_vec, cudaStream_t st) {
  const int tiles = cdiv(n, kBkT) + cdiv(m, kBkT);
  const dim3 grid(K == nullptr ? 1 : tiles, tiles, B);
  const bool vec = n % 4 == 0 && m % 4 == 0 && K != nullptr && aligned16(K) && aligned16(Q) && (m == 0 || aligned16(A0));
  if (vec) build_kkt_kernel<true><<<grid, 256, 0, st>>>(B, n, m, num_ineq, Q, p, A0, x, y, z, sched_t, sigma, K, rhs, rho_vec);
  else     build_kkt_kernel<false><<<grid, 256, 0, st>>>(B, n, m, num_ineq, Q, p, A0, x, y, z, sched_t, sigma, K, rhs, rho_vec);
  IADMM_LAUNCH_CHECK("build_kkt_kernel");
  return IADMM_OK;
}

// the -(1/rho_t) diagonal of the last m rows of a dense K (the only entries of K that depend on the iteration)
__global__ void __launch_bounds__(256) kkt_penalty_diag_kernel(int B, int n, int m, int num_ineq, const Sched* __restrict__ sched,
                                                               float* __restrict__ K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * m) return;
  const size_t b = idx / m, N = (size_t)n + m;
  const int i = (int)(idx - b * m);
  K[(b * N + n + i) * N + n + i] = -((i < num_ineq) ? sched->inv_rho_ineq : sched->inv_rho_eq);
}

int launch_kkt_penalty_diag(int B, int n, int m, int num_ineq, const Sched* sched_t, float* K, cudaStream_t st) {
  if (m == 0) return IADMM_OK;
  kkt_penalty_diag_kernel<<<(unsigned)(((size_t)B * m + 255) / 256), 256, 0, st>>>(B, n, m, num_ineq, sched_t, K);
  IADMM_LAUNCH_CHECK("kkt_penalty_diag_kernel");
  return IADMM_OK;
}

}  // namespace iadmm
