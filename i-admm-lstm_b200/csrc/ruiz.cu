// Modified Ruiz equilibration with cost normalisation, O(n^2) streaming form.
//
// Reference: methods/scaling.py:50-119.  The reference multiplies by dense diag matrices (6 [n,n]x[n,n]
// bmm per iteration).  Every one of those products has a diagonal factor, so each output entry is ONE
// rounded product; we therefore apply the same products element-wise, in the same order,
//     Q <- d_i * ((c_prev * Q_ij) * d_j)          A0 <- e_i * (A0_ij * d_j)
// and reproduce the reference's fp32 values exactly (the only order-dependent quantity is the mean of
// the n column norms in the cost scaling, accumulated here in double).  The cost factor c_t of
// iteration k is applied lazily by the matrix pass of iteration k+1 (rounding commutes with the
// column max because c_t > 0).
//
// Traffic.  Round 1 rescaled the matrices in place every iteration (one read + one write pass each: (1 + 2*ites + 1/2) *
// 4(n^2+mn) bytes, 21.5 passes at 10 iterations).  The entries themselves are only needed at the end: iteration k needs the
// inf-norms of the matrix scaled by the diagonals of iterations 0..k-1.  The "chain" passes therefore read the ORIGINAL
// matrices and re-apply, in registers and in the same order, the rounded products of all previous iterations (3 fp32
// multiplies per entry and iteration for Q, 2 for A0 -- the pass stays HBM bound up to 10 iterations), take the norms, and
// write nothing.  Re-applying ALL previous iterations turned out compute bound (measured: 11.8 ms against 9.0 ms in place at
// config 2), so the state is materialised every kRzPeriod = 3 iterations: 12 reads + 4 writes = 16 matrix passes instead of 21.5
// at 10 iterations, bit-identical to the in-place form (same operands, same IEEE multiplies, same order -- checked by hash,
// tools/ruiz_ab.py).  Needs the per-iteration diagonals ((n+m) floats per instance and iteration); more than kRzMaxHist
// iterations fall back to the in-place form.
#include "common.cuh"

namespace iadmm {

constexpr int kRzThreads = 256;
constexpr int kRzWarps   = 8;
constexpr int kRzUnroll  = 8;
constexpr int kRzChunkCols = 128 * kRzWarps;
constexpr int kRzPeriod = 3;           // chain passes: the state is materialised every kRzPeriod iterations
constexpr int kRzMaxChain = kRzPeriod; // steps a chain pass re-applies in registers
constexpr int kRzMaxHist = 32;         // iterations whose diagonals are kept (more: in-place form)
constexpr float kMinScaling = 1e-4f;   // scaling.py:12
constexpr float kMaxScaling = 1e4f;    // scaling.py:13

__device__ __forceinline__ float limit_scaling(float v) {   // scaling.py:31-38
  float w = fminf(fmaxf(v, kMinScaling), kMaxScaling);
  return (w == kMinScaling) ? 1.0f : w;
}

struct RuizWs {            // per-instance vectors, fp32
  float* sd;               // [hist][B,n] D_temp diagonal of iteration k at slot k*hist_on (hist = 1: the current one only)
  float* se;               // [hist][B,m] E_temp diagonal
  float* cprev;            // [hist+1][B] cost factor of iteration k-1 (applied lazily by iteration k) at slot k*hist_on
  int hist_on;             // 1: per-iteration history kept (chain passes), 0: slot 0 is overwritten every iteration
  int B;
  float* rowmax;           // [B,m] row inf-norms of the current A0
  float* partq;            // [B,chunks_q,n] column inf-norm partials of the current (pre-cost) Q
  float* parta;            // [B,chunks_a,n]
  int R, chunks_q, chunks_a;
  int scan_q, scan_a;      // chunk partials ruiz_vec_kernel still has to fold per column (1 after ruiz_fold_max_kernel)
};

enum { kRzNorm = 0, kRzScale = 1, kRzCost = 2 };

// grid = (chunks_q + chunks_a, B).  MODE kRzNorm: norms of src.  kRzScale: dst = scaled src + norms of
// dst.  kRzCost (Q chunks only): dst = cprev * src.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kRzThreads)
ruiz_pass_kernel(const float* Qsrc, const float* Asrc, float* Qdst, float* Adst, int n, int m, RuizWs W) {
  // src may alias dst (in-place rescaling): every element is read and then written by the same thread,
  // and the loads below are coherent (no .nc).
  __shared__ float rowpart[kRzWarps][128];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const bool isQ = chunk < W.chunks_q;
  if (MODE == kRzCost && !isQ) return;
  const int ca = isQ ? chunk : chunk - W.chunks_q;
  const int rows_total = isQ ? n : m;
  const int R = W.R, r0 = ca * R;
  const float* src = isQ ? Qsrc + (size_t)b * n * n : Asrc + (size_t)b * m * n;
  float*       dst = isQ ? (Qdst ? Qdst + (size_t)b * n * n : nullptr) : (Adst ? Adst + (size_t)b * m * n : nullptr);
  const float* sd = W.sd + (size_t)b * n;
  const float* srow = isQ ? sd : W.se + (size_t)b * m;       // left diagonal factor
  const float cprev = (MODE != kRzNorm && isQ) ? W.cprev[b] : 1.0f;
  float* colpart = isQ ? W.partq + ((size_t)b * W.chunks_q + ca) * n : W.parta + ((size_t)b * W.chunks_a + ca) * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int nchunk = (n + kRzChunkCols - 1) / kRzChunkCols;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kRzChunkCols + warp * 128 + lane * 4;
    const bool active = col < n;
    float sc[4] = {1.f, 1.f, 1.f, 1.f};
    if (MODE == kRzScale) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (col + e < n) sc[e] = sd[col + e];
    }
    float cmax[4] = {0.f, 0.f, 0.f, 0.f};
    for (int rg = 0; rg < R; rg += kRzUnroll) {
      float v[kRzUnroll][4];
#pragma unroll
      for (int u = 0; u < kRzUnroll; ++u) {
        const int row = r0 + rg + u;
        const bool ok = active && row < rows_total;
        if (ok && VEC) {
          const float4 t = ld_stream4_coherent(src + (size_t)row * n + col);
          v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[u][e] = (ok && col + e < n) ? src[(size_t)row * n + col + e] : 0.f;
        }
      }
      float rmax[kRzUnroll];
#pragma unroll
      for (int u = 0; u < kRzUnroll; ++u) {
        const int row = r0 + rg + u;
        const bool ok = active && row < rows_total;
        if (MODE == kRzScale) {
          const float sr = (row < rows_total) ? srow[row] : 1.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t = v[u][e];
            if (isQ) t = __fmul_rn(cprev, t);            // lazily applied c_{k-1} (scaling.py:101)
            t = __fmul_rn(t, sc[e]);                     // bmm(M, D_temp)       (scaling.py:80-81)
            v[u][e] = __fmul_rn(sr, t);                  // bmm(D_temp|E_temp, .)
          }
        } else if (MODE == kRzCost) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[u][e] = __fmul_rn(cprev, v[u][e]);
        }
        if (MODE != kRzNorm && ok) {
          float* drow = dst + (size_t)row * n + col;
          if (VEC) *reinterpret_cast<float4*>(drow) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
          else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (col + e < n) drow[e] = v[u][e];
          }
        }
        float rm = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = fabsf(v[u][e]);
          cmax[e] = fmaxf(cmax[e], a);
          rm = fmaxf(rm, a);
        }
        rmax[u] = rm;
      }
      if (MODE != kRzCost && !isQ) {
        const float tot = warp_transpose_reduce<kRzUnroll, true>(rmax, lane);
        if ((lane & 3) == 0) rowpart[warp][rg + (lane >> 2)] = tot;
      }
    }
    if (MODE != kRzCost && active) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (col + e < n) colpart[col + e] = cmax[e];
    }
    if (MODE != kRzCost && !isQ) {
      __syncthreads();
      for (int r = tid; r < R; r += kRzThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRzWarps; ++w) t = fmaxf(t, rowpart[w][r]);
        const int row = r0 + r;
        if (row < rows_total) {
          float* rm = W.rowmax + (size_t)b * m + row;
          *rm = (cc == 0) ? t : fmaxf(*rm, t);
        }
      }
      __syncthreads();
    }
  }
}

// Chain pass: reads a materialised state of the matrices (the originals, or Qs/A0s as written by an earlier chain pass),
// re-applies the scale steps slot0 .. slot0+steps-1 in registers (same rounded products, same order as the in-place form),
// and takes the norms of the result (NORM) and/or writes it (WRITE; with final_cost the pending last cost factor is applied
// first).  src may be dst (every entry is read and then written by the same thread; loads are coherent).
// grid = (chunks_q + chunks_a, B).
template <bool NORM, bool WRITE, bool VEC>
__global__ void __launch_bounds__(kRzThreads)
ruiz_chain_kernel(const float* Qsrc, const float* Asrc, float* Qdst, float* Adst, int n, int m, int slot0, int steps, int final_cost,
                  RuizWs W) {
  __shared__ float rowpart[kRzWarps][128];
  __shared__ float srow_s[kRzMaxChain][128];       // left diagonal factors of this CTA's rows, per step (R <= 128)
  __shared__ float cstep[kRzMaxChain + 1];         // Q: cost factor applied before step j; [steps] = the pending last one
  const int b = blockIdx.y, chunk = blockIdx.x;
  const bool isQ = chunk < W.chunks_q;
  const int ca = isQ ? chunk : chunk - W.chunks_q;
  const int rows_total = isQ ? n : m;
  const int R = W.R, r0 = ca * R;
  const float* src = isQ ? Qsrc + (size_t)b * n * n : Asrc + (size_t)b * m * n;
  float*       dst = WRITE ? (isQ ? Qdst + (size_t)b * n * n : Adst + (size_t)b * m * n) : nullptr;
  float* colpart = isQ ? W.partq + ((size_t)b * W.chunks_q + ca) * n : W.parta + ((size_t)b * W.chunks_a + ca) * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rlen = isQ ? n : m;
  const float* left = isQ ? W.sd : W.se;
  for (int i = tid; i < steps * R; i += kRzThreads) {
    const int j = i / R, r = i - j * R, row = r0 + r;
    srow_s[j][r] = (row < rows_total) ? left[((size_t)(slot0 + j) * W.B + b) * rlen + row] : 1.0f;
  }
  if (tid < steps + (final_cost ? 1 : 0)) cstep[tid] = isQ ? W.cprev[(size_t)(slot0 + tid) * W.B + b] : 1.0f;
  __syncthreads();

  const int nchunk = (n + kRzChunkCols - 1) / kRzChunkCols;
  for (int cc = 0; cc < nchunk; ++cc) {
    const int  col    = cc * kRzChunkCols + warp * 128 + lane * 4;
    const bool active = col < n;
    u64 sc[kRzMaxChain][2];                         // right diagonal factors of this lane's 4 columns, per step, as pairs
#pragma unroll
    for (int j = 0; j < kRzMaxChain; ++j) {
      float t[4] = {1.f, 1.f, 1.f, 1.f};
      if (j < steps) {
        const float* sdj = W.sd + ((size_t)(slot0 + j) * W.B + b) * n;
#pragma unroll
        for (int e = 0; e < 4; ++e) if (col + e < n) t[e] = sdj[col + e];
      }
      sc[j][0] = pk2(t[0], t[1]); sc[j][1] = pk2(t[2], t[3]);
    }
    float cmax[4] = {0.f, 0.f, 0.f, 0.f};
    for (int rg = 0; rg < R; rg += kRzUnroll) {
      u64 v[kRzUnroll][2];
#pragma unroll
      for (int u = 0; u < kRzUnroll; ++u) {
        const int row = r0 + rg + u;
        const bool ok = active && row < rows_total;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok && VEC) {
          const float4 q = ld_stream4_coherent(src + (size_t)row * n + col);
          t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
        } else if (ok) {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (col + e < n) t[e] = src[(size_t)row * n + col + e];
        }
        v[u][0] = pk2(t[0], t[1]); v[u][1] = pk2(t[2], t[3]);
      }
      float rmax[kRzUnroll];
#pragma unroll
      for (int u = 0; u < kRzUnroll; ++u) {
        const int row = r0 + rg + u;
        const bool ok = active && row < rows_total;
        u64 a0 = v[u][0], a1 = v[u][1];
#pragma unroll
        for (int j = 0; j < kRzMaxChain; ++j) {
          if (j < steps) {
            if (isQ) { const u64 cj = bc2(cstep[j]); a0 = mul2(cj, a0); a1 = mul2(cj, a1); }   // lazily applied c_{j-1} (scaling.py:101)
            a0 = mul2(a0, sc[j][0]); a1 = mul2(a1, sc[j][1]);                                     // bmm(M, D_temp)   (scaling.py:80-81)
            const u64 sr = bc2(srow_s[j][rg + u]);
            a0 = mul2(sr, a0); a1 = mul2(sr, a1);                                                 // bmm(D_temp|E_temp, .)
          }
        }
        float t[4];
        if (WRITE) {
          if (isQ && final_cost) { const u64 cl = bc2(cstep[steps]); a0 = mul2(cl, a0); a1 = mul2(cl, a1); }  // the last cost factor
          upk2(a0, t[0], t[1]); upk2(a1, t[2], t[3]);
          if (ok) {
            float* drow = dst + (size_t)row * n + col;
            if (VEC) *reinterpret_cast<float4*>(drow) = make_float4(t[0], t[1], t[2], t[3]);
            else {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (col + e < n) drow[e] = t[e];
            }
          }
          rmax[u] = 0.f;
        }
        if (NORM) {
          upk2(a0, t[0], t[1]); upk2(a1, t[2], t[3]);
          float rm = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = fabsf(t[e]);
            cmax[e] = fmaxf(cmax[e], a);
            rm = fmaxf(rm, a);
          }
          rmax[u] = rm;
        }
      }
      if (NORM && !isQ) {
        const float tot = warp_transpose_reduce<kRzUnroll, true>(rmax, lane);
        if ((lane & 3) == 0) rowpart[warp][rg + (lane >> 2)] = tot;
      }
    }
    if (NORM && active) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (col + e < n) colpart[col + e] = cmax[e];
    }
    if (NORM && !isQ) {
      __syncthreads();
      for (int r = tid; r < R; r += kRzThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRzWarps; ++w) t = fmaxf(t, rowpart[w][r]);
        const int row = r0 + r;
        if (row < rows_total) {
          float* rm = W.rowmax + (size_t)b * m + row;
          *rm = (cc == 0) ? t : fmaxf(*rm, t);
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per instance: finish iteration k-1 (cost normalisation, vector updates) and prepare the
// diagonal factors of iteration k.
constexpr int kRzVecThreads = 512;

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = sh[0];
  for (int w = 1; w < kRzVecThreads / 32; ++w) t = fmaxf(t, sh[w]);
  return t;
}
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kRzVecThreads / 32; ++w) t += sh[w];
  return t;
}

__global__ void __launch_bounds__(kRzVecThreads)
ruiz_vec_kernel(int n, int m, int finish_prev, int prepare_next, int k, float* __restrict__ p, float* __restrict__ zl,
                float* __restrict__ zu, float* __restrict__ d, float* __restrict__ e, float* __restrict__ c, RuizWs W) {
  // k = iteration being finished (finish_prev) and/or the one before the iteration being prepared (prepare_next, k = -1 at the start)
  __shared__ float shf[kRzVecThreads / 32];
  __shared__ double shd[kRzVecThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t kf = (size_t)(k < 0 ? 0 : k) * W.hist_on, kn = (size_t)(k + 1) * W.hist_on;     // history slots
  const float* sd = W.sd + (kf * W.B + b) * n;
  const float* se = W.se + (kf * W.B + b) * m;
  float* sd_next = W.sd + (kn * W.B + b) * n;
  float* se_next = W.se + (kn * W.B + b) * m;
  const float* partq = W.partq + (size_t)b * W.chunks_q * n;
  const float* parta = W.parta + (size_t)b * W.chunks_a * n;
  p += (size_t)b * n; d += (size_t)b * n;
  zl += (size_t)b * m; zu += (size_t)b * m; e += (size_t)b * m;

  float cprev = 1.0f;
  if (finish_prev) {
    // scaling.py:82-105 for the iteration whose matrix pass just ran
    double colsum = 0.0;
    float pmax = 0.f;
    for (int j = tid; j < n; j += kRzVecThreads) {
      float cq = 0.f;
      for (int ch = 0; ch < W.scan_q; ++ch) cq = fmaxf(cq, partq[(size_t)ch * n + j]);
      colsum += (double)cq;
      const float pj = __fmul_rn(sd[j], p[j]);
      p[j] = pj;
      pmax = fmaxf(pmax, fabsf(pj));
      d[j] = __fmul_rn(sd[j], d[j]);
    }
    for (int i = tid; i < m; i += kRzVecThreads) {
      zl[i] = __fmul_rn(se[i], zl[i]);
      zu[i] = __fmul_rn(se[i], zu[i]);
      e[i]  = __fmul_rn(se[i], e[i]);
    }
    colsum = block_sum_d(colsum, shd);
    pmax = block_max(pmax, shf);
    const float mean_col = __fdiv_rn((float)colsum, (float)n);
    const float cost = limit_scaling(fmaxf(limit_scaling(pmax), mean_col));
    cprev = __frcp_rn(cost);
    for (int j = tid; j < n; j += kRzVecThreads) p[j] = __fmul_rn(cprev, p[j]);
    if (tid == 0) { c[b] = __fmul_rn(cprev, c[b]); W.cprev[kn * W.B + b] = cprev; }
  } else {
    for (int j = tid; j < n; j += kRzVecThreads) d[j] = 1.0f;
    for (int i = tid; i < m; i += kRzVecThreads) e[i] = 1.0f;
    if (tid == 0) { c[b] = 1.0f; W.cprev[kn * W.B + b] = 1.0f; }
  }
  if (prepare_next) {
    // scaling.py:66-69 on the matrix as it stands (Q carries the pending factor cprev)
    for (int j = tid; j < n; j += kRzVecThreads) {
      float cq = 0.f, ca = 0.f;
      for (int ch = 0; ch < W.scan_q; ++ch) cq = fmaxf(cq, partq[(size_t)ch * n + j]);
      for (int ch = 0; ch < W.scan_a; ++ch) ca = fmaxf(ca, parta[(size_t)ch * n + j]);
      const float nrm = fmaxf(__fmul_rn(cprev, cq), ca);
      sd_next[j] = __frcp_rn(__fsqrt_rn(limit_scaling(nrm)));   // reciprocal(sqrt(.)), two roundings like scaling.py:68-69
    }
    const float* rowmax = W.rowmax + (size_t)b * m;
    for (int i = tid; i < m; i += kRzVecThreads) se_next[i] = __frcp_rn(__fsqrt_rn(limit_scaling(rowmax[i])));
  }
}

// With many row chunks (n = 5000: 79 + 79) the one-CTA-per-instance vector kernel above spent 370 us per call walking the chunk
// partials of its columns (24 instances = 24 CTAs on 148 SMs, a dependent strided load per chunk).  This folds them first with
// one thread per (instance, column) over the whole GPU, into chunk 0 -- a maximum, so the order is immaterial and the result
// bit-identical -- and the vector kernel reads one value per column.
__global__ void __launch_bounds__(256) ruiz_fold_max_kernel(float* __restrict__ partq, int chunks_q, float* __restrict__ parta,
                                                            int chunks_a, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t b = blockIdx.y;
  float* pq = partq + b * chunks_q * n + j;
  float cq = 0.f;
#pragma unroll 8
  for (int ch = 0; ch < chunks_q; ++ch) cq = fmaxf(cq, pq[(size_t)ch * n]);
  pq[0] = cq;
  if (chunks_a > 0) {
    float* pa = parta + b * chunks_a * n + j;
    float ca = 0.f;
#pragma unroll 8
    for (int ch = 0; ch < chunks_a; ++ch) ca = fmaxf(ca, pa[(size_t)ch * n]);
    pa[0] = ca;
  }
}
constexpr int kRzFoldChunks = 24;      // fold first when an instance has at least this many chunk partials per column

size_t ruiz_ws_floats(int B, int n, int m, int* R_out, int* cq_out, int* ca_out) {
  const KktDims kd = make_kkt_dims(B, n, m, 0);
  int R = kd.rows_per_chunk > 128 ? 128 : kd.rows_per_chunk;
  const int cq = cdiv(n, R), ca = cdiv(m, R);
  if (R_out) { *R_out = R; *cq_out = cq; *ca_out = ca; }
  const size_t hist = kRzMaxHist + 1;           // per-iteration diagonals for the chain passes
  return hist * ((size_t)B * n + (size_t)B * m + B) + (size_t)B * m + (size_t)B * cq * n + (size_t)B * ca * n + 64 + 4 * 8;
}

int ruiz_impl(const float* Q, const float* p, const float* A0, const float* zl, const float* zu, float* Qs, float* ps,
              float* A0s, float* zls, float* zus, float* d, float* e, float* c, int B, int n, int m, int iterations,
              void* workspace, size_t workspace_bytes, cudaStream_t st) {
  RuizWs W;
  const size_t need = ruiz_ws_floats(B, n, m, &W.R, &W.chunks_q, &W.chunks_a) * sizeof(float);
  if (workspace_bytes < need) IADMM_FAIL(IADMM_EWORK, "ruiz workspace too small: %zu < %zu", workspace_bytes, need);
  float* base = static_cast<float*>(workspace);
  auto take = [&](size_t cnt) { float* q = base; base += (cnt + 3) / 4 * 4; return q; };
  const size_t hist = kRzMaxHist + 1;
  W.B = B;
  W.sd = take(hist * (size_t)B * n); W.se = take(hist * (size_t)B * m); W.cprev = take(hist * (size_t)B); W.rowmax = take((size_t)B * m);
  W.partq = take((size_t)B * W.chunks_q * n); W.parta = take((size_t)B * W.chunks_a * n);
  const char* sw = dev_env("IADMM_RUIZ_CHAIN");                       // development switch: 0 = in-place form (round 1)
  const bool chain = iterations <= kRzMaxHist && !(sw && sw[0] == '0');
  W.hist_on = chain ? 1 : 0;
  W.scan_q = W.chunks_q; W.scan_a = W.chunks_a;
  // the vector stage after a matrix pass (its column partials are fresh): fold them GPU-wide first when there are many
  const char* swf = dev_env("IADMM_RUIZ_FOLD");                       // development switch: 0 = the vector kernel walks the partials
  const bool fold = W.chunks_q + W.chunks_a >= kRzFoldChunks && !(swf && swf[0] == '0');
  auto launch_vec = [&](int finish_prev, int prepare_next, int k) -> int {
    RuizWs Wv = W;
    if ((finish_prev || prepare_next) && fold) {
      ruiz_fold_max_kernel<<<dim3(cdiv(n, 256), B), 256, 0, st>>>(W.partq, W.chunks_q, W.parta, m > 0 ? W.chunks_a : 0, n);
      IADMM_LAUNCH_CHECK("ruiz_fold_max_kernel");
      Wv.scan_q = 1; Wv.scan_a = (m > 0 && W.chunks_a > 0) ? 1 : 0;
    }
    ruiz_vec_kernel<<<B, kRzVecThreads, 0, st>>>(n, m, finish_prev, prepare_next, k, ps, zls, zus, d, e, c, Wv);
    IADMM_LAUNCH_CHECK("ruiz_vec_kernel");
    return IADMM_OK;
  };
  int rcv;

  IADMM_CUDA(cudaMemcpyAsync(ps, p, (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (m > 0) {
    IADMM_CUDA(cudaMemcpyAsync(zls, zl, (size_t)B * m * sizeof(float), cudaMemcpyDeviceToDevice, st));
    IADMM_CUDA(cudaMemcpyAsync(zus, zu, (size_t)B * m * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  const bool vec = (n % 4 == 0) && aligned16(Q) && aligned16(A0) && aligned16(Qs) && aligned16(A0s);
  const dim3 grid(W.chunks_q + W.chunks_a, B);
  if (iterations == 0) {
    IADMM_CUDA(cudaMemcpyAsync(Qs, Q, (size_t)B * n * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (m > 0) IADMM_CUDA(cudaMemcpyAsync(A0s, A0, (size_t)B * m * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return launch_vec(0, 0, -1);
  }
  if (vec) ruiz_pass_kernel<kRzNorm, true><<<grid, kRzThreads, 0, st>>>(Q, A0, nullptr, nullptr, n, m, W);
  else     ruiz_pass_kernel<kRzNorm, false><<<grid, kRzThreads, 0, st>>>(Q, A0, nullptr, nullptr, n, m, W);
  IADMM_LAUNCH_CHECK("ruiz_pass_kernel<norm>");
  if ((rcv = launch_vec(0, 1, -1))) return rcv;
  if (chain) {
    // Iteration k needs the norms of the matrices under the scale steps 0..k.  They are taken from the last MATERIALISED state
    // (the originals, later Qs/A0s) by re-applying the steps since then in registers; every kRzPeriod-th iteration the pass also
    // writes its result, which bounds the re-applied steps at kRzPeriod (more and the pass turns compute bound: measured 0.16
    // ms per step and pass at config 2 against 0.33 ms of HBM time per read).  10 iterations: 12 reads + 4 writes instead of
    // the 11 + 10.5 of the in-place form.
    const float* qsrc = Q;
    const float* asrc = A0;
    int slot0 = 0;
    for (int k = 0; k < iterations; ++k) {
      const int steps = k - slot0 + 1;
      const bool write = (steps == kRzPeriod) && (k + 1 < iterations);
      if (write) {
        if (vec) ruiz_chain_kernel<true, true, true><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, slot0, steps, 0, W);
        else     ruiz_chain_kernel<true, true, false><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, slot0, steps, 0, W);
      } else {
        if (vec) ruiz_chain_kernel<true, false, true><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, nullptr, nullptr, n, m, slot0, steps, 0, W);
        else     ruiz_chain_kernel<true, false, false><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, nullptr, nullptr, n, m, slot0, steps, 0, W);
      }
      IADMM_LAUNCH_CHECK("ruiz_chain_kernel<norm>");
      if ((rcv = launch_vec(1, (k + 1 < iterations) ? 1 : 0, k))) return rcv;
      if (write) { qsrc = Qs; asrc = A0s; slot0 = k + 1; }
    }
    // the last write: the steps since the last materialised state and the last cost factor
    if (vec) ruiz_chain_kernel<false, true, true><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, slot0, iterations - slot0, 1, W);
    else     ruiz_chain_kernel<false, true, false><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, slot0, iterations - slot0, 1, W);
    IADMM_LAUNCH_CHECK("ruiz_chain_kernel<write>");
    return IADMM_OK;
  }
  for (int k = 0; k < iterations; ++k) {
    const float* qsrc = (k == 0) ? Q : Qs;
    const float* asrc = (k == 0) ? A0 : A0s;
    if (vec) ruiz_pass_kernel<kRzScale, true><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, W);
    else     ruiz_pass_kernel<kRzScale, false><<<grid, kRzThreads, 0, st>>>(qsrc, asrc, Qs, A0s, n, m, W);
    IADMM_LAUNCH_CHECK("ruiz_pass_kernel<scale>");
    if ((rcv = launch_vec(1, (k + 1 < iterations) ? 1 : 0, k))) return rcv;
  }
  if (vec) ruiz_pass_kernel<kRzCost, true><<<grid, kRzThreads, 0, st>>>(Qs, A0s, Qs, A0s, n, m, W);
  else     ruiz_pass_kernel<kRzCost, false><<<grid, kRzThreads, 0, st>>>(Qs, A0s, Qs, A0s, n, m, W);
  IADMM_LAUNCH_CHECK("ruiz_pass_kernel<cost>");
  return IADMM_OK;
}

}  // namespace iadmm
