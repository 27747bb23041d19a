// Stage II ("feasibility restoration") linear algebra: batched dense LU with partial pivoting and the
// triangular solves, hand-written.
//
// Reference: models/lu.py:28-35 -- `torch.lu(A_tild, pivot=True)` once, then `torch.lu_solve(b_tild, lu, piv)` in
// every Stage-II iteration (main.py:1035-1115), i.e. LAPACK getrf/getrs semantics on the dense KKT matrix
// K = [[Q + sigma I, A0^T], [A0, -diag(1/rho)]] of every instance ([B, N, N], N = n + m, row-major).
//
// Factorisation: right-looking, 16-column panels.
//   lu_panel_kernel   one CTA (1024 threads) per instance keeps the whole panel (N-k0 rows x 16 columns) in
//                     REGISTERS, one or more rows per thread: per column a block-wide arg-max (first maximum,
//                     like LAPACK's isamax), the pivot row broadcast through shared memory, a predicated
//                     register swap and the rank-1 update of the remaining panel columns.
//   lu_swap_kernel    applies the panel's 16 row interchanges to the columns outside the panel.
//   lu_trsm_kernel    U12 = L11^-1 A12 (one thread per column, L11 in shared memory).
//   lu_update_kernel  A22 -= L21 U12 (rank-16 update, memory bound: one read+write of the trailing matrix).
// Solve: lu_perm_kernel turns the pivot sequence into a permutation once; lu_solve_kernel (one CTA per instance,
// right-hand side in shared memory) runs left-looking forward and backward substitution in 32-row blocks: the
// rows of L and U are contiguous, so every row does ONE coalesced dot product with the already known part of the
// solution, then the 32x32 diagonal block is solved by a single warp from shared memory.
#include "common.cuh"

namespace iadmm {

constexpr int kLuNb = 16;           // panel width
constexpr int kLuThreads = 1024;    // panel kernel threads
constexpr int kLuMaxSlots = 4;      // rows per thread held in registers -> N - k0 <= 4096

template <int SLOTS>
__global__ void __launch_bounds__(kLuThreads, 1)
lu_panel_kernel(float* __restrict__ Aall, int* __restrict__ piv_all, int* __restrict__ info, int N, int k0) {
  __shared__ float s_val[32];
  __shared__ int   s_idx[32];
  __shared__ float s_prow[kLuNb], s_orow[kLuNb];
  __shared__ int   s_piv;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* A = Aall + (size_t)b * N * N;
  int* piv = piv_all + (size_t)b * N;
  const int kb = min(kLuNb, N - k0);

  float a[SLOTS][kLuNb];
  int   rowid[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int r = k0 + tid + s * kLuThreads;
    rowid[s] = r;
#pragma unroll
    for (int c = 0; c < kLuNb; ++c) a[s][c] = (r < N && c < kb) ? A[(size_t)r * N + k0 + c] : 0.f;
  }

#pragma unroll
  for (int j = 0; j < kLuNb; ++j) {
    if (j < kb) {                                    // uniform
      const int prow_id = k0 + j;
      // ---- pivot search: first row with the largest |a[r][j]|, r >= k0 + j
      float best = -1.f; int bidx = 0x7fffffff;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (rowid[s] >= prow_id && rowid[s] < N) {
          const float v = fabsf(a[s][j]);
          if (v > best || (v == best && rowid[s] < bidx)) { best = v; bidx = rowid[s]; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(kFullMask, best, o);
        const int   oi = __shfl_xor_sync(kFullMask, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
      }
      if (lane == 0) { s_val[warp] = best; s_idx[warp] = bidx; }
      __syncthreads();
      if (warp == 0) {
        best = s_val[lane]; bidx = s_idx[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(kFullMask, best, o);
          const int   oi = __shfl_xor_sync(kFullMask, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) {
          s_piv = bidx;
          piv[prow_id] = bidx;
          if (best == 0.f && info) atomicCAS(info + b, 0, prow_id + 1);   // exactly singular (LAPACK info > 0)
        }
      }
      __syncthreads();
      const int p = s_piv;
      // ---- publish the pivot row and the row it replaces
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (rowid[s] == p) {
#pragma unroll
          for (int c = 0; c < kLuNb; ++c) s_prow[c] = a[s][c];
        }
        if (rowid[s] == prow_id) {
#pragma unroll
          for (int c = 0; c < kLuNb; ++c) s_orow[c] = a[s][c];
        }
      }
      __syncthreads();
      // ---- interchange (registers) and eliminate below the pivot
      const float pv = s_prow[j];
      const float rinv = (pv != 0.f) ? 1.0f / pv : 0.f;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (p != prow_id) {
          if (rowid[s] == p) {
#pragma unroll
            for (int c = 0; c < kLuNb; ++c) a[s][c] = s_orow[c];
          } else if (rowid[s] == prow_id) {
#pragma unroll
            for (int c = 0; c < kLuNb; ++c) a[s][c] = s_prow[c];
          }
        }
        if (rowid[s] > prow_id && rowid[s] < N) {
          const float l = a[s][j] * rinv;
          a[s][j] = l;
#pragma unroll
          for (int c = j + 1; c < kLuNb; ++c) a[s][c] = fmaf(-l, s_prow[c], a[s][c]);
        }
      }
      __syncthreads();                               // s_prow / s_orow / s_piv are reused by the next column
    }
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int r = rowid[s];
    if (r < N) {
#pragma unroll
      for (int c = 0; c < kLuNb; ++c)
        if (c < kb) A[(size_t)r * N + k0 + c] = a[s][c];
    }
  }
}

// apply the interchanges of panel [k0, k0+kb) to the columns outside the panel.  grid (ceil(N/256), B)
__global__ void __launch_bounds__(256) lu_swap_kernel(float* __restrict__ Aall, const int* __restrict__ piv_all, int N, int k0,
                                                      int kb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N || (c >= k0 && c < k0 + kb)) return;
  float* A = Aall + (size_t)blockIdx.y * N * N;
  const int* piv = piv_all + (size_t)blockIdx.y * N;
  for (int j = 0; j < kb; ++j) {
    const int r = k0 + j, p = piv[r];
    if (p != r) {
      const float t = A[(size_t)r * N + c];
      A[(size_t)r * N + c] = A[(size_t)p * N + c];
      A[(size_t)p * N + c] = t;
    }
  }
}

// U12 = L11^-1 A12 for the columns right of the panel.  grid (ceil((N-k0-kb)/256), B)
__global__ void __launch_bounds__(256) lu_trsm_kernel(float* __restrict__ Aall, int N, int k0, int kb) {
  __shared__ float L11[kLuNb][kLuNb + 1];
  float* A = Aall + (size_t)blockIdx.y * N * N;
  for (int i = threadIdx.x; i < kLuNb * kLuNb; i += blockDim.x) {
    const int r = i / kLuNb, c = i % kLuNb;
    L11[r][c] = (r < kb && c < kb) ? A[(size_t)(k0 + r) * N + k0 + c] : 0.f;
  }
  __syncthreads();
  const int c = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float u[kLuNb];
#pragma unroll
  for (int i = 0; i < kLuNb; ++i) {
    if (i < kb) {
      float v = A[(size_t)(k0 + i) * N + c];
#pragma unroll
      for (int t = 0; t < kLuNb; ++t)
        if (t < i) v = fmaf(-L11[i][t], u[t], v);
      u[i] = v;
      A[(size_t)(k0 + i) * N + c] = v;
    } else {
      u[i] = 0.f;
    }
  }
}

// A22 -= L21 U12.  block = 256 columns x 32 rows; grid (ceil(cols/256), ceil(rows/32), B)
__global__ void __launch_bounds__(256) lu_update_kernel(float* __restrict__ Aall, int N, int k0, int kb) {
  __shared__ float L21[32][kLuNb + 1];
  float* A = Aall + (size_t)blockIdx.z * N * N;
  const int r0 = k0 + kb + blockIdx.y * 32;
  for (int i = threadIdx.x; i < 32 * kLuNb; i += blockDim.x) {
    const int r = i / kLuNb, t = i % kLuNb;
    L21[r][t] = (r0 + r < N && t < kb) ? A[(size_t)(r0 + r) * N + k0 + t] : 0.f;
  }
  __syncthreads();
  const int c = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float u[kLuNb];
#pragma unroll
  for (int t = 0; t < kLuNb; ++t) u[t] = (t < kb) ? A[(size_t)(k0 + t) * N + c] : 0.f;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    if (r0 + r < N) {
      float v = A[(size_t)(r0 + r) * N + c];
#pragma unroll
      for (int t = 0; t < kLuNb; ++t) v = fmaf(-L21[r][t], u[t], v);
      A[(size_t)(r0 + r) * N + c] = v;
    }
  }
}

// pivot sequence -> permutation (row i of P*K is row perm[i] of K).  One CTA per instance, sequential in smem.
__global__ void __launch_bounds__(256) lu_perm_kernel(const int* __restrict__ piv_all, int* __restrict__ perm_all, int N) {
  extern __shared__ int s_perm[];
  const int* piv = piv_all + (size_t)blockIdx.x * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_perm[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; ++i) {
      const int p = piv[i];
      if (p != i) { const int t = s_perm[i]; s_perm[i] = s_perm[p]; s_perm[p] = t; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) perm_all[(size_t)blockIdx.x * N + i] = s_perm[i];
}

// x = U^-1 L^-1 P b.  One CTA (512 threads) per instance; dynamic smem: x[N] + diag[32][33]
constexpr int kLuSolveThreads = 512;
__global__ void __launch_bounds__(kLuSolveThreads) lu_solve_kernel(const float* __restrict__ LUall, const int* __restrict__ perm_all,
                                                                   float* __restrict__ rhs_all, int N) {
  extern __shared__ float s_mem[];
  float* x = s_mem;
  float (*diag)[33] = reinterpret_cast<float (*)[33]>(s_mem + ((N + 31) / 32) * 32);
  const float* LU = LUall + (size_t)blockIdx.x * N * N;
  const int* perm = perm_all + (size_t)blockIdx.x * N;
  float* rhs = rhs_all + (size_t)blockIdx.x * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kLuSolveThreads / 32;
  for (int i = tid; i < N; i += kLuSolveThreads) x[i] = rhs[perm[i]];
  __syncthreads();
  const int nblk = (N + 31) / 32;
  // ---- forward substitution, unit lower triangle
  for (int blk = 0; blk < nblk; ++blk) {
    const int r0 = blk * 32;
    for (int rr = warp; rr < 32; rr += kWarps) {        // rows of this block: dot with the known part x[0, r0)
      const int r = r0 + rr;
      if (r < N) {
        float acc = 0.f;
        const float* row = LU + (size_t)r * N;
        for (int c = lane; c < r0; c += 32) acc = fmaf(row[c], x[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) x[r] -= acc;
        const int c = r0 + lane;                         // and stage the diagonal block
        diag[rr][lane] = (c < N) ? row[c] : 0.f;
      }
    }
    __syncthreads();
    if (warp == 0) {
      float xv = (r0 + lane < N) ? x[r0 + lane] : 0.f;
      for (int i = 0; i < 32; ++i) {
        const float xi = __shfl_sync(kFullMask, xv, i);
        if (lane > i) xv = fmaf(-diag[lane][i], xi, xv);
      }
      if (r0 + lane < N) x[r0 + lane] = xv;
    }
    __syncthreads();
  }
  // ---- backward substitution, upper triangle
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int r0 = blk * 32;
    const int c_hi = r0 + 32;                            // columns >= c_hi are known
    for (int rr = warp; rr < 32; rr += kWarps) {
      const int r = r0 + rr;
      if (r < N) {
        float acc = 0.f;
        const float* row = LU + (size_t)r * N;
        for (int c = c_hi + lane; c < N; c += 32) acc = fmaf(row[c], x[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) x[r] -= acc;
        const int c = r0 + lane;
        diag[rr][lane] = (c < N) ? row[c] : 0.f;
      }
    }
    __syncthreads();
    if (warp == 0) {
      float xv = (r0 + lane < N) ? x[r0 + lane] : 0.f;
      for (int i = 31; i >= 0; --i) {
        if (r0 + i < N) {
          float xi = 0.f;
          if (lane == i) { xv = xv / diag[i][i]; }
          xi = __shfl_sync(kFullMask, xv, i);
          if (lane < i) xv = fmaf(-diag[lane][i], xi, xv);
        }
      }
      if (r0 + lane < N) x[r0 + lane] = xv;
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += kLuSolveThreads) rhs[i] = x[i];
}

}  // namespace iadmm

using namespace iadmm;

extern "C" {

int iadmm_lu_factor(float* K, int* piv, int* perm, int* info, int B, int N, void* stream) {
  if (B <= 0 || N <= 0) IADMM_FAIL(IADMM_ESHAPE, "lu_factor: B=%d N=%d", B, N);
  if (B > 65535) IADMM_FAIL(IADMM_ESHAPE, "lu_factor: batch %d > 65535", B);
  if (N > kLuThreads * kLuMaxSlots) IADMM_FAIL(IADMM_EMODE, "lu_factor: N=%d exceeds the register-resident panel limit %d", N, kLuThreads * kLuMaxSlots);
  if (!K || !piv || !perm) IADMM_FAIL(IADMM_EALIGN, "lu_factor: NULL pointer");
  int dev = 0, major = 0;
  IADMM_CUDA(cudaGetDevice(&dev));
  IADMM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) IADMM_FAIL(IADMM_EARCH, "device %d has compute capability %d.x; libiadmm_b200 needs sm_100 (B200)", dev, major);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (info) IADMM_CUDA(cudaMemsetAsync(info, 0, (size_t)B * sizeof(int), st));
  for (int k0 = 0; k0 < N; k0 += kLuNb) {
    const int kb = (N - k0 < kLuNb) ? N - k0 : kLuNb;
    const int rows = N - k0;
    if (rows <= kLuThreads)          lu_panel_kernel<1><<<B, kLuThreads, 0, st>>>(K, piv, info, N, k0);
    else if (rows <= 2 * kLuThreads) lu_panel_kernel<2><<<B, kLuThreads, 0, st>>>(K, piv, info, N, k0);
    else                             lu_panel_kernel<4><<<B, kLuThreads, 0, st>>>(K, piv, info, N, k0);
    IADMM_LAUNCH_CHECK("lu_panel_kernel");
    lu_swap_kernel<<<dim3(cdiv(N, 256), B), 256, 0, st>>>(K, piv, N, k0, kb);
    IADMM_LAUNCH_CHECK("lu_swap_kernel");
    const int rest = N - k0 - kb;
    if (rest > 0) {
      lu_trsm_kernel<<<dim3(cdiv(rest, 256), B), 256, 0, st>>>(K, N, k0, kb);
      IADMM_LAUNCH_CHECK("lu_trsm_kernel");
      lu_update_kernel<<<dim3(cdiv(rest, 256), cdiv(rest, 32), B), 256, 0, st>>>(K, N, k0, kb);
      IADMM_LAUNCH_CHECK("lu_update_kernel");
    }
  }
  lu_perm_kernel<<<B, 256, (size_t)N * sizeof(int), st>>>(piv, perm, N);
  IADMM_LAUNCH_CHECK("lu_perm_kernel");
  return IADMM_OK;
}

int iadmm_lu_solve(const float* LU, const int* perm, float* rhs, int B, int N, void* stream) {
  if (B <= 0 || N <= 0) IADMM_FAIL(IADMM_ESHAPE, "lu_solve: B=%d N=%d", B, N);
  if (!LU || !perm || !rhs) IADMM_FAIL(IADMM_EALIGN, "lu_solve: NULL pointer");
  const size_t smem = ((size_t)((N + 31) / 32) * 32 + 32 * 33) * sizeof(float);
  if (smem > 200 * 1024) IADMM_FAIL(IADMM_EMODE, "lu_solve: N=%d does not fit the shared-memory right-hand side", N);
  static PerDeviceOnce attr;
  int rc;
  if ((rc = ensure_dyn_smem(lu_solve_kernel, 200 * 1024, &attr))) return rc;
  lu_solve_kernel<<<B, kLuSolveThreads, smem, static_cast<cudaStream_t>(stream)>>>(LU, perm, rhs, N);
  IADMM_LAUNCH_CHECK("lu_solve_kernel");
  return IADMM_OK;
}

}  // extern "C"
