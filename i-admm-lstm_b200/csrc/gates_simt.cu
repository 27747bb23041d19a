// Coordinate-wise LSTM cell, fp32 CUDA-core path (IADMM_GATES_SIMT_FP32).
//
// Reference: models/lstm.py:74-80.  For every coordinate r of every instance (rows = B*(n+m)):
//     pre_g = [xv_r, grad_r] @ W_g + H_r @ U_g + b_g        g in {i,f,o,u}
//     C_r   = sigmoid(pre_i) * tanh(pre_u) + sigmoid(pre_f) * C_r
//     H_r   = sigmoid(pre_o) * tanh(C_r)
//     head_r = H_r . W_h  (+ b_h, added by the tail kernel)
// i.e. one [rows,h] x [h,4h] GEMM whose epilogue is the whole cell.  This file is the bit-stable
// validation path (plain fp32 FMA, accurate expf/tanhf); the production path is gates_tc.cu.
// A CTA computes 128 rows x 32 hidden units (128 interleaved gate columns); a thread 8 rows x 2 units.
#include "common.cuh"

namespace iadmm {

constexpr int kSgBM = 128;          // rows per CTA
constexpr int kSgBU = 32;           // hidden units per CTA (x4 gate columns)
constexpr int kSgBN = kSgBU * 4;    // 128 gate columns
constexpr int kSgBK = 16;
constexpr int kSgThreads = 256;

int simt_gate_tiles(int h) { return cdiv(h, kSgBU); }

__global__ void __launch_bounds__(kSgThreads)
gates_simt_kernel(const float* __restrict__ u32, const float* __restrict__ wc, const float* __restrict__ bias,
                  const float* __restrict__ wh, const float* __restrict__ xv, const float* __restrict__ gvec,
                  const float* __restrict__ H_in, float* __restrict__ H_out, float* __restrict__ C,
                  float* __restrict__ head_part, float* __restrict__ gates_out, long rows, int h, int tiles_u) {
  __shared__ float As[kSgBK][kSgBM + 4];
  __shared__ float Bs[kSgBK][kSgBN];

  const long tile = blockIdx.x;
  const int  ut   = (int)(tile % tiles_u);
  const long rt   = tile / tiles_u;
  const long row0 = rt * kSgBM;
  const int  col0 = ut * kSgBN;                // first interleaved gate column
  const int  h4   = 4 * h;
  const int  tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < h; k0 += kSgBK) {
#pragma unroll
    for (int e = 0; e < (kSgBM * kSgBK) / kSgThreads; ++e) {
      const int idx = tid + e * kSgThreads;
      const int k = idx % kSgBK, r = idx / kSgBK;
      const long row = row0 + r;
      As[k][r] = (row < rows && k0 + k < h) ? H_in[row * h + k0 + k] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < (kSgBN * kSgBK) / kSgThreads; ++e) {
      const int idx = tid + e * kSgThreads;
      const int c = idx % kSgBN, k = idx / kSgBN;
      Bs[k][c] = (k0 + k < h && col0 + c < h4) ? u32[(size_t)(k0 + k) * h4 + col0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // fused cell epilogue
  const int unit0 = ut * kSgBU + tx * 2;
  float w0[8], w1[8], bb[8], whv[2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col0 + tx * 8 + j;
    const bool ok = c < h4;
    w0[j] = ok ? wc[c] : 0.f;
    w1[j] = ok ? wc[h4 + c] : 0.f;
    bb[j] = ok ? bias[c] : 0.f;
  }
  whv[0] = (unit0 < h) ? wh[unit0] : 0.f;
  whv[1] = (unit0 + 1 < h) ? wh[unit0 + 1] : 0.f;

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long row = row0 + ty * 8 + i;
    float hp = 0.f;
    if (row < rows) {
      const float xr = xv[row], gr = gvec[row];
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const int unit = unit0 + uu;
        if (unit < h) {
          float pre[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int j = uu * 4 + g;
            const float iw = fmaf(gr, w1[j], __fmul_rn(xr, w0[j]));
            pre[g] = __fadd_rn(__fadd_rn(iw, acc[i][j]), bb[j]);
          }
          const float gi = sigmoid_ref(pre[0]);
          const float gf = sigmoid_ref(pre[1]);
          const float go = sigmoid_ref(pre[2]);
          const float gu = tanhf(pre[3]);
          const size_t o = (size_t)row * h + unit;
          if (gates_out) {   // training: keep the gate activations for the backward pass ([rows][4h], column 4j+g)
            *reinterpret_cast<float4*>(gates_out + (size_t)row * h4 + 4 * (size_t)unit) = make_float4(gi, gf, go, gu);
          }
          const float cn = __fadd_rn(__fmul_rn(gi, gu), __fmul_rn(gf, C[o]));
          const float hn = __fmul_rn(go, tanhf(cn));
          C[o] = cn;
          H_out[o] = hn;
          hp = fmaf(hn, whv[uu], hp);
        }
      }
    }
    // sum over the 16 threads (tx) sharing this row: lanes of one 16-lane half-warp
    hp += __shfl_xor_sync(kFullMask, hp, 8);
    hp += __shfl_xor_sync(kFullMask, hp, 4);
    hp += __shfl_xor_sync(kFullMask, hp, 2);
    hp += __shfl_xor_sync(kFullMask, hp, 1);
    if (tx == 0 && row < rows) head_part[(size_t)ut * rows + row] = hp;
  }
}

int launch_gates_simt(const void* packed, const WeightLayout& L, const float* xv, const float* g, const float* H_in,
                      float* H_out, float* C, float* head_part, long rows, int h, cudaStream_t st, float* gates_out) {
  const char* base = static_cast<const char*>(packed);
  const int tiles_u = simt_gate_tiles(h);
  const long row_tiles = (rows + kSgBM - 1) / kSgBM;
  const long grid = row_tiles * tiles_u;
  if (grid > 0x7fffffffL) IADMM_FAIL(IADMM_ESHAPE, "gate grid too large (%ld tiles)", grid);
  gates_simt_kernel<<<(unsigned)grid, kSgThreads, 0, st>>>(
      reinterpret_cast<const float*>(base + L.off_u32), reinterpret_cast<const float*>(base + L.off_wc),
      reinterpret_cast<const float*>(base + L.off_bias), reinterpret_cast<const float*>(base + L.off_wh), xv, g, H_in,
      H_out, C, head_part, gates_out, rows, h, tiles_u);
  IADMM_LAUNCH_CHECK("gates_simt_kernel");
  return IADMM_OK;
}

}  // namespace iadmm
