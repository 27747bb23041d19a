// Device helpers shared by the tensor-core LSTM-cell kernels (gates_tc.cu, resident.cu): the epilogue's
// transcendentals, 256-bit global accesses and the fp16 / e4m3 operand images of the hidden state.
#pragma once

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace iadmm {

// Transcendentals of the epilogue: MUFU ex2/rcp based, relative error ~2e-7 (measured against fp64 on the
// host for the polynomial; the gate-GEMM split error and fp32 summation order are larger).
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) {            // 1 / (1 + 2^(-x log2 e))
  return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f));
}
// tanh: odd minimax polynomial x + x^3 q(x^2) for |x| < 0.55 (rel err 1e-7 in fp32), 1 - 2/(e^{2x}+1) beyond;
// both evaluated, selected without a branch.
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = x * x;
  float q = fmaf(t, 0.016433170300270403f, -0.052669384762106176f);
  q = fmaf(q, t, 0.133206865150314f);
  q = fmaf(q, t, -0.33332945121698027f);
  const float small = fmaf(x * t, q, x);
  const float big = fmaf(-2.0f, rcp_approx(ex2_approx(x * 2.8853900817779268f) + 1.0f), 1.0f);
  return (fabsf(x) < 0.55f) ? small : big;
}

// exp-only tanh, 1 - 2 / (e^{2x} + 1): 5 instructions, ABSOLUTE error <= 3e-7 (the relative error grows as x -> 0,
// where the cancellation sits).  Used by the on-chip-resident kernel, which is instruction-issue bound in its
// epilogue; the streaming kernels keep the branch-free polynomial/exp pair above.
__device__ __forceinline__ float tanh_exp(float x) {
  return fmaf(-2.0f, rcp_approx(ex2_approx(x * 2.8853900817779268f) + 1.0f), 1.0f);
}

// The four gate activations of one hidden unit with ONE reciprocal: sigmoid(p) = 1/(1+e^-p), tanh(p) = 1 - 2/(1+e^2p);
// 1/a, 1/b, 1/c, 1/d follow from r = 1/(abcd) by multiplications.  The special-function unit (16 lanes/clk/SM) is
// the bottleneck of the resident kernel's epilogue: this form needs 4 ex2 + 1 rcp instead of 4 + 4.  Exponent
// arguments are clamped at 30 so the product stays below 2^121 (sigmoid(-20.8) = 9e-10 is returned for anything
// smaller: an absolute error below 1e-9).  Relative error ~5 ulp.
__device__ __forceinline__ void gates4_shared_rcp(float pi, float pf, float po, float pu, float& gi, float& gf, float& go, float& gu) {
  const float kL = 1.4426950408889634f;
  const float a = 1.0f + ex2_approx(fminf(pi * -kL, 30.0f));
  const float b = 1.0f + ex2_approx(fminf(pf * -kL, 30.0f));
  const float c = 1.0f + ex2_approx(fminf(po * -kL, 30.0f));
  const float d = 1.0f + ex2_approx(fminf(pu * (2.0f * kL), 30.0f));
  const float ab = a * b, cd = c * d;
  const float r = rcp_approx(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  gi = rab * b; gf = rab * a; go = rcd * d;
  gu = fmaf(-2.0f, rcd * c, 1.0f);
}

// Same with the pre-activations delivered as packed pairs (p_i, p_f) and (p_o, p_u) and the multiplies / adds issued two-wide
// (FMUL2 / FADD2): lane by lane the same operations as gates4_shared_rcp, 15 instead of 22 issue slots.
__device__ __forceinline__ void gates4_shared_rcp_x2(u64 pif, u64 pou, float& gi, float& gf, float& go, float& gu) {
  const float kL = 1.4426950408889634f;
  float t0, t1, t2, t3;
  upk2(mul2(pif, bc2(-kL)), t0, t1);
  upk2(mul2(pou, pk2(-kL, 2.0f * kL)), t2, t3);
  float a, b, c, d;
  upk2(add2(pk2(ex2_approx(fminf(t0, 30.0f)), ex2_approx(fminf(t1, 30.0f))), bc2(1.0f)), a, b);
  upk2(add2(pk2(ex2_approx(fminf(t2, 30.0f)), ex2_approx(fminf(t3, 30.0f))), bc2(1.0f)), c, d);
  const float ab = a * b, cd = c * d;
  const float r = rcp_approx(ab * cd);
  float rab, rcd;
  upk2(mul2(bc2(r), pk2(cd, ab)), rab, rcd);
  upk2(mul2(bc2(rab), pk2(b, a)), gi, gf);
  float guh;
  upk2(mul2(bc2(rcd), pk2(d, c)), go, guh);
  gu = fmaf(-2.0f, guh, 1.0f);
}

// 256-bit global accesses (sm_100): one request per 32-byte sector instead of two
__device__ __forceinline__ void ld_global_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_global_v8u(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// fp16 / e4m3 images of 8 hidden values for the next iteration's MMAs:
//   hi  = fp16(H*2^14)                                   (4 x half2)
//   lo  = fp16(H*2^14 - hi)                              (NPROD 3)
//   res = e4m3((H*2^14 - hi) * 2^5), crs = e4m3(H*2^14 * 2^-6)   (NPROD 2; 2 words each)
template <int NPROD>
__device__ __forceinline__ void split_hidden8(const float (&hnew)[8], uint32_t (&hi)[4], uint32_t (&lo)[4], uint32_t (&res)[2],
                                              uint32_t (&crs)[2]) {
  const float hs = (float)(1 << kHShift);
  __nv_fp8x2_storage_t r2[4], c2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float s0 = hnew[2 * u] * hs, s1 = hnew[2 * u + 1] * hs;
    const __half2 hh = __floats2half2_rn(s0, s1);
    hi[u] = *reinterpret_cast<const uint32_t*>(&hh);
    const float2 back = __half22float2(hh);
    if (NPROD == 3) {
      const __half2 hl = __floats2half2_rn(s0 - back.x, s1 - back.y);
      lo[u] = *reinterpret_cast<const uint32_t*>(&hl);
    }
    if (NPROD == 2) {
      r2[u] = __nv_cvt_float2_to_fp8x2(make_float2((s0 - back.x) * 32.0f, (s1 - back.y) * 32.0f), __NV_SATFINITE, __NV_E4M3);
      c2[u] = __nv_cvt_float2_to_fp8x2(make_float2(s0 * 0.015625f, s1 * 0.015625f), __NV_SATFINITE, __NV_E4M3);
    }
  }
  if (NPROD == 2) {
    res[0] = (uint32_t)r2[0] | ((uint32_t)r2[1] << 16); res[1] = (uint32_t)r2[2] | ((uint32_t)r2[3] << 16);
    crs[0] = (uint32_t)c2[0] | ((uint32_t)c2[1] << 16); crs[1] = (uint32_t)c2[2] | ((uint32_t)c2[3] << 16);
  }
}

}  // namespace iadmm
