// Epilogue arithmetic shared by the tcgen05 kernels (GEMM, fused ConvFFN): bf16 packing and the fast
// GELU / SiLU forms evaluated on the accumulators.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace fvla {
namespace epi {

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// ---- epilogue activations -----------------------------------------------------------------------
// The XU (MUFU) pipe is the scarce resource of the small-K GEMM epilogues: ncu shows it 100 % busy at
// ~4 tanh/clk/SM (profiles/r01_gemm_gelu_k192_ncu.txt), i.e. 8192 cycles for a 128x256 tile whose MMAs
// take 1536-3072.  ex2+rcp is no better (two MUFU ops).  GELU is therefore evaluated two ways and the
// elements of a thread are split between them so that the XU and FMA pipes finish together:
//   3 of 8 elements:  h + h*tanh(h*Q(h^2))          1 MUFU + 7 FP32 slots   (|err| <= 2.5e-5 + 2^-11 rel)
//   5 of 8 elements:  h + h*th*R(th^2), th=clamp(h) 0 MUFU + 11 FP32 slots  (|err| <= 1.9e-4)
// both in terms of h = x/2 (GELU GEMMs are packed with weights/bias pre-halved, exact in bf16), both
// approximations of the erf form nn.GELU() computes.  The fp32 parity mode keeps erff.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh form: Phi(x) = (1 + tanh(P(x)))/2 with P fitted to atanh(erf(x/sqrt2)), h^2 clamped where tanh has
// saturated; polynomial form: erf(x/sqrt2) ~= x*R(x^2) on |x| <= 4 (odd minimax, 7 coefficients), +-1 outside.
// ---- packed fp32 pairs (sm_100 FMUL2 / FFMA2: two fp32 lanes per issue slot, same rounding as FMUL / FFMA) ----
// With warps decoupled the small-K epilogues are issue-bound (~16 instructions per output element at ~70 % issue
// utilisation), so the activation polynomials run on register pairs: 6.5 instead of 11 slots per element for the
// polynomial form, 4.5 instead of 7 for the tanh form.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 splat2(float c) { return pk2(c, c); }
__device__ __forceinline__ void gelu_half_mufu2(float& x0, float& x1) {
  const f32x2 h = pk2(x0, x1);
  float u0, u1;
  upk2(mul2(h, h), u0, u1);
  const f32x2 u = pk2(fminf(u0, 16.0f), fminf(u1, 16.0f));
  const f32x2 q = fma2(u, fma2(u, splat2(-1.124853725e-02f), splat2(2.96045168e-01f)), splat2(1.594015768f));
  float a0, a1;
  upk2(mul2(h, q), a0, a1);
  upk2(fma2(h, pk2(tanh_approx(a0), tanh_approx(a1)), h), x0, x1);
}
__device__ __forceinline__ void gelu_half_poly2(float& x0, float& x1) {
  const f32x2 h = pk2(x0, x1);
  const f32x2 th = pk2(fminf(fmaxf(x0, -2.0f), 2.0f), fminf(fmaxf(x1, -2.0f), 2.0f));
  const f32x2 s = mul2(th, th);
  f32x2 r = fma2(s, splat2(3.732471752e-04f), splat2(-6.547819094e-03f));
  r = fma2(s, r, splat2(4.910614436e-02f));
  r = fma2(s, r, splat2(-2.083871470e-01f));
  r = fma2(s, r, splat2(5.614317921e-01f));
  r = fma2(s, r, splat2(-1.033169161e+00f));
  r = fma2(s, r, splat2(1.591533285e+00f));
  upk2(fma2(h, mul2(th, r), h), x0, x1);
}
// 3 of every 8 pairs through the XU (tanh), 5 through the FMA pipe: both pipes finish together
template <int N>
__device__ __forceinline__ void gelu_half_hybrid(float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    if ((j & 7) % 3 == 0) gelu_half_mufu2(v[2 * j], v[2 * j + 1]);
    else gelu_half_poly2(v[2 * j], v[2 * j + 1]);
  }
}
// x * sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU (SwiGLU has one per TWO accumulators)
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}


// ---- GELU on packed half pairs, for hidden tensors that are stored as fp16 ---------------------------------------
// Where the consumer of a GELU output is another tensor-core GEMM of ours, the hidden tensor can be fp16 instead of
// bf16 (3 more mantissa bits; ConvFFN activations are O(1..100), far inside fp16 range; the consuming weights are
// converted to fp16 at pack time).  The whole activation then runs two elements per instruction:
//   h2 = cvt.f16x2(x0/2, x1/2);  u = min(h2*h2, 16);  q = c0 + u*(c1 + u*c2);  t = tanh.approx.f16x2(h2*q);
//   out = h2 + h2*t                                   -- 8 issue slots and ONE MUFU op per PAIR (fp32 form: 26 / 2)
// Same tanh-form fit as gelu_half_mufu2; tanh.approx.f16x2 carries ~2^-11 absolute error, the same order as the
// fp32 tanh.approx it replaces, and the fp16 result is stored without a second rounding.
__device__ __forceinline__ uint32_t h2_splat(float c) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(c));
  return r;
}
__device__ __forceinline__ uint32_t gelu_half_f16x2(float h_lo, float h_hi) {
  uint32_t h, u, q, a, t, o;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(h_hi), "f"(h_lo));
  const uint32_t c0 = h2_splat(1.594015768f), c1 = h2_splat(2.96045168e-01f), c2 = h2_splat(-1.124853725e-02f);
  const uint32_t k16 = h2_splat(16.0f);
  asm("mul.rn.f16x2 %0, %1, %1;" : "=r"(u) : "r"(h));
  asm("min.f16x2 %0, %1, %2;" : "=r"(u) : "r"(u), "r"(k16));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(u), "r"(c2), "r"(c1));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(u), "r"(q), "r"(c0));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(h), "r"(q));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(a));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(o) : "r"(h), "r"(t));
  return o;
}

// same, with the (pre-halved) bias added as a packed half pair after the conversion: 1 HADD2 per pair instead of 2 FADD
__device__ __forceinline__ uint32_t gelu_half_f16x2_b(float h_lo, float h_hi, uint32_t bias2) {
  uint32_t h, u, q, a, t, o;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(h_hi), "f"(h_lo));
  asm("add.rn.f16x2 %0, %0, %1;" : "+r"(h) : "r"(bias2));
  const uint32_t c0 = h2_splat(1.594015768f), c1 = h2_splat(2.96045168e-01f), c2 = h2_splat(-1.124853725e-02f);
  const uint32_t k16 = h2_splat(16.0f);
  asm("mul.rn.f16x2 %0, %1, %1;" : "=r"(u) : "r"(h));
  asm("min.f16x2 %0, %1, %2;" : "=r"(u) : "r"(u), "r"(k16));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(u), "r"(c2), "r"(c1));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(u), "r"(q), "r"(c0));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(h), "r"(q));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(a));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(o) : "r"(h), "r"(t));
  return o;
}

}  // namespace epi
}  // namespace fvla
