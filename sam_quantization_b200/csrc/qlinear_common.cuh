// Device helpers shared by the 1-CTA (qlinear.cu) and 2-CTA (qlinear2.cu) dequant-GEMM kernels.
#pragma once
#include "common.cuh"

namespace samq {
namespace {

// Exact-erf GELU, x * Phi(x), with erfc from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on
// erf):  Phi(x) = 1 - g (x >= 0) | g (x < 0),  g = 0.5 erfc(|x|/sqrt2) = poly(t) exp(-x^2/2),
// t = 1/(1 + p |x|/sqrt2).  gelu(x) = max(x, 0) - g |x|.  Max abs error 3.4e-7 over [-12, 12]
// (checked in tests/test_gelu_approx.py): far below the fp16 output rounding.  14 instructions
// (2 MUFU) instead of ~40 for erff, which made the lin1 epilogue issue-bound.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752440f, 1.0f)));
  float poly = 0.5f * 1.061405429f;
  poly = fmaf(poly, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  poly *= t;
  const float u = ax * 0.84932180028801904272f;   // |x| * sqrt(log2(e) / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * u));
  return fmaf(-(poly * e), ax, fmaxf(x, 0.f));
}

// (a & 0x000f000f) | 0x64006400  ->  two fp16 values 1024 + nibble
__device__ __forceinline__ uint32_t nib_to_h2(uint32_t w) {
  return lop3_and_or(w, 0x000f000fu, 0x64006400u);
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2_dup(__half h) {
  const uint32_t u = __half_as_ushort(h);
  return u | (u << 16);
}

}  // namespace
}  // namespace samq
