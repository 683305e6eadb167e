// Device helpers shared by the 1-CTA (qlinear.cu) and 2-CTA (qlinear2.cu) dequant-GEMM kernels.
#pragma once
#include "common.cuh"

namespace samq {

// window-order GEMM row -> image-order row (see EpiBlock below); ws == 0 means identity.
// Defined outside the anonymous namespace: it appears in cross-file launcher signatures.
struct RowMap {
  int ws, H, W, nH, nW;
  int to_windows;   // 0: GEMM rows are windowed tokens, stored in image order (pad tokens dropped);
                    // 1: GEMM rows are image-order tokens, stored at their windowed position
};

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

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction) ----
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_dup(float v) { return f32x2_pack(v, v); }
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// (a & 0x000f000f) | 0x64006400  ->  two fp16 values 1024 + nibble
__device__ __forceinline__ uint32_t nib_to_h2(uint32_t w) {
  return lop3_and_or(w, 0x000f000fu, 0x64006400u);
}
// (a & 0x00f000f0) | 0x54005400  ->  two fp16 values 64 + nibble (nibbles 1 and 5 of the word, unshifted)
__device__ __forceinline__ uint32_t nib_hi_to_h2(uint32_t w) {
  return lop3_and_or(w, 0x00f000f0u, 0x54005400u);
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

// ---- epilogue: one 32(token) x 32(feature) block per warp ------------------------------------
// Optional row remap = window_unpartition + crop fused into the store
// (image_encoder.py:201-204, 309-333): GEMM rows are tokens in window order
// [B*nH*nW, ws, ws]; the destination row is the token's place in [B, H, W] and padding tokens
// are dropped.  ws == 0 means identity.
__device__ __forceinline__ int map_row(int m, int M, const RowMap& rm) {
  if (m >= M) return -1;
  if (rm.ws == 0) return m;
  if (rm.to_windows) {
    // window_partition (image_encoder.py:282-306) fused into the store: token (b, h, w) goes to
    // window (h / ws, w / ws), position (h % ws, w % ws); the zero-padding rows are not touched
    const int w = m % rm.W, t = m / rm.W;
    const int h = t % rm.H, b = t / rm.H;
    const int wh = h / rm.ws, i = h - wh * rm.ws;
    const int ww = w / rm.ws, j = w - ww * rm.ws;
    return (((b * rm.nH + wh) * rm.nW + ww) * rm.ws + i) * rm.ws + j;
  }
  const int per_win = rm.ws * rm.ws;
  const int win = m / per_win, within = m - win * per_win;
  const int i = within / rm.ws, j = within - i * rm.ws;
  const int ww = win % rm.nW, t = win / rm.nW;
  const int wh = t % rm.nH, b = t / rm.nH;
  const int h = wh * rm.ws + i, w = ww * rm.ws + j;
  return (h < rm.H && w < rm.W) ? (b * rm.H + h) * rm.W + w : -1;
}

// Lanes own features (TMEM lanes), registers own tokens; the block is transposed through a
// 2 KB shared-memory tile so that global accesses are 16-byte row segments.
// (ncu source view of the first version: the residual "prefetch" was a predicated load followed
// by a predicated move, i.e. it waited for DRAM on the spot -- 21% of the proj GEMM's samples;
// the tile was addressed through generic pointers.  Hence the explicit predicated / shared PTX.)
// RES: 0 = the kernel has no residual, 1 = it always has one (unconditional loads: rows that are
// not stored read row 0 and are ignored), 2 = decided at run time (predicated loads).  ptxas turns
// a predicated load into "load to a temporary + predicated move", and the move waits for DRAM on
// the spot, so the kernels that live on the residual path are compiled with RES = 1.
template <bool GELU, int RES = 2>
struct EpiBlock {
  int dest[4];     // destination rows of the 4 row segments this lane stores (-1: skip)
  uint4 rv[4];     // residual values for them

  // before the TMEM load: destination rows (one map_row per lane, shuffled to where it is
  // needed) and the residual loads, whose latency then overlaps the load + math + transpose
  __device__ __forceinline__ void prefetch(int m0, int M, int N, int col0, int lane,
                                           const __half* residual, const RowMap& rm) {
    const int d = map_row(m0 + lane, M, rm);
    const int q = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      dest[it] = __shfl_sync(0xffffffffu, d, it * 8 + (lane >> 2));
      if (RES == 0) continue;
      const int row = dest[it] >= 0 ? dest[it] : 0;
      const __half* src = residual + static_cast<size_t>(row) * N + (col0 + q * 8);
      if (RES == 1) {
        asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rv[it].x), "=r"(rv[it].y), "=r"(rv[it].z), "=r"(rv[it].w)
                     : "l"(src));
      } else {
        rv[it] = make_uint4(0, 0, 0, 0);
        const int on = (residual != nullptr) & (dest[it] >= 0);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "@p ld.global.v4.u32 {%0, %1, %2, %3}, [%4];\n\t"
            "}\n"
            : "+r"(rv[it].x), "+r"(rv[it].y), "+r"(rv[it].z), "+r"(rv[it].w)
            : "l"(src), "r"(on));
      }
    }
  }

  __device__ __forceinline__ void finish(const uint32_t (&r)[32], float bv, __half* stage, int N,
                                         int col0, int lane, bool has_residual, __half* y) {
    const uint32_t st = smem_u32(stage);
    const uint32_t wr = st + lane * 2;
    if (!GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = __uint_as_float(r[j]) + bv;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(wr + j * 64), "h"(__half_as_ushort(__float2half_rn(v))) : "memory");
      }
    } else {
      // gelu_erf() over the 32 tokens, two tokens per instruction: Blackwell's packed fp32 pipe
      // (fma.rn.f32x2 = two IEEE fp32 FMAs per issue slot) halves the FMA-pipe instruction count
      // of the polynomial, which is what the lin1 GEMM loses tensor cycles to (the GELU warps
      // took 45 % of the SM's issue slots; the MMA-issuing warp was `not_selected` 22 % of its
      // samples).  Same formula as gelu_erf(); max(x, 0) is formed exactly as 0.5 |x| + 0.5 x so
      // that it needs no scalar max.  Software-pipelined by hand in three stages so that the four
      // MUFU ops of a pair (2 rcp, 2 ex2: one warp-instruction per 8 clk each) are spread between
      // the packed FMA-pipe instructions of its neighbours (volatile asm keeps the order).
      const uint64_t bv2 = f32x2_dup(bv);
      const uint64_t kP = f32x2_dup(0.3275911f * 0.70710678118654752440f), kOne = f32x2_dup(1.0f);
      const uint64_t kC4 = f32x2_dup(-0.5f * 1.061405429f), kC3 = f32x2_dup(-0.5f * -1.453152027f);
      const uint64_t kC2 = f32x2_dup(-0.5f * 1.421413741f), kC1 = f32x2_dup(-0.5f * -0.284496736f);
      const uint64_t kC0 = f32x2_dup(-0.5f * 0.254829592f), kHalf = f32x2_dup(0.5f);
      const uint64_t kNU2 = f32x2_dup(-0.72134752044448170368f);   // -log2(e) / 2:  exp(-x^2/2) = 2^(kNU2 x^2)
      uint64_t xv[16], av[16], tv[16], qv[16];
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        if (j < 16) {
          xv[j] = f32x2_add(f32x2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), bv2);
          av[j] = xv[j] & 0x7fffffff7fffffffull;                  // |x|
          float d0, d1, t0, t1;
          f32x2_unpack(f32x2_fma(av[j], kP, kOne), d0, d1);
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
          tv[j] = f32x2_pack(t0, t1);
        }
        if (j >= 1 && j < 17) {
          const int i = j - 1;
          const uint64_t t = tv[i];
          uint64_t poly = f32x2_fma(kC4, t, kC3);                 // -(0.5 erfc polynomial), Horner
          poly = f32x2_fma(poly, t, kC2);
          poly = f32x2_fma(poly, t, kC1);
          poly = f32x2_fma(poly, t, kC0);
          poly = f32x2_mul(poly, t);
          float w0, w1, e0, e1;
          f32x2_unpack(f32x2_mul(f32x2_mul(xv[i], xv[i]), kNU2), w0, w1);
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
          qv[i] = f32x2_mul(poly, f32x2_pack(e0, e1));            // -g
        }
        if (j >= 2) {
          const int i = j - 2;
          float g0, g1;
          const uint64_t relu = f32x2_fma(av[i], kHalf, f32x2_mul(xv[i], kHalf));   // exact
          f32x2_unpack(f32x2_fma(av[i], qv[i], relu), g0, g1);
          const __half2 h = __floats2half2_rn(g0, g1);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(wr + (2 * i) * 64), "h"(__half_as_ushort(__low2half(h))) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(wr + (2 * i + 1) * 64), "h"(__half_as_ushort(__high2half(h))) : "memory");
        }
      }
    }
    __syncwarp();
    const int q = lane & 3;
    const uint32_t rd = st + ((lane >> 2) * 32 + q * 8) * 2;
    uint4 val[4];
#pragma unroll
    for (int it = 0; it < 4; ++it)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(val[it].x), "=r"(val[it].y), "=r"(val[it].z), "=r"(val[it].w)
                   : "r"(rd + it * 512));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (dest[it] >= 0) {
        if (RES == 1 || (RES == 2 && has_residual)) {   // fp16 add of two fp16 values, as the reference's `shortcut + x`
          val[it].x = h2_add(val[it].x, rv[it].x);
          val[it].y = h2_add(val[it].y, rv[it].y);
          val[it].z = h2_add(val[it].z, rv[it].z);
          val[it].w = h2_add(val[it].w, rv[it].w);
        }
        *reinterpret_cast<uint4*>(y + static_cast<size_t>(dest[it]) * N + (col0 + q * 8)) = val[it];
      }
    }
    __syncwarp();
  }
};

// non-blocking probe of an mbarrier phase (the blocking try_wait costs ~90 clk even when the
// phase is already complete; probing early lets that latency overlap useful work)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---- in-register unpack of one k-block (64 k) of ONE output feature -> 32 fp16 pairs (k, k+1) ----
// `q` holds the thread's 2*BITS packed words of the k-block (GPTQ layout: fields LSB-first along k,
// 32/BITS per word; 3-bit: 32 fields as a 96-bit little-endian stream over 3 words, quant.py:160-180).
// Every format goes through the same two fp16 steps as dequant.cu, on pairs:
//   q = (1024 + q) - 1024  (exact: the field is OR-ed into the mantissa of 1024.0),
//   w = fma(q, s, -fp16((z+1) s))  -- the rounding of the reference's Triton kernel.
template <int BITS>
__device__ __forceinline__ void unpack_kblock(const uint32_t (&q)[2 * BITS], uint32_t s2, uint32_t nzs2,
                                              uint32_t (&out)[32]) {
  if constexpr (BITS == 4) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t w = q[r];
      // (k0,k4) (k1,k5) (k2,k6) (k3,k7) as fp16 pairs 1024+q (even nibbles) / 64+q (odd nibbles, in
      // place: no shift), minus the magic constant = q exactly, then ONE fma per pair
      const uint32_t w8 = w >> 8;
      uint32_t a = nib_to_h2(w), b = nib_hi_to_h2(w), c = nib_to_h2(w8), d = nib_hi_to_h2(w8);
      a = h2_fma(h2_add(a, 0xe400e400u), s2, nzs2);   // -1024
      b = h2_fma(h2_add(b, 0xd400d400u), s2, nzs2);   // -64
      c = h2_fma(h2_add(c, 0xe400e400u), s2, nzs2);
      d = h2_fma(h2_add(d, 0xd400d400u), s2, nzs2);
      out[4 * r + 0] = prmt(a, b, 0x5410);  // (k0,k1)
      out[4 * r + 1] = prmt(c, d, 0x5410);  // (k2,k3)
      out[4 * r + 2] = prmt(a, b, 0x7632);  // (k4,k5)
      out[4 * r + 3] = prmt(c, d, 0x7632);  // (k6,k7)
    }
  } else if constexpr (BITS == 8) {
    // a word is 4 consecutive k as bytes: prmt interleaves them with 0x64 -> fp16 1024 + byte
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const uint32_t lo = prmt(q[r], 0x64646464u, 0x4140), hi = prmt(q[r], 0x64646464u, 0x4342);
      out[2 * r + 0] = h2_fma(h2_add(lo, 0xe400e400u), s2, nzs2);   // (k0,k1)
      out[2 * r + 1] = h2_fma(h2_add(hi, 0xe400e400u), s2, nzs2);   // (k2,k3)
    }
  } else if constexpr (BITS == 2) {
    // a word is 16 consecutive k: (w >> 2j) & 0x00030003 = fields j and j + 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        h[j] = h2_fma(h2_add(lop3_and_or(q[r] >> (2 * j), 0x00030003u, 0x64006400u), 0xe400e400u), s2, nzs2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        out[8 * r + j] = prmt(h[2 * j], h[2 * j + 1], 0x5410);       // (k_2j, k_2j+1)
        out[8 * r + 4 + j] = prmt(h[2 * j], h[2 * j + 1], 0x7632);   // (k_8+2j, k_8+2j+1)
      }
    }
  } else {
    static_assert(BITS == 3, "bits");
    // 32 fields per 3 words; field i sits at bit 3 i of the 96-bit stream (fields 10 and 21 straddle)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        uint32_t f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int bit = 3 * (2 * pr + e), word = bit >> 5, off = bit & 31;
          uint32_t v = q[3 * c + word] >> off;
          if (off + 3 > 32) v |= q[3 * c + word + 1] << (32 - off);
          f[e] = v & 7u;
        }
        const uint32_t h = (f[0] | (f[1] << 16)) | 0x64006400u;
        out[16 * c + pr] = h2_fma(h2_add(h, 0xe400e400u), s2, nzs2);
      }
    }
  }
}

// The integer fields of one k-block (64 k) of one feature as EXACT fp16 pairs (q_k, q_k+1), before
// scale / zero are applied: the first half of unpack_kblock, for callers whose scale differs per k
// (act-order groups in the standalone unpack kernel, dequant.cu).
template <int BITS>
__device__ __forceinline__ void unpack_q_h2(const uint32_t (&q)[2 * BITS], uint32_t (&out)[32]) {
  if constexpr (BITS == 4) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t w = q[r], w8 = w >> 8;
      const uint32_t a = h2_add(nib_to_h2(w), 0xe400e400u), b = h2_add(nib_hi_to_h2(w), 0xd400d400u);
      const uint32_t c = h2_add(nib_to_h2(w8), 0xe400e400u), d = h2_add(nib_hi_to_h2(w8), 0xd400d400u);
      out[4 * r + 0] = prmt(a, b, 0x5410);
      out[4 * r + 1] = prmt(c, d, 0x5410);
      out[4 * r + 2] = prmt(a, b, 0x7632);
      out[4 * r + 3] = prmt(c, d, 0x7632);
    }
  } else if constexpr (BITS == 8) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      out[2 * r + 0] = h2_add(prmt(q[r], 0x64646464u, 0x4140), 0xe400e400u);
      out[2 * r + 1] = h2_add(prmt(q[r], 0x64646464u, 0x4342), 0xe400e400u);
    }
  } else if constexpr (BITS == 2) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = h2_add(lop3_and_or(q[r] >> (2 * j), 0x00030003u, 0x64006400u), 0xe400e400u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        out[8 * r + j] = prmt(h[2 * j], h[2 * j + 1], 0x5410);
        out[8 * r + 4 + j] = prmt(h[2 * j], h[2 * j + 1], 0x7632);
      }
    }
  } else {
    static_assert(BITS == 3, "bits");
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        uint32_t f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int bit = 3 * (2 * pr + e), word = bit >> 5, off = bit & 31;
          uint32_t v = q[3 * c + word] >> off;
          if (off + 3 > 32) v |= q[3 * c + word + 1] << (32 - off);
          f[e] = v & 7u;
        }
        out[16 * c + pr] = h2_add((f[0] | (f[1] << 16)) | 0x64006400u, 0xe400e400u);
      }
    }
  }
}

// One dequant warp's main loop (shared by the 1-CTA and 2-CTA kernels).
//
// `set` (0/1) takes the CTA-wide k-blocks kbc = set, set+2, ...  For each of them the warp
// (TMEM lane quadrant q4) reads its 8 packed words per thread from the weight ring, unpacks
// to fp16 (bit-identical to dequant.cu) and writes TMEM A stage kbc % kAStages.
//   n_of_tile(tl)  -> first output feature of this CTA in its tl-th tile
//   arrive_full(as) -> signal "A stage written" to the MMA issuer (local or remote barrier)
// No integer division on the per-k-block path: (tile, kb, group) are tracked incrementally
// (two dependent runtime divisions per k-block cost ~250 clk of latency in the first version).
template <int kWStages, int kAStages, int kWStageBytes, int BITS = 4, class NOfTile, class ArriveFull>
__device__ __forceinline__ void dequant_warp_loop(
    int set, int q4, int lane, int total_kb, int num_kb, int N, int groupsize,
    const __half* __restrict__ scales, const int32_t* __restrict__ qzeros, const uint8_t* sw,
    uint64_t* w_full, uint64_t* w_empty, uint64_t* a_empty, uint32_t tmem_a_base,
    NOfTile n_of_tile, ArriveFull arrive_full) {
  static_assert((kWStages & (kWStages - 1)) == 0 && (kAStages & (kAStages - 1)) == 0, "ring sizes");
  static_assert(kWStageBytes == 2 * BITS * 128 * 4, "packed tile = 2*BITS words x 128 features");
  const int tid = q4 * 32 + lane;  // 0..127 == TMEM lane == feature within the CTA's tile
  // zero point of feature n: field n of the qzeros row (same packing as the weights, along n).
  // A tile starts at a multiple of 128 features, so the field's place inside its word (or, for
  // 3 bits, inside its 3-word chunk) depends on tid only.
  constexpr int kFieldsPerWord = BITS == 3 ? 1 : 32 / BITS;
  const int zwords = BITS == 3 ? N / 32 * 3 : N / kFieldsPerWord;
  const int zbit = BITS == 3 ? 3 * (tid & 31) : (tid % kFieldsPerWord) * BITS;
  const int zword_in_tile = BITS == 3 ? (tid >> 5) * 3 + (zbit >> 5) : tid / kFieldsPerWord;
  const int zshift = zbit & 31;
  const bool zstraddle = BITS == 3 && zshift + 3 > 32;
  const int kb_per_group = groupsize / 64;

  // cursor of the k-block whose constants are being PREFETCHED (one own-k-block ahead)
  int p_tl = 0, p_kb = set, p_g = 0, p_kig = set;   // tile, kb in tile, group, kb in group
  auto normalise = [&]() {
    while (p_kb >= num_kb) { p_kb -= num_kb; ++p_tl; p_g = 0; p_kig = p_kb; }
    while (p_kig >= kb_per_group) { p_kig -= kb_per_group; ++p_g; }
  };
  normalise();
  __half s_next = __float2half(0.f);
  uint32_t zw_next = 0, zw_next_hi = 0;
  auto prefetch = [&]() {
    const int n0 = n_of_tile(p_tl);
    s_next = scales[static_cast<int64_t>(p_g) * N + n0 + tid];
    const int32_t* zp = qzeros + static_cast<int64_t>(p_g) * zwords +
                        (BITS == 3 ? n0 / 32 * 3 : n0 / kFieldsPerWord) + zword_in_tile;
    zw_next = static_cast<uint32_t>(zp[0]);
    if (BITS == 3 && zstraddle) zw_next_hi = static_cast<uint32_t>(zp[1]);
  };
  if (set < total_kb) prefetch();

  for (int kbc = set; kbc < total_kb; kbc += 2) {
    const int ws = kbc & (kWStages - 1);
    const uint32_t wph = (kbc / kWStages) & 1;
    const int as = kbc & (kAStages - 1);
    const uint32_t aph = (kbc / kAStages) & 1;
    // early, non-blocking probes; their latency hides behind the constant setup / unpack
    const bool w_ready = mbar_test(&w_full[ws], wph);
    const bool a_ready = mbar_test(&a_empty[as], aph ^ 1);

    const __half s = s_next;
    uint32_t z = zw_next >> zshift;
    if (BITS == 3 && zstraddle) z |= zw_next_hi << (32 - zshift);
    z &= (1u << BITS) - 1u;
    const __half zs = __hmul_rn(__uint2half_rn(z + 1u), s);
    const uint32_t s2 = h2_dup(s);
    const uint32_t nzs2 = h2_dup(__hneg(zs));
    if (kbc + 2 < total_kb) {   // raw loads for the next own k-block: nothing depends on them yet
      p_kb += 2;
      p_kig += 2;
      normalise();
      prefetch();
    }

    if (!w_ready) mbar_wait(&w_full[ws], wph);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(sw + ws * kWStageBytes) + tid;
    uint32_t q[2 * BITS];
#pragma unroll
    for (int r = 0; r < 2 * BITS; ++r) q[r] = wp[r * 128];
    __syncwarp();
    if (lane == 0) mbar_arrive(&w_empty[ws]);

    uint32_t out[32];
    unpack_kblock<BITS>(q, s2, nzs2, out);
    if (!a_ready) mbar_wait(&a_empty[as], aph ^ 1);   // MMAs that read this A stage are done
    tc_fence_after();
    tmem_st_x32(tmem_a_base + as * 32 + (static_cast<uint32_t>(q4 * 32) << 16), out);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive_full(as);
  }
}

}  // namespace
}  // namespace samq
