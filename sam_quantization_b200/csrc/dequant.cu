// Standalone unpack + dequantise of GPTQ-packed weights (HBM-bound).
//
// Arithmetic follows the reference kernel literally (quant_linear.py:334-339):
//     zeros = (z + 1) * scales         -> fp16 rounding
//     b     = q * scales - zeros       -> fp16 rounding after the product and after
//                                         the subtraction (no FMA contraction)
// which is what PyTorch produces for the same expression on fp16 tensors; the
// oracle (oracle/quant.py, dequant form "stepwise") is the bit-exact checker.
//
// Packing (gptq4sam.py:472-495): 2/4/8-bit fields LSB-first, 32/bits consecutive
// k per int32 of qweight[k/f, n]; 32/bits consecutive n per int32 of
// qzeros[g, n/f] storing zero-1.  3-bit (extension, quant.py:160-180): 32 values
// form a 96-bit little-endian bit stream over 3 consecutive words.
#include "common.cuh"

namespace samq {

// field `idx` of a packed bit stream whose words are `stride` int32 apart
__device__ __forceinline__ uint32_t extract_field(const int32_t* __restrict__ base, int64_t stride,
                                                  int idx, int bits) {
  const int chunk = idx >> 5, j = idx & 31;
  const int p = bits * j;
  const int word = p >> 5, off = p & 31;
  const int64_t w = static_cast<int64_t>(chunk) * bits + word;
  uint32_t v = static_cast<uint32_t>(base[w * stride]) >> off;
  if (off + bits > 32) v |= static_cast<uint32_t>(base[(w + 1) * stride]) << (32 - off);
  return v & ((1u << bits) - 1u);
}

__device__ __forceinline__ __half dequant_one(uint32_t q, __half s, __half zs) {
  // fp16(fp16(q*s) - zs); *_rn intrinsics forbid contraction into an FMA
  return __hsub_rn(__hmul_rn(__uint2half_rn(q), s), zs);
}

// One thread: 32 consecutive k (one packing chunk) of 2 adjacent output features.
// blockDim.x threads cover 2*blockDim.x columns; blockIdx.y = k chunk.
template <int BITS, bool TRANSPOSED>
__global__ void __launch_bounds__(128)
unpack_dequant_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                      const __half* __restrict__ scales, const int32_t* __restrict__ g_idx,
                      __half* __restrict__ w_out, int K, int N, int groupsize) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int k0 = blockIdx.y * 32;
  if (n >= N) return;
  const int kcount = min(32, K - k0);
  const int zstride_words = (N * BITS) / 32;  // words per qzeros row

  // packed words of this chunk for columns n, n+1 (coalesced 8-byte loads)
  uint32_t w0[BITS], w1[BITS];
  const int rows_total = (K * BITS + 31) / 32;
#pragma unroll
  for (int r = 0; r < BITS; ++r) {
    const int row = blockIdx.y * BITS + r;
    if (row < rows_total) {
      const int2 v = *reinterpret_cast<const int2*>(qweight + static_cast<int64_t>(row) * N + n);
      w0[r] = static_cast<uint32_t>(v.x);
      w1[r] = static_cast<uint32_t>(v.y);
    } else {
      w0[r] = 0;
      w1[r] = 0;
    }
  }

  int last_g = -1;
  __half s0 = __float2half(0.f), s1 = s0, zs0 = s0, zs1 = s0;
  __align__(16) __half out0[32];
  __align__(16) __half out1[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < kcount) {
      const int k = k0 + j;
      const int g = g_idx ? g_idx[k] : k / groupsize;
      if (g != last_g) {
        last_g = g;
        const __half2 sv = *reinterpret_cast<const __half2*>(scales + static_cast<int64_t>(g) * N + n);
        s0 = __low2half(sv);
        s1 = __high2half(sv);
        const int32_t* zrow = qzeros + static_cast<int64_t>(g) * zstride_words;
        const uint32_t z0 = extract_field(zrow, 1, n, BITS);
        const uint32_t z1 = extract_field(zrow, 1, n + 1, BITS);
        zs0 = __hmul_rn(__uint2half_rn(z0 + 1u), s0);
        zs1 = __hmul_rn(__uint2half_rn(z1 + 1u), s1);
      }
      const int p = BITS * j;
      const int word = p >> 5, off = p & 31;
      uint32_t q0 = w0[word] >> off, q1 = w1[word] >> off;
      if (off + BITS > 32) {  // only the 3-bit stream straddles words
        q0 |= w0[(word + 1) % BITS] << (32 - off);
        q1 |= w1[(word + 1) % BITS] << (32 - off);
      }
      q0 &= (1u << BITS) - 1u;
      q1 &= (1u << BITS) - 1u;
      out0[j] = dequant_one(q0, s0, zs0);
      out1[j] = dequant_one(q1, s1, zs1);
    } else {
      out0[j] = __float2half(0.f);
      out1[j] = out0[j];
    }
  }

  if (!TRANSPOSED) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < kcount)
        *reinterpret_cast<__half2*>(w_out + static_cast<int64_t>(k0 + j) * N + n) =
            __halves2half2(out0[j], out1[j]);
  } else {
    __half* r0 = w_out + static_cast<int64_t>(n) * K + k0;
    __half* r1 = r0 + K;
    if (kcount == 32) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        *reinterpret_cast<uint4*>(r0 + v * 8) = *reinterpret_cast<const uint4*>(out0 + v * 8);
        *reinterpret_cast<uint4*>(r1 + v * 8) = *reinterpret_cast<const uint4*>(out1 + v * 8);
      }
    } else {
      for (int j = 0; j < kcount; ++j) {
        r0[j] = out0[j];
        r1[j] = out1[j];
      }
    }
  }
}

template <int BITS>
static int launch_unpack(const int32_t* qweight, const int32_t* qzeros, const __half* scales,
                         const int32_t* g_idx, __half* w_out, int K, int N, int groupsize,
                         int transposed, cudaStream_t st) {
  dim3 block(128);
  dim3 grid((N / 2 + 127) / 128, (K + 31) / 32);
  if (transposed)
    unpack_dequant_kernel<BITS, true><<<grid, block, 0, st>>>(qweight, qzeros, scales, g_idx, w_out, K, N, groupsize);
  else
    unpack_dequant_kernel<BITS, false><<<grid, block, 0, st>>>(qweight, qzeros, scales, g_idx, w_out, K, N, groupsize);
  count_launch();
  return check_launch("unpack_dequant_kernel");
}

int unpack_dequant(const int32_t* qweight, const int32_t* qzeros, const void* scales,
                   const int32_t* g_idx, void* w_out, int K, int N, int bits, int groupsize,
                   int transposed, cudaStream_t st) {
  SAMQ_REQUIRE(qweight && qzeros && scales && w_out, SAMQ_ERR_BAD_ARG, "unpack_dequant: null pointer");
  SAMQ_REQUIRE(bits == 2 || bits == 3 || bits == 4 || bits == 8, SAMQ_ERR_UNSUPPORTED_BITS,
               "unpack_dequant: bits must be 2, 3, 4 or 8 (got %d)", bits);
  SAMQ_REQUIRE(K > 0 && N > 0, SAMQ_ERR_BAD_SHAPE, "unpack_dequant: K=%d N=%d", K, N);
  if (groupsize == -1) groupsize = K;
  SAMQ_REQUIRE(groupsize > 0, SAMQ_ERR_BAD_SHAPE, "unpack_dequant: groupsize=%d", groupsize);
  SAMQ_REQUIRE(K % 32 == 0, SAMQ_ERR_BAD_SHAPE, "unpack_dequant: K=%d must be a multiple of 32", K);
  SAMQ_REQUIRE(N % 32 == 0, SAMQ_ERR_BAD_SHAPE, "unpack_dequant: N=%d must be a multiple of 32", N);
  if (transposed)
    SAMQ_REQUIRE(reinterpret_cast<uintptr_t>(w_out) % 16 == 0, SAMQ_ERR_BAD_ARG,
                 "unpack_dequant: w_out must be 16-byte aligned");
  const __half* s = reinterpret_cast<const __half*>(scales);
  __half* w = reinterpret_cast<__half*>(w_out);
  switch (bits) {
    case 2: return launch_unpack<2>(qweight, qzeros, s, g_idx, w, K, N, groupsize, transposed, st);
    case 3: return launch_unpack<3>(qweight, qzeros, s, g_idx, w, K, N, groupsize, transposed, st);
    case 4: return launch_unpack<4>(qweight, qzeros, s, g_idx, w, K, N, groupsize, transposed, st);
    default: return launch_unpack<8>(qweight, qzeros, s, g_idx, w, K, N, groupsize, transposed, st);
  }
}

}  // namespace samq

extern "C" int samq_unpack_dequant(const int32_t* qweight, const int32_t* qzeros,
                                   const void* scales, const int32_t* g_idx, void* w_out, int K,
                                   int N, int bits, int groupsize, int transposed, void* stream) {
  return samq::unpack_dequant(qweight, qzeros, scales, g_idx, w_out, K, N, bits, groupsize,
                              transposed, reinterpret_cast<cudaStream_t>(stream));
}
