// HBM-bound token-wise kernels of Block.forward (image_encoder.py:189-207):
// LayerNorm, LayerNorm fused with window_partition (zero padding applied after
// the norm, image_encoder.py:191-195, 282-306), window_unpartition + crop +
// residual add (image_encoder.py:201-204, 309-333) and a plain residual add.
// One warp per token, 16-byte vector accesses, fp32 statistics (two-pass over
// registers: mean, then sum of squared deviations).
#include "common.cuh"

namespace samq {
namespace {

constexpr int kMaxVec = 8;  // 16-byte vectors per lane -> C <= 2048

struct alignas(16) H8 {
  __half2 v[4];
};

// normalise one token held by a warp; `src`/`dst` point at the token's C halfs
__device__ __forceinline__ void warp_layernorm(const __half* __restrict__ src,
                                               const __half* __restrict__ gamma,
                                               const __half* __restrict__ beta,
                                               __half* __restrict__ dst, int C, float eps,
                                               int lane) {
  const int nvec = C >> 3;  // 16-byte vectors per token
  float x[kMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const H8 h = *reinterpret_cast<const H8*>(src + v * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h.v[j]);
        x[i][2 * j] = f.x;
        x[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(C);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[i][j] - mean;
        var += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const H8 g = *reinterpret_cast<const H8*>(gamma + v * 8);
      const H8 b = *reinterpret_cast<const H8*>(beta + v * 8);
      H8 o;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 gf = __half22float2(g.v[j]);
        const float2 bf = __half22float2(b.v[j]);
        const float y0 = (x[i][2 * j] - mean) * rstd * gf.x + bf.x;
        const float y1 = (x[i][2 * j + 1] - mean) * rstd * gf.y + bf.y;
        o.v[j] = __floats2half2_rn(y0, y1);
      }
      *reinterpret_cast<H8*>(dst + v * 8) = o;
    }
  }
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, __half* __restrict__ y, int64_t rows, int C,
                 float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  warp_layernorm(x + row * C, gamma, beta, y + row * C, C, eps, lane);
}

// output token r of [B*nH*nW, ws, ws, C]  <-  input token (b, h, w) or zeros
__global__ void __launch_bounds__(256)
layernorm_partition_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                           const __half* __restrict__ beta, __half* __restrict__ y, int B, int H,
                           int W, int C, int ws, int nH, int nW, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = static_cast<int64_t>(B) * nH * nW * ws * ws;
  if (r >= total) return;
  const int j = static_cast<int>(r % ws);
  int64_t q = r / ws;
  const int i = static_cast<int>(q % ws);
  q /= ws;
  const int ww = static_cast<int>(q % nW);
  q /= nW;
  const int wh = static_cast<int>(q % nH);
  const int b = static_cast<int>(q / nH);
  const int h = wh * ws + i, w = ww * ws + j;
  __half* dst = y + r * C;
  if (h < H && w < W) {
    const __half* src = x + ((static_cast<int64_t>(b) * H + h) * W + w) * C;
    warp_layernorm(src, gamma, beta, dst, C, eps, lane);
  } else {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int v = lane; v < (C >> 3); v += 32) *reinterpret_cast<uint4*>(dst + v * 8) = z;
  }
}

// one thread per 8 channels of one output token
__global__ void __launch_bounds__(256)
unpartition_residual_kernel(const __half* __restrict__ windows, const __half* shortcut,
                            __half* out, int B, int H, int W, int C, int ws, int nH, int nW) {
  const int cv = C >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * H * W * cv;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % cv);
  int64_t tkn = idx / cv;
  const int w = static_cast<int>(tkn % W);
  tkn /= W;
  const int h = static_cast<int>(tkn % H);
  const int b = static_cast<int>(tkn / H);
  const int64_t win = (static_cast<int64_t>(b) * nH + h / ws) * nW + w / ws;
  const int64_t src_tok = (win * ws + h % ws) * ws + w % ws;
  const H8 a = *reinterpret_cast<const H8*>(windows + src_tok * C + v * 8);
  const int64_t o = ((static_cast<int64_t>(b) * H + h) * W + w) * C + v * 8;
  const H8 s = *reinterpret_cast<const H8*>(shortcut + o);
  H8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = __hadd2_rn(s.v[j], a.v[j]);
  *reinterpret_cast<H8*>(out + o) = r;
}

__global__ void __launch_bounds__(256)
add_kernel(const __half* a, const __half* b, __half* out, int64_t nvec) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const H8 x = *reinterpret_cast<const H8*>(a + i * 8);
  const H8 y = *reinterpret_cast<const H8*>(b + i * 8);
  H8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) r.v[j] = __hadd2_rn(x.v[j], y.v[j]);
  *reinterpret_cast<H8*>(out + i * 8) = r;
}

// ---- fast path: C == NVEC * 256 (768 / 1024 / 1280 ...), two tokens per warp ----------
// Each lane keeps NVEC 16-byte vectors per token in registers; both tokens' loads are
// issued before any arithmetic, so a warp has 2 * NVEC * 512 B in flight (the generic kernel
// above had half of that and twice the registers, and reached only ~40 % of HBM bandwidth).
//
// A token is converted to fp32 ONCE (the fp32 pairs replace its raw vectors in registers) and every
// arithmetic pass runs on the packed fp32 pipe (add / mul / fma.rn.f32x2: two IEEE operations per
// issue slot, same rounding as the scalar forms); the operation order per element is the reference's,
// ((x - mean) * rstd) * g + b.  7.5 instead of 12 instructions per element.  Measured in situ (behind
// the power-capped GEMMs, SM clock 1.3-1.4 GHz; A/B on one box): 119.8 -> 116.0 us per 131072 x 1280
// call -- the kernel is NOT issue-bound there (a version that also staged gamma / beta as fp32 in
// shared memory, 5.5 instructions per element but one block barrier, took 126 us); what it runs at
// in situ (5.8 TB/s) is what the memory system delivers at that clock.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpk2(uint64_t v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int NVEC>
__device__ __forceinline__ void ln_token_from_regs(const uint4 (&raw)[NVEC], const __half* __restrict__ gamma,
                                                   const __half* __restrict__ beta, __half* __restrict__ dst,
                                                   float inv_c, float eps, int lane) {
  uint64_t f[NVEC][4];                       // the token's values as fp32 pairs
  uint64_t s2 = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = __half22float2(h[j]);
      f[i][j] = pk2(v.x, v.y);
      s2 = add2(s2, f[i][j]);
    }
  }
  const float2 sp = unpk2(s2);
  float sum = sp.x + sp.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * inv_c;
  const uint64_t nm2 = pk2(-mean, -mean);
  uint64_t v2 = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[i][j] = add2(f[i][j], nm2);          // x - mean
      v2 = fma2(f[i][j], f[i][j], v2);
    }
  const float2 vp = unpk2(v2);
  float var = vp.x + vp.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var * inv_c + eps);
  const uint64_t r2 = pk2(rstd, rstd);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int v = lane + i * 32;
    const uint4 g4 = *reinterpret_cast<const uint4*>(gamma + v * 8);
    const uint4 b4 = *reinterpret_cast<const uint4*>(beta + v * 8);
    const __half2* g = reinterpret_cast<const __half2*>(&g4);
    const __half2* b = reinterpret_cast<const __half2*>(&b4);
    uint4 o4;
    __half2* o = reinterpret_cast<__half2*>(&o4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 gf = __half22float2(g[j]);
      const float2 bf = __half22float2(b[j]);
      const float2 r = unpk2(fma2(mul2(f[i][j], r2), pk2(gf.x, gf.y), pk2(bf.x, bf.y)));
      o[j] = __floats2half2_rn(r.x, r.y);
    }
    *reinterpret_cast<uint4*>(dst + v * 8) = o4;
  }
}

// PARTITION == false: rows are tokens of x;  true: rows are OUTPUT tokens of the windowed
// layout [B*nH*nW, ws, ws, C] (source token looked up, padding written as zeros).
// Three resident blocks per SM for C <= 1280 (<= 80 registers): in situ, behind the power-capped GEMMs at
// ~1.35 GHz, the reduce / normalise phase of a warp takes 35 % longer than at the 1.8 GHz ncu sees, and 16
// warps per SM no longer cover it (A/B on one box: 131 -> 118 us per 131072 x 1280 call, + 0.4 % of the step;
// one token per warp at four blocks per SM: 124 us).
template <int NVEC, bool PARTITION>
__global__ void __launch_bounds__(256, NVEC <= 5 ? 3 : 2)
layernorm_fast_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                      const __half* __restrict__ beta, __half* __restrict__ y, int64_t rows, int B,
                      int H, int W, int ws, int nH, int nW, float eps) {
  constexpr int C = NVEC * 256;
  const int lane = threadIdx.x & 31;
  // Rows are walked from the END of the tensor (block 0 takes the last rows).  The producer of x -- a
  // GEMM epilogue -- wrote the tensor front to back, so its last ~100 MB are still in the 126 MB L2
  // when this kernel starts; and the consumer of y (the next GEMM) starts at row 0, which this
  // kernel then writes last.  A 335 MB activation does not fit in L2, its two ends do.
  const int64_t blk = PARTITION ? static_cast<int64_t>(blockIdx.x) : static_cast<int64_t>(gridDim.x) - 1 - blockIdx.x;
  const int64_t r0 = (blk * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  const __half* src[2];
  bool live[2], pad[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int64_t r = r0 + t;
    live[t] = r < rows;
    pad[t] = false;
    src[t] = x;
    if (live[t]) {
      if (PARTITION) {
        const int j = static_cast<int>(r % ws);
        int64_t q = r / ws;
        const int i = static_cast<int>(q % ws);
        q /= ws;
        const int ww = static_cast<int>(q % nW);
        q /= nW;
        const int wh = static_cast<int>(q % nH);
        const int b = static_cast<int>(q / nH);
        const int h = wh * ws + i, w = ww * ws + j;
        pad[t] = !(h < H && w < W);
        if (!pad[t]) src[t] = x + ((static_cast<int64_t>(b) * H + h) * W + w) * C;
      } else {
        src[t] = x + r * C;
      }
    }
  }
  uint4 raw[2][NVEC];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < NVEC; ++i)
      raw[t][i] = (live[t] && !pad[t]) ? *reinterpret_cast<const uint4*>(src[t] + (lane + i * 32) * 8)
                                       : make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    if (!live[t]) continue;
    __half* dst = y + (r0 + t) * C;
    if (pad[t]) {
#pragma unroll
      for (int i = 0; i < NVEC; ++i) *reinterpret_cast<uint4*>(dst + (lane + i * 32) * 8) = make_uint4(0, 0, 0, 0);
    } else {
      ln_token_from_regs<NVEC>(raw[t], gamma, beta, dst, 1.0f / C, eps, lane);
    }
  }
}

template <bool PARTITION>
bool launch_ln_fast(const __half* x, const __half* g, const __half* b, __half* y, int64_t rows, int C, int B,
                    int H, int W, int ws, int nH, int nW, float eps, cudaStream_t st) {
  if (C % 256 != 0 || C > 2048) return false;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + 2 * wpb - 1) / (2 * wpb));
#define SAMQ_LN_CASE(NV)                                                                              \
  case NV:                                                                                            \
    layernorm_fast_kernel<NV, PARTITION><<<grid, wpb * 32, 0, st>>>(x, g, b, y, rows, B, H, W, ws, nH, nW, eps); \
    return true;
  switch (C / 256) {
    SAMQ_LN_CASE(1) SAMQ_LN_CASE(2) SAMQ_LN_CASE(3) SAMQ_LN_CASE(4) SAMQ_LN_CASE(5) SAMQ_LN_CASE(6)
    SAMQ_LN_CASE(7) SAMQ_LN_CASE(8)
  }
#undef SAMQ_LN_CASE
  return false;
}

int check_ln(const void* x, const void* g, const void* b, const void* y, int C, const char* who) {
  SAMQ_REQUIRE(x && g && b && y, SAMQ_ERR_BAD_ARG, "%s: null pointer", who);
  SAMQ_REQUIRE(C > 0 && C % 8 == 0 && C <= kMaxVec * 256, SAMQ_ERR_BAD_SHAPE,
               "%s: C=%d must be a multiple of 8 and <= %d", who, C, kMaxVec * 256);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) |
                reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "%s: pointers must be 16-byte aligned", who);
  return SAMQ_OK;
}

}  // namespace
}  // namespace samq

extern "C" int samq_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y,
                                  int64_t rows, int C, float eps, void* stream) {
  using namespace samq;
  int rc = check_ln(x, gamma, beta, y, C, "samq_layernorm_fwd");
  if (rc != SAMQ_OK) return rc;
  SAMQ_REQUIRE(rows >= 0, SAMQ_ERR_BAD_SHAPE, "samq_layernorm_fwd: rows=%lld", (long long)rows);
  if (rows == 0) return SAMQ_OK;
  if (launch_ln_fast<false>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
                            reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), rows, C, 0, 0,
                            0, 1, 1, 1, eps, reinterpret_cast<cudaStream_t>(stream))) {
    count_launch();
    return check_launch("layernorm_fast_kernel");
  }
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  layernorm_kernel<<<grid, wpb * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
      reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), rows, C, eps);
  count_launch();
  return check_launch("layernorm_kernel");
}

extern "C" int samq_layernorm_partition_fwd(const void* x, const void* gamma, const void* beta,
                                            void* y, int B, int H, int W, int C, int ws, float eps,
                                            void* stream) {
  using namespace samq;
  int rc = check_ln(x, gamma, beta, y, C, "samq_layernorm_partition_fwd");
  if (rc != SAMQ_OK) return rc;
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && ws > 0, SAMQ_ERR_BAD_SHAPE,
               "samq_layernorm_partition_fwd: B=%d H=%d W=%d ws=%d", B, H, W, ws);
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  const int64_t total = static_cast<int64_t>(B) * nH * nW * ws * ws;
  if (launch_ln_fast<true>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
                           reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), total, C, B, H, W,
                           ws, nH, nW, eps, reinterpret_cast<cudaStream_t>(stream))) {
    count_launch();
    return check_launch("layernorm_fast_kernel<partition>");
  }
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((total + wpb - 1) / wpb);
  layernorm_partition_kernel<<<grid, wpb * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
      reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), B, H, W, C, ws, nH, nW,
      eps);
  count_launch();
  return check_launch("layernorm_partition_kernel");
}

extern "C" int samq_unpartition_residual(const void* windows, const void* shortcut, void* out,
                                         int B, int H, int W, int C, int ws, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(windows && shortcut && out, SAMQ_ERR_BAD_ARG, "samq_unpartition_residual: null pointer");
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && ws > 0 && C > 0 && C % 8 == 0, SAMQ_ERR_BAD_SHAPE,
               "samq_unpartition_residual: B=%d H=%d W=%d C=%d ws=%d", B, H, W, C, ws);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(windows) | reinterpret_cast<uintptr_t>(shortcut) |
                reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "samq_unpartition_residual: pointers must be 16-byte aligned");
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  const int64_t total = static_cast<int64_t>(B) * H * W * (C / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  unpartition_residual_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(windows), reinterpret_cast<const __half*>(shortcut),
      reinterpret_cast<__half*>(out), B, H, W, C, ws, nH, nW);
  count_launch();
  return check_launch("unpartition_residual_kernel");
}

namespace samq {
namespace {
// Non-overlapping P x P patches of an NCHW image as GEMM rows (PatchEmbed's conv16x16/16,
// image_encoder.py:434-442, is exactly  rows x [C*P*P] . W^T):  out[(b,ph,pw), (c,i,j)] =
// x[b, c, ph*P+i, pw*P+j].  One thread per 16-byte chunk; consecutive threads walk an image
// row, so reads are contiguous and writes are whole 32-byte sectors.
__global__ void __launch_bounds__(256)
patchify_kernel(const __half* __restrict__ x, __half* __restrict__ out, int B, int C, int H, int W, int P) {
  const int cpr = W / 8;                       // 16-byte chunks per image row
  const int64_t total = static_cast<int64_t>(B) * C * H * cpr;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % cpr);
  int64_t t = idx / cpr;
  const int row = static_cast<int>(t % H);
  t /= H;
  const int c = static_cast<int>(t % C);
  const int b = static_cast<int>(t / C);
  const int col = chunk * 8;
  const int ph = row / P, i = row % P, pw = col / P, j = col % P;
  const uint4 v = *reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(b) * C + c) * H + row) * W + col);
  const int64_t orow = (static_cast<int64_t>(b) * (H / P) + ph) * (W / P) + pw;
  *reinterpret_cast<uint4*>(out + orow * (static_cast<int64_t>(C) * P * P) + (c * P + i) * P + j) = v;
}
}  // namespace
}  // namespace samq

extern "C" int samq_patchify_fwd(const void* x, void* out, int B, int C, int H, int W, int P, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(x && out, SAMQ_ERR_BAD_ARG, "samq_patchify_fwd: null pointer");
  SAMQ_REQUIRE(B > 0 && C > 0 && P > 0 && P % 8 == 0 && H % P == 0 && W % P == 0, SAMQ_ERR_BAD_SHAPE,
               "samq_patchify_fwd: B=%d C=%d H=%d W=%d P=%d (P %% 8 == 0, H,W multiples of P)", B, C, H, W, P);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, SAMQ_ERR_BAD_ARG,
               "samq_patchify_fwd: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * C * H * (W / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  patchify_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), B, C, H, W, P);
  count_launch();
  return check_launch("patchify_kernel");
}

namespace samq {
namespace {
// 3x3 / stride 1 / zero-padding 1 neighbourhoods of an NHWC tensor as GEMM rows (the neck's second
// convolution, image_encoder.py:96-103, is  rows x [9 C] . W[O, (ky, kx, c)]^T):
// out[(b, h, w), (ky*3 + kx) * C + c] = x[b, h + ky - 1, w + kx - 1, c], zero outside the image.
// One thread per 16-byte chunk; a warp writes 512 contiguous bytes and reads 512 contiguous bytes.
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const __half* __restrict__ x, __half* __restrict__ out, int B, int H, int W, int C) {
  const int cpt = C / 8;                       // 16-byte chunks per tap
  const int64_t total = static_cast<int64_t>(B) * H * W * 9 * cpt;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % cpt);
  int64_t t = idx / cpt;
  const int tap = static_cast<int>(t % 9);
  t /= 9;                                      // output pixel (b, h, w)
  const int w = static_cast<int>(t % W);
  const int h = static_cast<int>((t / W) % H);
  const int b = static_cast<int>(t / (static_cast<int64_t>(W) * H));
  const int hs = h + tap / 3 - 1, ws = w + tap % 3 - 1;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (hs >= 0 && hs < H && ws >= 0 && ws < W)
    v = *reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(b) * H + hs) * W + ws) * C + chunk * 8);
  *reinterpret_cast<uint4*>(out + t * (9 * static_cast<int64_t>(C)) + tap * C + chunk * 8) = v;
}
}  // namespace
}  // namespace samq

extern "C" int samq_im2col3x3_fwd(const void* x, void* out, int B, int H, int W, int C, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(x && out, SAMQ_ERR_BAD_ARG, "samq_im2col3x3_fwd: null pointer");
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, SAMQ_ERR_BAD_SHAPE,
               "samq_im2col3x3_fwd: B=%d H=%d W=%d C=%d (C %% 8 == 0)", B, H, W, C);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, SAMQ_ERR_BAD_ARG,
               "samq_im2col3x3_fwd: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * H * W * 9 * (C / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  im2col3x3_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(out), B, H, W, C);
  count_launch();
  return check_launch("im2col3x3_kernel");
}

extern "C" int samq_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(a && b && out, SAMQ_ERR_BAD_ARG, "samq_add: null pointer");
  SAMQ_REQUIRE(n >= 0 && n % 8 == 0, SAMQ_ERR_BAD_SHAPE, "samq_add: n=%lld must be a multiple of 8", (long long)n);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "samq_add: pointers must be 16-byte aligned");
  if (n == 0) return SAMQ_OK;
  const int64_t nvec = n / 8;
  const unsigned grid = static_cast<unsigned>((nvec + 255) / 256);
  add_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(a), reinterpret_cast<const __half*>(b),
      reinterpret_cast<__half*>(out), nvec);
  count_launch();
  return check_launch("add_kernel");
}
