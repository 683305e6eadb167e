// Prompt-side kernels of SAM's mask decoder: the two-way transformer's
// attention between a handful of prompt tokens and the 64x64 image tokens
// (segment_anything/modeling/transformer.py:185-240) and the skinny linears on the prompt tokens
// (transformer.py:151-153, mask_decoder.py:155-178).  < 1 % of the encoder's FLOPs and a few
// hundred KB of operands: SIMT kernels with fp32 math; the 4096-token linears of the decoder go
// through the tcgen05 dense GEMM (samq_dense_linear_fwd), its LayerNorms through the LN kernel.
#include "common.cuh"

namespace samq {
namespace {

// ---- attention without positional bias: out = softmax(q k^T * scale) v ------------------------
// q [B, Nq, heads*HD], k / v [B, Nk, heads*HD], out [B, Nq, heads*HD], fp16, row-major.
// One warp per (batch, head, query): the lanes split the keys (online softmax per lane, fp32),
// then the 32 partial (max, sum, acc[HD]) triples are merged with shuffles.  Serves both shapes of
// the decoder: few queries x 4096 keys (tokens -> image) and 4096 queries x few keys (image -> tokens).
template <int HD>
__global__ void __launch_bounds__(128)
attn_small_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                  __half* __restrict__ out, int B, int heads, int Nq, int Nk, float scale_log2e) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  const int64_t total = static_cast<int64_t>(B) * heads * Nq;
  if (wid >= total) return;
  const int iq = static_cast<int>(wid % Nq);
  const int head = static_cast<int>((wid / Nq) % heads);
  const int b = static_cast<int>(wid / (static_cast<int64_t>(Nq) * heads));
  const int D = heads * HD;
  float qv[HD];
  {
    const uint4* src = reinterpret_cast<const uint4*>(q + (static_cast<int64_t>(b) * Nq + iq) * D + head * HD);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = src[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        qv[c * 8 + 2 * e] = f.x * scale_log2e;
        qv[c * 8 + 2 * e + 1] = f.y * scale_log2e;
      }
    }
  }
  float m = -INFINITY, l = 0.f, acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const __half* kb = k + static_cast<int64_t>(b) * Nk * D + head * HD;
  const __half* vb = v + static_cast<int64_t>(b) * Nk * D + head * HD;
  for (int t = lane; t < Nk; t += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(t) * D);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = kr[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        s = fmaf(qv[c * 8 + 2 * e], f.x, s);
        s = fmaf(qv[c * 8 + 2 * e + 1], f.y, s);
      }
    }
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    m = mn;
    l = fmaf(l, corr, p);
    const uint4* vr = reinterpret_cast<const uint4*>(vb + static_cast<int64_t>(t) * D);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = vr[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc[c * 8 + 2 * e] = fmaf(acc[c * 8 + 2 * e], corr, p * f.x);
        acc[c * 8 + 2 * e + 1] = fmaf(acc[c * 8 + 2 * e + 1], corr, p * f.y);
      }
    }
  }
  // merge the lanes: common maximum, rescale, sum
  float mw = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  const float r = m == -INFINITY ? 0.f : exp2f(m - mw);
  l *= r;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
  __half* dst = out + (static_cast<int64_t>(b) * Nq + iq) * D + head * HD;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    float a = acc[d] * r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == (d & 31)) dst[d] = __float2half_rn(a * inv);
  }
}

// Few keys (image -> tokens: 4096 queries per image, ~9 keys): one THREAD per (batch, head, query), the
// keys walked sequentially -- a warp per query would keep 9 of 32 lanes busy and launch 2 M warps.
template <int HD>
__global__ void __launch_bounds__(128)
attn_fewkeys_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                    __half* __restrict__ out, int B, int heads, int Nq, int Nk, float scale_log2e) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * heads * Nq;
  if (tid >= total) return;
  // consecutive threads = consecutive heads of one query: their q / out segments are contiguous
  const int head = static_cast<int>(tid % heads);
  const int iq = static_cast<int>((tid / heads) % Nq);
  const int b = static_cast<int>(tid / (static_cast<int64_t>(heads) * Nq));
  const int D = heads * HD;
  float qv[HD], acc[HD];
  {
    const uint4* src = reinterpret_cast<const uint4*>(q + (static_cast<int64_t>(b) * Nq + iq) * D + head * HD);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = src[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        qv[c * 8 + 2 * e] = f.x * scale_log2e;
        qv[c * 8 + 2 * e + 1] = f.y * scale_log2e;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  const __half* kb = k + static_cast<int64_t>(b) * Nk * D + head * HD;
  const __half* vb = v + static_cast<int64_t>(b) * Nk * D + head * HD;
  for (int t = 0; t < Nk; ++t) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(t) * D);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = kr[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        s = fmaf(qv[c * 8 + 2 * e], f.x, s);
        s = fmaf(qv[c * 8 + 2 * e + 1], f.y, s);
      }
    }
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    m = mn;
    l = fmaf(l, corr, p);
    const uint4* vr = reinterpret_cast<const uint4*>(vb + static_cast<int64_t>(t) * D);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = vr[c];
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc[c * 8 + 2 * e] = fmaf(acc[c * 8 + 2 * e], corr, p * f.x);
        acc[c * 8 + 2 * e + 1] = fmaf(acc[c * 8 + 2 * e + 1], corr, p * f.y);
      }
    }
  }
  const float inv = 1.f / l;
  uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * Nq + iq) * D + head * HD);
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    __align__(16) __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(acc[c * 8 + 2 * e] * inv, acc[c * 8 + 2 * e + 1] * inv);
    dst[c] = *reinterpret_cast<const uint4*>(h);
  }
}

// Few reduction terms (K <= 64: the mask product hyper_in . upscaled, K = 32): one THREAD per output.
__global__ void __launch_bounds__(256)
small_linear_shortk_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                           const __half* __restrict__ residual, __half* __restrict__ y, int64_t M, int N, int K, int act) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (tid >= M * N) return;
  const int64_t m = tid / N;
  const int n = static_cast<int>(tid % N);
  const uint4* xr = reinterpret_cast<const uint4*>(x + m * K);
  const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n) * K);
  float s = 0.f;
  for (int c = 0; c < K / 8; ++c) {
    const uint4 a = xr[c], b = wr[c];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
      s = fmaf(fa.x, fb.x, s);
      s = fmaf(fa.y, fb.y, s);
    }
  }
  if (bias) s += __half2float(bias[n]);
  if (act == 1) s = 0.5f * s * (1.f + erff(s * 0.70710678118654752440f));
  else if (act == 2) s = fmaxf(s, 0.f);
  __half h = __float2half_rn(s);
  if (residual) h = __hadd(h, residual[m * N + n]);
  y[m * N + n] = h;
}

// ---- skinny linear: y[m, n] = act(x[m, :] . w[n, :] + bias[n]) (+ residual[m, n]) --------------
// x fp16 [M, K], w fp16 [N, K] (nn.Linear layout), fp32 accumulation, any N, K % 8 == 0.
// One warp per output element pair is wasteful for long M; this kernel is for the decoder's
// prompt tokens (M = batch x ~8) and hypernetwork heads (N = 32, 4), where the 128-feature
// tensor-core tile would be mostly padding.  act: 0 none, 1 exact-erf GELU, 2 ReLU.
__global__ void __launch_bounds__(256)
small_linear_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                    const __half* __restrict__ residual, __half* __restrict__ y, int64_t M, int N, int K, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (wid >= M * N) return;
  const int64_t m = wid / N;
  const int n = static_cast<int>(wid % N);
  const uint4* xr = reinterpret_cast<const uint4*>(x + m * K);
  const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n) * K);
  float s = 0.f;
  for (int c = lane; c < K / 8; c += 32) {
    const uint4 a = xr[c], b = wr[c];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
      s = fmaf(fa.x, fb.x, s);
      s = fmaf(fa.y, fb.y, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (bias) s += __half2float(bias[n]);
    if (act == 1) s = 0.5f * s * (1.f + erff(s * 0.70710678118654752440f));
    else if (act == 2) s = fmaxf(s, 0.f);
    __half h = __float2half_rn(s);
    if (residual) h = __hadd(h, residual[m * N + n]);   // fp16 + fp16 like the reference's `queries + mlp_out`
    y[m * N + n] = h;
  }
}

// exact-erf GELU, elementwise on fp16 (the activation after the LayerNorm2d of the mask decoder's
// output_upscaling, mask_decoder.py:55-59; the other GELUs ride in GEMM epilogues)
__global__ void __launch_bounds__(256) gelu_kernel(const __half2* __restrict__ x, __half2* __restrict__ y, int64_t n2) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n2) return;
  const float2 f = __half22float2(x[i]);
  y[i] = __floats2half2_rn(0.5f * f.x * (1.f + erff(f.x * 0.70710678118654752440f)),
                           0.5f * f.y * (1.f + erff(f.y * 0.70710678118654752440f)));
}

}  // namespace
}  // namespace samq

extern "C" int samq_gelu_fwd(const void* x, void* y, int64_t n, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(x && y, SAMQ_ERR_BAD_ARG, "samq_gelu_fwd: null pointer");
  SAMQ_REQUIRE(n > 0 && n % 2 == 0, SAMQ_ERR_BAD_SHAPE, "samq_gelu_fwd: n=%lld must be positive and even", (long long)n);
  gelu_kernel<<<static_cast<unsigned>((n / 2 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half2*>(x), reinterpret_cast<__half2*>(y), n / 2);
  count_launch();
  return check_launch("gelu_kernel");
}

extern "C" int samq_attn_small_fwd(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq,
                                   int Nk, int hd, float scale, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(q && k && v && out, SAMQ_ERR_BAD_ARG, "samq_attn_small_fwd: null pointer");
  SAMQ_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0, SAMQ_ERR_BAD_SHAPE, "samq_attn_small_fwd: B=%d heads=%d Nq=%d Nk=%d",
               B, heads, Nq, Nk);
  SAMQ_REQUIRE(hd == 16 || hd == 32 || hd == 64, SAMQ_ERR_BAD_SHAPE, "samq_attn_small_fwd: head dim %d (16, 32 or 64)", hd);
  for (const void* p : {q, k, v, static_cast<const void*>(out)})
    SAMQ_REQUIRE(reinterpret_cast<uintptr_t>(p) % 16 == 0, SAMQ_ERR_BAD_ARG, "samq_attn_small_fwd: pointers must be 16-byte aligned");
  const int64_t warps = static_cast<int64_t>(B) * heads * Nq;
  const unsigned grid = static_cast<unsigned>((warps + 3) / 4);
  const float sl = scale * 1.4426950408889634f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half *qh = reinterpret_cast<const __half*>(q), *kh = reinterpret_cast<const __half*>(k),
               *vh = reinterpret_cast<const __half*>(v);
  __half* oh = reinterpret_cast<__half*>(out);
  if (Nk <= 32 && warps >= 4096 && hd <= 32) {     // many queries, few keys: a thread per query
    const unsigned g2 = static_cast<unsigned>((warps + 127) / 128);
    if (hd == 16) attn_fewkeys_kernel<16><<<g2, 128, 0, st>>>(qh, kh, vh, oh, B, heads, Nq, Nk, sl);
    else attn_fewkeys_kernel<32><<<g2, 128, 0, st>>>(qh, kh, vh, oh, B, heads, Nq, Nk, sl);
    count_launch();
    return check_launch("attn_fewkeys_kernel");
  }
  if (hd == 16) attn_small_kernel<16><<<grid, 128, 0, st>>>(qh, kh, vh, oh, B, heads, Nq, Nk, sl);
  else if (hd == 32) attn_small_kernel<32><<<grid, 128, 0, st>>>(qh, kh, vh, oh, B, heads, Nq, Nk, sl);
  else attn_small_kernel<64><<<grid, 128, 0, st>>>(qh, kh, vh, oh, B, heads, Nq, Nk, sl);
  count_launch();
  return check_launch("attn_small_kernel");
}

extern "C" int samq_small_linear_fwd(const void* x, const void* w, const void* bias, const void* residual, void* y,
                                     int64_t M, int N, int K, int act, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(x && w && y, SAMQ_ERR_BAD_ARG, "samq_small_linear_fwd: null pointer");
  SAMQ_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && M * N < (1ll << 34), SAMQ_ERR_BAD_SHAPE,
               "samq_small_linear_fwd: M=%lld N=%d K=%d (K must be a multiple of 8)", (long long)M, N, K);
  SAMQ_REQUIRE(act >= 0 && act <= 2, SAMQ_ERR_BAD_ARG, "samq_small_linear_fwd: act=%d", act);
  SAMQ_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0, SAMQ_ERR_BAD_ARG,
               "samq_small_linear_fwd: x and w must be 16-byte aligned");
  if (K <= 64) {
    small_linear_shortk_kernel<<<static_cast<unsigned>((M * N + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(w), reinterpret_cast<const __half*>(bias),
        reinterpret_cast<const __half*>(residual), reinterpret_cast<__half*>(y), M, N, K, act);
    count_launch();
    return check_launch("small_linear_shortk_kernel");
  }
  const int64_t warps = M * N;
  small_linear_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(w), reinterpret_cast<const __half*>(bias),
      reinterpret_cast<const __half*>(residual), reinterpret_cast<__half*>(y), M, N, K, act);
  count_launch();
  return check_launch("small_linear_kernel");
}
