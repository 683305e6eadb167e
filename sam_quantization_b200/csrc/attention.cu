// Flash-style attention with SAM's decomposed relative-position bias on sm_100a.
//
// Replaces QuantAttention.forward's q-slice copy, add_decomposed_rel_pos (2 x
// get_rel_pos + 2 batched matmuls through HBM) and the Triton kernel _fwd_kernel1
// (gptq_triton/fused_attention.py:46-80, 107-133, 159-358).
//
//   out[b, m, head, :] = softmax_n( scale * q[m].k[n] + rel_h[m, n / E] + rel_w[m, n % E] ) v[n]
//   rel_h[m, kh] = fp16( q[m] . rel_pos_h[h(m) - kh + E-1] )
//   rel_w[m, kw] = fp16( q[m] . rel_pos_w[r(m) - kw + E-1] ),  r(m) = h(m) in "reference"
//                  mode (the fork's matmul broadcasting, fused_attention.py:76-78)
//                  or w(m) in "upstream" mode.
//
// One CTA = one (batch/window, head, 128-query tile).  tcgen05 everywhere:
//   prologue: T_h = Q . rel_pos_h^T, T_w = Q . rel_pos_w^T  (two MMAs into the S
//             buffers), rounded to fp16 and bounced through shared memory so each
//             softmax thread (= one query row) can gather its E + E bias values;
//   loop    : S_j = Q . K_j^T (SS MMA, fp32 in TMEM, double buffered)
//             softmax threads read S_j from TMEM (one row per thread: no shuffles),
//             add bias, online softmax in base 2 (fused_attention.py:219,275-293)
//             with lazy rescaling, write P_j as fp16 back into the S_j columns
//             O += P_j . V_j (TS MMA: P from TMEM, V tile MN-major in shared memory)
//   epilogue: O / l -> fp16 -> [B, S, heads*hd].
// head_dim 80 is handled without padding to 128 (the reference pads,
// fused_attention.py:323): K = 64 (128B-swizzle tile) + 16 (32B-swizzle tile).
//
// Translation units: attention_win.cu (persistent 14x14 kernel), attention_glob.cu (64x64 kernel),
// this file (C ABI + dispatch); attention_ablations.cu holds the earlier designs and is compiled
// only with `make ABLATIONS=1`.
#include "attention_common.cuh"

extern "C" int samq_attn_relpos_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                                    void* out, int B, int H, int W, int heads, int hd, float scale,
                                    int relw_mode, void* stream) {
  using namespace samq;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SAMQ_REQUIRE(qkv && rel_pos_h && rel_pos_w && out, SAMQ_ERR_BAD_ARG, "samq_attn_relpos_fwd: null pointer");
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(rel_pos_h) |
                reinterpret_cast<uintptr_t>(rel_pos_w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "samq_attn_relpos_fwd: pointers must be 16-byte aligned");
  SAMQ_REQUIRE(relw_mode == SAMQ_RELW_REFERENCE || relw_mode == SAMQ_RELW_UPSTREAM, SAMQ_ERR_BAD_ARG,
               "samq_attn_relpos_fwd: bad relw_mode %d", relw_mode);
  SAMQ_REQUIRE(B > 0 && B <= 65535 && heads > 0 && heads <= 65535, SAMQ_ERR_BAD_SHAPE,
               "samq_attn_relpos_fwd: B=%d heads=%d out of range", B, heads);
  SAMQ_REQUIRE(hd == 64 || hd == 80, SAMQ_ERR_BAD_SHAPE,
               "samq_attn_relpos_fwd: head_dim %d not supported (64 or 80)", hd);
  const bool glob = (H == 64 && W == 64), win = (H == 14 && W == 14);
  SAMQ_REQUIRE(glob || win, SAMQ_ERR_BAD_SHAPE,
               "samq_attn_relpos_fwd: (H,W)=(%d,%d) not supported ((64,64) or (14,14))", H, W);
#ifdef SAMQ_ABLATIONS
  const int gen = glob ? config().attn_glob : config().attn_win;   // SAMQ_ATTN_GLOB / SAMQ_ATTN_WIN = v1 | v2
  if (gen == 1 || gen == 2)
    return attn_ablation_dispatch(gen, glob, hd, qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
#endif
  if (glob) return attn_glob3_dispatch(hd, qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
  return attn_win3_dispatch(hd, qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, 0, 0, st);
}

extern "C" int samq_attn_relpos_unpartition_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                                                void* out, int B, int H, int W, int ws, int heads, int hd,
                                                float scale, int relw_mode, void* stream) {
  using namespace samq;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* who = "samq_attn_relpos_unpartition_fwd";
  SAMQ_REQUIRE(qkv && rel_pos_h && rel_pos_w && out, SAMQ_ERR_BAD_ARG, "%s: null pointer", who);
  SAMQ_REQUIRE((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(rel_pos_h) |
                reinterpret_cast<uintptr_t>(rel_pos_w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "%s: pointers must be 16-byte aligned", who);
  SAMQ_REQUIRE(relw_mode == SAMQ_RELW_REFERENCE || relw_mode == SAMQ_RELW_UPSTREAM, SAMQ_ERR_BAD_ARG,
               "%s: bad relw_mode %d", who, relw_mode);
  SAMQ_REQUIRE(ws == 14, SAMQ_ERR_BAD_SHAPE, "%s: window size %d not supported (14)", who, ws);
  SAMQ_REQUIRE(hd == 64 || hd == 80, SAMQ_ERR_BAD_SHAPE, "%s: head_dim %d not supported (64 or 80)", who, hd);
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && heads <= 65535, SAMQ_ERR_BAD_SHAPE,
               "%s: B=%d H=%d W=%d heads=%d out of range", who, B, H, W, heads);
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  const int64_t windows = static_cast<int64_t>(B) * nH * nW;
  SAMQ_REQUIRE(windows * heads < (1ll << 31), SAMQ_ERR_BAD_SHAPE, "%s: too many (window, head) items", who);
  return attn_win3_dispatch(hd, qkv, rel_pos_h, rel_pos_w, out, static_cast<int>(windows), heads, scale, relw_mode,
                            H, W, st);
}
