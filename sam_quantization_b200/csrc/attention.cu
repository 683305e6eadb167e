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
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace samq {
namespace {

constexpr int kAttThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

template <int HD, bool WIN>
struct ACfg {
  static constexpr int E = WIN ? 14 : 64;           // H == W
  static constexpr int S = E * E;                   // tokens per image / window
  static constexpr int kQTiles = (S + 127) / 128;
  static constexpr int kKVTiles = (S + 127) / 128;
  static constexpr int kTail = HD - 64;             // 0 or 16
  static constexpr int kMainBytes = 128 * 128;      // 128 rows x 64 fp16, 128B swizzle
  static constexpr int kTailBytes = kTail ? 128 * 32 : 0;  // 128 rows x 16 fp16, 32B swizzle
  static constexpr int kTileBytes = kMainBytes + kTailBytes;
  static constexpr int kRpRows = WIN ? 32 : 128;    // rel-pos table rows (2E-1) padded
  static constexpr int kRpMainBytes = kRpRows * 128;
  static constexpr int kRpTailBytes = kTail ? kRpRows * 32 : 0;
  static constexpr int kRpBytes = kRpMainBytes + kRpTailBytes;
  static constexpr int kStages = (HD == 64) ? 3 : 2;
  static constexpr int kBounceWords = WIN ? 17 : 65;  // row stride (32-bit words), odd: conflict-free
  static constexpr int kBounceBytes = ((128 * kBounceWords * 4 + 1023) / 1024) * 1024;
  // shared memory carve (all tile bases 1024-aligned)
  static constexpr int oQ = 0;
  static constexpr int oRph = oQ + kTileBytes;
  static constexpr int oRpw = oRph + ((kRpBytes + 1023) / 1024) * 1024;
  static constexpr int oKV = oRpw + ((kRpBytes + 1023) / 1024) * 1024;
  static constexpr int oTh = oKV + kStages * 2 * kTileBytes;
  static constexpr int oTw = oTh + kBounceBytes;
  static constexpr int oBars = oTw + kBounceBytes;
  static constexpr int kNumBars = 1 + 2 * kStages + 2 + 2 + 2 + 2;
  static constexpr int kSmemBytes = oBars + kNumBars * 8 + 16 + 1024;
  // TMEM columns
  static constexpr int cS0 = 0, cS1 = 128, cO = 256;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2v(float x) {   // volatile: keeps its place among other volatile asm
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int HD, bool WIN>
__global__ void __launch_bounds__(kAttThreads, 1)
attn_relpos_kernel(const __grid_constant__ CUtensorMap map_qkv_main,
                   const __grid_constant__ CUtensorMap map_qkv_tail,
                   const __grid_constant__ CUtensorMap map_rph_main,
                   const __grid_constant__ CUtensorMap map_rph_tail,
                   const __grid_constant__ CUtensorMap map_rpw_main,
                   const __grid_constant__ CUtensorMap map_rpw_tail, __half* __restrict__ out,
                   int heads, float scale, int relw_mode) {
  using C = ACfg<HD, WIN>;
  constexpr int E = C::E, S = C::S, T = C::kKVTiles;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + C::oQ;
  uint8_t* sRph = smem + C::oRph;
  uint8_t* sRpw = smem + C::oRpw;
  uint8_t* sKV = smem + C::oKV;
  uint32_t* sTh = reinterpret_cast<uint32_t*>(smem + C::oTh);
  uint32_t* sTw = reinterpret_cast<uint32_t*>(smem + C::oTw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBars);
  uint64_t* q_full = bars;
  uint64_t* kv_full = q_full + 1;
  uint64_t* kv_empty = kv_full + C::kStages;
  uint64_t* t_full = kv_empty + C::kStages;
  uint64_t* t_done = t_full + 1;
  uint64_t* s_full = t_done + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int D = heads * HD;
  const int m0 = q_tile * 128;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_done, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_qkv_main);
    tma_prefetch_desc(&map_rph_main);
    tma_prefetch_desc(&map_rpw_main);
    if (C::kTail) {
      tma_prefetch_desc(&map_qkv_tail);
      tma_prefetch_desc(&map_rph_tail);
      tma_prefetch_desc(&map_rpw_tail);
    }
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::kTileBytes + 2 * C::kRpBytes);
      tma_load_3d(sQ, &map_qkv_main, q_full, head * HD, m0, b);
      tma_load_2d(sRph, &map_rph_main, q_full, 0, 0);
      tma_load_2d(sRpw, &map_rpw_main, q_full, 0, 0);
      if (C::kTail) {
        tma_load_3d(sQ + C::kMainBytes, &map_qkv_tail, q_full, head * HD + 64, m0, b);
        tma_load_2d(sRph + C::kRpMainBytes, &map_rph_tail, q_full, 64, 0);
        tma_load_2d(sRpw + C::kRpMainBytes, &map_rpw_tail, q_full, 64, 0);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * C::kTileBytes);
        uint8_t* sK = sKV + s * 2 * C::kTileBytes;
        uint8_t* sV = sK + C::kTileBytes;
        tma_load_3d(sK, &map_qkv_main, &kv_full[s], D + head * HD, j * 128, b);
        tma_load_3d(sV, &map_qkv_main, &kv_full[s], 2 * D + head * HD, j * 128, b);
        if (C::kTail) {
          tma_load_3d(sK + C::kMainBytes, &map_qkv_tail, &kv_full[s], D + head * HD + 64, j * 128, b);
          tma_load_3d(sV + C::kMainBytes, &map_qkv_tail, &kv_full[s], 2 * D + head * HD + 64, j * 128, b);
        }
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // whole warp runs the loop convergently (uniform registers), one elected lane issues
    {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0);
      constexpr uint32_t idesc_t = make_idesc_f16(128, C::kRpRows, 0);
      constexpr uint32_t idesc_pv_main = make_idesc_f16(128, 64, 1);
      constexpr uint32_t idesc_pv_tail = make_idesc_f16(128, 16, 1);
      const uint64_t q_main = make_smem_desc(smem_u32(sQ), 0, 1024, kLayoutSw128);
      const uint64_t q_tail = make_smem_desc(smem_u32(sQ + C::kMainBytes), 0, 256, kLayoutSw32);

      // S_buf = Q . B^T for a K-major B tile (K tile or rel-pos table)
      auto mma_q_times = [&](uint32_t d_tmem, const uint8_t* b_main_ptr, const uint8_t* b_tail_ptr,
                             uint32_t idesc, uint64_t* done_bar) {
        const uint64_t b_main = make_smem_desc(smem_u32(b_main_ptr), 0, 1024, kLayoutSw128);
        const uint64_t b_tail = make_smem_desc(smem_u32(b_tail_ptr), 0, 256, kLayoutSw32);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_ss(d_tmem, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc, k > 0);
          if (C::kTail) tc_mma_ss(d_tmem, q_tail, b_tail, idesc, 1);
          if (done_bar) tc_commit(done_bar);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      tc_fence_after();
      mma_q_times(tmem_base + C::cS0, sRph, sRph + C::kRpMainBytes, idesc_t, nullptr);
      mma_q_times(tmem_base + C::cS1, sRpw, sRpw + C::kRpMainBytes, idesc_t, t_full);

      mbar_wait(&kv_full[0], 0);
      mbar_wait(t_done, 0);  // softmax threads have copied T_h / T_w out of the S buffers
      tc_fence_after();
      mma_q_times(tmem_base + C::cS0, sKV, sKV + C::kMainBytes, idesc_qk, &s_full[0]);

      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          int s1 = s + 1;
          uint32_t ph1 = ph;
          if (s1 == C::kStages) { s1 = 0; ph1 ^= 1; }
          mbar_wait(&kv_full[s1], ph1);
          tc_fence_after();
          const uint8_t* sK1 = sKV + s1 * 2 * C::kTileBytes;
          // in-order tensor pipe: this overwrite of S[(j+1)&1] is ordered after PV(j-1)
          mma_q_times(tmem_base + (((j + 1) & 1) ? C::cS1 : C::cS0), sK1, sK1 + C::kMainBytes, idesc_qk,
                      &s_full[(j + 1) & 1]);
        }
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint8_t* sV = sKV + s * 2 * C::kTileBytes + C::kTileBytes;
        const uint32_t p_tmem = tmem_base + ((j & 1) ? C::cS1 : C::cS0);
        const int keys = (S - j * 128) < 128 ? (S - j * 128) : 128;
        const int ksteps = (keys + 15) / 16;
        const uint64_t v_main0 = make_smem_desc(smem_u32(sV), C::kMainBytes, 1024, kLayoutSw128);
        const uint64_t v_tail0 = make_smem_desc(smem_u32(sV + C::kMainBytes), C::kTailBytes, 256, kLayoutSw32);
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
            tc_mma_ts(tmem_base + C::cO, p_tmem + ks * 8, v_main0 + (ks * 2048 >> 4), idesc_pv_main, acc);
            if (C::kTail)
              tc_mma_ts(tmem_base + C::cO + 64, p_tmem + ks * 8, v_tail0 + (ks * 512 >> 4), idesc_pv_tail, acc);
          }
          tc_commit(&kv_empty[s]);
          tc_commit(&pv_done[j & 1]);
        }
        __syncwarp();
        if (++s == C::kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ============================ softmax / epilogue ============================
    const int e = warp - 4;
    const int row = e * 32 + lane;            // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(e * 32) << 16;
    const int m = m0 + row;
    const bool valid = m < S;
    const int mh = valid ? m / E : 0, mw = valid ? m % E : 0;
    const float c_scale = scale * kLog2e;

    // ---- rel-pos tables: TMEM -> fp16 -> shared (own row only) ----
    mbar_wait(t_full, 0);
    tc_fence_after();
    uint32_t* my_th = sTh + row * C::kBounceWords;
    uint32_t* my_tw = sTw + row * C::kBounceWords;
#pragma unroll
    for (int c = 0; c < C::kRpRows / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_base + C::cS0 + c * 32 + lane_off, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        my_th[c * 16 + i] = pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      tmem_ld_x32(tmem_base + C::cS1 + c * 32 + lane_off, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        my_tw[c * 16 + i] = pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_done);

    const __half* th_row = reinterpret_cast<const __half*>(my_th);
    const __half* tw_row = reinterpret_cast<const __half*>(my_tw);
    const int rw = (relw_mode == SAMQ_RELW_UPSTREAM) ? mw : mh;
    float bw[E];   // log2e * rel_w[m, kw]
#pragma unroll
    for (int kw = 0; kw < E; ++kw) bw[kw] = kLog2e * __half2float(tw_row[rw - kw + E - 1]);
    float bh_win[WIN ? E : 1];  // windowed: log2e * rel_h[m, kh] for all kh
    if (WIN) {
#pragma unroll
      for (int kh = 0; kh < E; ++kh) bh_win[kh] = kLog2e * __half2float(th_row[mh - kh + E - 1]);
    }

    float m_used = -INFINITY, l = 0.f;
    // windowed: 2 tiles, fully unrolled so key -> (kh, kw) is resolved at compile time
#pragma unroll(WIN ? 2 : 1)
    for (int j = 0; j < T; ++j) {
      const uint32_t s_tmem = tmem_base + ((j & 1) ? C::cS1 : C::cS0) + lane_off;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float bh0 = 0.f, bh1 = 0.f;
      if (!WIN) {
        bh0 = kLog2e * __half2float(th_row[mh - 2 * j + E - 1]);
        bh1 = kLog2e * __half2float(th_row[mh - 2 * j - 1 + E - 1]);
      }
      // ---- pass 1: tile maximum ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(s_tmem + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = j * 128 + c * 32 + i;  // key index (compile-time in WIN mode)
          if (WIN) {
            if (n < S) mx0 = fmaxf(mx0, fmaf(__uint_as_float(r[i]), c_scale, bh_win[(n / E) % E] + bw[n % E]));
          } else {
            const float x = fmaf(__uint_as_float(r[i]), c_scale, bw[(c * 32 + i) % E]);
            if (c < 2) mx0 = fmaxf(mx0, x); else mx1 = fmaxf(mx1, x);
          }
        }
      }
      const float m_tile = WIN ? mx0 : fmaxf(mx0 + bh0, mx1 + bh1);
      const float m_new = fmaxf(m_used, m_tile);
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, m_new > m_used + 8.f)) {
        // lazy rescale of the running output (rare once the maximum has settled)
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const float alpha = ex2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
        const uint32_t o_tmem = tmem_base + C::cO + lane_off;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_x32(o_tmem + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x32(o_tmem + c * 32, r);
        }
        if (C::kTail) {
          uint32_t r[16];
          tmem_ld_x16(o_tmem + 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x16(o_tmem + 64, r);
        }
        tmem_st_wait();
      }
      // ---- pass 2: P = 2^(x - m) as fp16 into the S columns, row sum ----
      const float mm0 = m_used - bh0, mm1 = m_used - bh1;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(s_tmem + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int n = j * 128 + c * 32 + i + u;
            if (WIN) {
              p[u] = (n < S) ? ex2(fmaf(__uint_as_float(r[i + u]), c_scale, bh_win[(n / E) % E] + bw[n % E]) - m_used)
                             : 0.f;
            } else {
              const float x = fmaf(__uint_as_float(r[i + u]), c_scale, bw[(c * 32 + i + u) % E]);
              p[u] = ex2(x - (c < 2 ? mm0 : mm1));
            }
            sum += p[u];
          }
          pk[i >> 1] = pack_h2(p[0], p[1]);
        }
        tmem_st_x16(s_tmem + c * 16, pk);
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }

    // ---- epilogue: O / l ----
    mbar_wait(&pv_done[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const uint32_t o_tmem = tmem_base + C::cO + lane_off;
    __half* dst = out + (static_cast<size_t>(b) * S + (valid ? m : 0)) * D + head * HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_x32(o_tmem + c * 32, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          o.x = pack_h2(__uint_as_float(r[8 * v + 0]) * inv_l, __uint_as_float(r[8 * v + 1]) * inv_l);
          o.y = pack_h2(__uint_as_float(r[8 * v + 2]) * inv_l, __uint_as_float(r[8 * v + 3]) * inv_l);
          o.z = pack_h2(__uint_as_float(r[8 * v + 4]) * inv_l, __uint_as_float(r[8 * v + 5]) * inv_l);
          o.w = pack_h2(__uint_as_float(r[8 * v + 6]) * inv_l, __uint_as_float(r[8 * v + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = o;
        }
      }
    }
    if (C::kTail) {
      uint32_t r[16];
      tmem_ld_x16(o_tmem + 64, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 o;
          o.x = pack_h2(__uint_as_float(r[8 * v + 0]) * inv_l, __uint_as_float(r[8 * v + 1]) * inv_l);
          o.y = pack_h2(__uint_as_float(r[8 * v + 2]) * inv_l, __uint_as_float(r[8 * v + 3]) * inv_l);
          o.z = pack_h2(__uint_as_float(r[8 * v + 4]) * inv_l, __uint_as_float(r[8 * v + 5]) * inv_l);
          o.w = pack_h2(__uint_as_float(r[8 * v + 6]) * inv_l, __uint_as_float(r[8 * v + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + 64 + v * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#ifdef SAMQ_ATTN_PROFILE
// developer-only wait-time breakdown (tests/micro/attn_prof.cu); never compiled into libsamq.so
__device__ long long g_attn_prof[12][8];
#define PROF_DECL long long pt0 = 0, pstart = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_BEGIN pt0 = clock64()
#ifdef SAMQ_ATTN_STAMPS
#define PROF_END(i)
#else
#define PROF_END(i) pacc[i] += clock64() - pt0
#endif
#define PROF_STAMP(i) pacc[i] = clock64() - pstart
#define PROF_ADD(i, d) pacc[i] += (d)
#define PROF_FLUSH                                                        \
  if (lane == 0 && blockIdx.x == (gridDim.x > 3 ? 3 : 0) && blockIdx.y == (gridDim.y > 1 ? 1 : 0) && blockIdx.z == 0) \
    for (int i_ = 0; i_ < 8; ++i_) g_attn_prof[warp][i_] = pacc[i_]
#else
#define PROF_DECL
#define PROF_BEGIN
#define PROF_END(i)
#define PROF_STAMP(i)
#define PROF_ADD(i, d)
#define PROF_FLUSH
#endif
// ===========================================================================================
// Windowed attention, second design: the whole 14x14 window (196 keys, padded to 208) is ONE
// key tile, so the softmax is exact single-pass (no online rescaling), and the CTA is small
// enough -- 95 KB of shared memory, 256 TMEM columns, <= 128 registers -- that TWO CTAs share
// an SM: one CTA's softmax overlaps the other's TMA / MMA / prologue.  (The first design ran
// one 130 KB / 512-column CTA per SM and was prologue-bound: 151 TFLOP/s.)
//   TMEM columns: S [0,208) fp32  ->  P [0,104) fp16 pairs (aliases S, written chunk by chunk
//   behind the read pointer);  O [128, 128+hd) is written by the PV MMAs only after every S
//   column has been consumed;  rel-pos tables T_h [0,32), T_w [32,64) live there before S.
//   Shared memory: Q | rel_pos_h | rel_pos_w | K (208 rows) | V (208 rows); the fp16 bounce
//   buffers of the rel-pos tables alias the V region (V's TMA is issued after they are read).
// ===========================================================================================
template <int HD>
struct WCfg {
  static constexpr int E = 14, S = 196, SP = 208;    // keys padded to a multiple of 16
  static constexpr int kTail = HD - 64;
  static constexpr int kQMain = 128 * 128, kQTail = kTail ? 128 * 32 : 0, kQBytes = kQMain + kQTail;
  static constexpr int kKMain = SP * 128, kKTail = kTail ? SP * 32 : 0;
  static constexpr int kKMainPad = ((kKMain + 1023) / 1024) * 1024;          // 26624 -> 26624
  static constexpr int kKBytes = kKMainPad + ((kKTail + 1023) / 1024) * 1024;
  static constexpr int kRpMain = 32 * 128, kRpTail = kTail ? 32 * 32 : 0;
  static constexpr int kRpBytes = ((kRpMain + kRpTail + 1023) / 1024) * 1024;
  static constexpr int oQ = 0;
  static constexpr int oRph = oQ + ((kQBytes + 1023) / 1024) * 1024;
  static constexpr int oRpw = oRph + kRpBytes;
  static constexpr int oK = oRpw + kRpBytes;
  static constexpr int oV = oK + kKBytes;
  static constexpr int oBars = oV + kKBytes;
  static constexpr int kBounceWords = 17;                                   // per row, odd stride
  static constexpr int kSmemBytes = oBars + 16 * 8 + 16 + 1024;
  static_assert(2 * 128 * kBounceWords * 4 <= kKBytes, "bounce buffers must fit in the V region");
  static constexpr int cS = 0, cTh = 0, cTw = 32, cO = 128;
};

template <int HD>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_win_kernel(const __grid_constant__ CUtensorMap map_q_main, const __grid_constant__ CUtensorMap map_q_tail,
                const __grid_constant__ CUtensorMap map_kv_main, const __grid_constant__ CUtensorMap map_kv_tail,
                const __grid_constant__ CUtensorMap map_rph_main, const __grid_constant__ CUtensorMap map_rph_tail,
                const __grid_constant__ CUtensorMap map_rpw_main, const __grid_constant__ CUtensorMap map_rpw_tail,
                __half* __restrict__ out, int heads, float scale, int relw_mode) {
  using C = WCfg<HD>;
  constexpr int E = C::E, S = C::S, SP = C::SP;
  PROF_DECL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + C::oQ;
  uint8_t* sRph = smem + C::oRph;
  uint8_t* sRpw = smem + C::oRpw;
  uint8_t* sK = smem + C::oK;
  uint8_t* sV = smem + C::oV;
  uint32_t* sTh = reinterpret_cast<uint32_t*>(sV);                       // aliases V (see above)
  uint32_t* sTw = sTh + 128 * C::kBounceWords;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBars);
  uint64_t* q_full = bars + 0;      // Q + rel-pos tables landed
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* t_full = bars + 3;      // T_h / T_w MMAs done
  uint64_t* t_done = bars + 4;      // softmax warps copied T out of TMEM (count 4)
  uint64_t* b_done = bars + 5;      // softmax warps read their bias values from the bounce (count 4)
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;      // count 4
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int D = heads * HD;
  const int m0 = q_tile * 128;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(t_full, 1);
    mbar_init(t_done, 4); mbar_init(b_done, 4); mbar_init(s_full, 1); mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::kQBytes + 2 * (C::kRpMain + C::kRpTail));
      tma_load_3d(sQ, &map_q_main, q_full, head * HD, m0, b);
      tma_load_2d(sRph, &map_rph_main, q_full, 0, 0);
      tma_load_2d(sRpw, &map_rpw_main, q_full, 0, 0);
      if (C::kTail) {
        tma_load_3d(sQ + C::kQMain, &map_q_tail, q_full, head * HD + 64, m0, b);
        tma_load_2d(sRph + C::kRpMain, &map_rph_tail, q_full, 64, 0);
        tma_load_2d(sRpw + C::kRpMain, &map_rpw_tail, q_full, 64, 0);
      }
      mbar_arrive_expect_tx(k_full, C::kKMain + C::kKTail);
      tma_load_3d(sK, &map_kv_main, k_full, D + head * HD, 0, b);
      if (C::kTail) tma_load_3d(sK + C::kKMainPad, &map_kv_tail, k_full, D + head * HD + 64, 0, b);
      mbar_wait(b_done, 0);                      // bounce buffers (aliasing V) are no longer needed
      mbar_arrive_expect_tx(v_full, C::kKMain + C::kKTail);
      tma_load_3d(sV, &map_kv_main, v_full, 2 * D + head * HD, 0, b);
      if (C::kTail) tma_load_3d(sV + C::kKMainPad, &map_kv_tail, v_full, 2 * D + head * HD + 64, 0, b);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_t = make_idesc_f16(128, 32, 0);
    constexpr uint32_t idesc_qk = make_idesc_f16(128, SP, 0);
    constexpr uint32_t idesc_pv_main = make_idesc_f16(128, 64, 1);
    constexpr uint32_t idesc_pv_tail = make_idesc_f16(128, 16, 1);
    const uint64_t q_main = make_smem_desc(smem_u32(sQ), 0, 1024, kLayoutSw128);
    const uint64_t q_tail = make_smem_desc(smem_u32(sQ + C::kQMain), 0, 256, kLayoutSw32);
    auto mma_q_times = [&](uint32_t d_tmem, const uint8_t* b_main_ptr, const uint8_t* b_tail_ptr, uint32_t idesc,
                           uint64_t* done_bar) {
      const uint64_t b_main = make_smem_desc(smem_u32(b_main_ptr), 0, 1024, kLayoutSw128);
      const uint64_t b_tail = make_smem_desc(smem_u32(b_tail_ptr), 0, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(d_tmem, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc, k > 0);
        if (C::kTail) tc_mma_ss(d_tmem, q_tail, b_tail, idesc, 1);
        if (done_bar) tc_commit(done_bar);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    mma_q_times(tmem_base + C::cTh, sRph, sRph + C::kRpMain, idesc_t, nullptr);
    mma_q_times(tmem_base + C::cTw, sRpw, sRpw + C::kRpMain, idesc_t, t_full);
    mbar_wait(k_full, 0);
    mbar_wait(t_done, 0);
    tc_fence_after();
    mma_q_times(tmem_base + C::cS, sK, sK + C::kKMainPad, idesc_qk, s_full);
    mbar_wait(v_full, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
    const uint64_t v_main0 = make_smem_desc(smem_u32(sV), C::kKMainPad, 1024, kLayoutSw128);
    const uint64_t v_tail0 = make_smem_desc(smem_u32(sV + C::kKMainPad), 4096, 256, kLayoutSw32);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < SP / 16; ++ks) {
        tc_mma_ts(tmem_base + C::cO, tmem_base + ks * 8, v_main0 + (ks * 2048 >> 4), idesc_pv_main, ks > 0);
        if (C::kTail)
          tc_mma_ts(tmem_base + C::cO + 64, tmem_base + ks * 8, v_tail0 + (ks * 512 >> 4), idesc_pv_tail, ks > 0);
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================ softmax / epilogue ============================
    const int e = warp - 4;
    const int row = e * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(e * 32) << 16;
    const int m = m0 + row;
    const bool valid = m < S;
    const bool warp_valid = (m0 + e * 32) < S;        // warp-uniform: any valid row in this warp
    const int mh = valid ? m / E : 0, mw = valid ? m % E : 0;
    const float c_scale = scale * kLog2e;

    PROF_STAMP(0);
    mbar_wait(t_full, 0);
    PROF_STAMP(1);
    tc_fence_after();
    uint32_t* my_th = sTh + row * C::kBounceWords;
    uint32_t* my_tw = sTw + row * C::kBounceWords;
    float bh[E], bw[E];
    if (warp_valid) {
      uint32_t r[32];
      tmem_ld_x32(tmem_base + C::cTh + lane_off, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) my_th[i] = pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      tmem_ld_x32(tmem_base + C::cTw + lane_off, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) my_tw[i] = pack_h2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_done);
    {
      const __half* th_row = reinterpret_cast<const __half*>(my_th);
      const __half* tw_row = reinterpret_cast<const __half*>(my_tw);
      const int rw = (relw_mode == SAMQ_RELW_UPSTREAM) ? mw : mh;
#pragma unroll
      for (int k = 0; k < E; ++k) {
        bh[k] = warp_valid ? kLog2e * __half2float(th_row[mh - k + E - 1]) : 0.f;
        bw[k] = warp_valid ? kLog2e * __half2float(tw_row[rw - k + E - 1]) : 0.f;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(b_done);
    PROF_STAMP(2);

    mbar_wait(s_full, 0);
    PROF_STAMP(3);
    tc_fence_after();
    const uint32_t s_tmem = tmem_base + C::cS + lane_off;
    float l = 0.f;
    if (warp_valid) {
      // 7 steps of 2 key rows (28 keys): S columns [28i, 28i+28), bias = bh[2i | 2i+1] + bw[kw]
      // ---- pass 1: row maximum over the 196 real keys ----
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        uint32_t r[32];
        tmem_ld_x32(s_tmem + 28 * i, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 28; ++j)
          mx = fmaxf(mx, fmaf(__uint_as_float(r[j]), c_scale, bw[j % E]) + bh[2 * i + (j >= E ? 1 : 0)]);
      }
      PROF_STAMP(4);
      // (the bias is added as fma(s, c, bw) + bh so that nothing but the 28 table values is
      // loop-invariant: summing bh + bw first made the compiler keep 196 sums alive and spill)
#pragma unroll
      for (int k = 0; k < E; ++k) bh[k] -= mx;
      // ---- pass 2: P = 2^(x - max) as fp16 pairs, written behind the read pointer ----
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        uint32_t r[32];
        tmem_ld_x32(s_tmem + 28 * i, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 28; j += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(r[j]), c_scale, bw[j % E]) + bh[2 * i + (j >= E ? 1 : 0)]);
          const float p1 = ex2(fmaf(__uint_as_float(r[j + 1]), c_scale, bw[(j + 1) % E]) + bh[2 * i + (j + 1 >= E ? 1 : 0)]);
          l += p0 + p1;
          pk[j >> 1] = pack_h2(p0, p1);
        }
        pk[14] = 0;   // the two extra columns belong to the next step (rewritten there) or are
        pk[15] = 0;   // the zero padding after key 195
        tmem_st_x16(s_tmem + 14 * i, pk);   // 14i+15 < 28(i+1): never ahead of the read pointer
      }
      {
        // padded keys 200..207 (P columns 100..103) must be exact zeros for the K = 208 PV MMA
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s_tmem + 100), "r"(0u)
                     : "memory");
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);
    PROF_STAMP(5);

    // ---- epilogue: O / l ----
    mbar_wait(o_full, 0);
    PROF_STAMP(6);
    tc_fence_after();
    if (warp_valid) {
      const float inv_l = 1.f / l;
      const uint32_t o_tmem = tmem_base + C::cO + lane_off;
      __half* dst = out + (static_cast<size_t>(b) * S + (valid ? m : 0)) * D + head * HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_x32(o_tmem + c * 32, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o;
            o.x = pack_h2(__uint_as_float(r[8 * v + 0]) * inv_l, __uint_as_float(r[8 * v + 1]) * inv_l);
            o.y = pack_h2(__uint_as_float(r[8 * v + 2]) * inv_l, __uint_as_float(r[8 * v + 3]) * inv_l);
            o.z = pack_h2(__uint_as_float(r[8 * v + 4]) * inv_l, __uint_as_float(r[8 * v + 5]) * inv_l);
            o.w = pack_h2(__uint_as_float(r[8 * v + 6]) * inv_l, __uint_as_float(r[8 * v + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = o;
          }
        }
      }
      if (C::kTail) {
        uint32_t r[16];
        tmem_ld_x16(o_tmem + 64, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            uint4 o;
            o.x = pack_h2(__uint_as_float(r[8 * v + 0]) * inv_l, __uint_as_float(r[8 * v + 1]) * inv_l);
            o.y = pack_h2(__uint_as_float(r[8 * v + 2]) * inv_l, __uint_as_float(r[8 * v + 3]) * inv_l);
            o.z = pack_h2(__uint_as_float(r[8 * v + 4]) * inv_l, __uint_as_float(r[8 * v + 5]) * inv_l);
            o.w = pack_h2(__uint_as_float(r[8 * v + 6]) * inv_l, __uint_as_float(r[8 * v + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + 64 + v * 8) = o;
          }
        }
      }
    }
  }

  PROF_STAMP(7);
  PROF_FLUSH;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int HD>
int launch_attn_win(const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads, float scale,
                    int relw_mode, cudaStream_t st) {
  using C = WCfg<HD>;
  const int D = heads * HD;
  const uint64_t row_bytes = static_cast<uint64_t>(3) * D * 2;
  uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(C::S), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {row_bytes, row_bytes * C::S};
  uint32_t q_main[3] = {64, 128, 1}, q_tail[3] = {16, 128, 1};
  uint32_t kv_main[3] = {64, static_cast<uint32_t>(C::SP), 1}, kv_tail[3] = {16, static_cast<uint32_t>(C::SP), 1};
  const CUtensorMap* mq = get_tensor_map_nd(qkv, 3, dims, strides, q_main, 2, 3);
  const CUtensorMap* mkv = get_tensor_map_nd(qkv, 3, dims, strides, kv_main, 2, 3);
  const CUtensorMap* mh = get_tensor_map_2d(rph, 27, HD, HD * 2, 32, 64, 2, 3);
  const CUtensorMap* mw = get_tensor_map_2d(rpw, 27, HD, HD * 2, 32, 64, 2, 3);
  if (!mq || !mkv || !mh || !mw) return SAMQ_ERR_LAUNCH;
  const CUtensorMap *mqt = mq, *mkvt = mkv, *mht = mh, *mwt = mw;
  if (C::kTail) {
    mqt = get_tensor_map_nd(qkv, 3, dims, strides, q_tail, 2, 1);
    mkvt = get_tensor_map_nd(qkv, 3, dims, strides, kv_tail, 2, 1);
    mht = get_tensor_map_2d(rph, 27, HD, HD * 2, 32, 16, 2, 1);
    mwt = get_tensor_map_2d(rpw, 27, HD, HD * 2, 32, 16, 2, 1);
    if (!mqt || !mkvt || !mht || !mwt) return SAMQ_ERR_LAUNCH;
  }
  auto kern = attn_win_kernel<HD>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "attn_win"); rc != SAMQ_OK) return rc;
  dim3 grid(2, heads, B);
  kern<<<grid, kAttThreads, C::kSmemBytes, st>>>(*mq, *mqt, *mkv, *mkvt, *mh, *mht, *mw, *mwt,
                                                reinterpret_cast<__half*>(out), heads, scale, relw_mode);
  count_launch();
  return check_launch("attn_win_kernel");
}

// ===========================================================================================
// Windowed attention, third design: PERSISTENT CTAs, one per SM, streaming (window, head) items.
//
// Timeline of the second design (clock64 stamps, tests/micro/attn_prof.cu): per CTA 0.9k clk
// set-up, 3.0k waiting for the first TMA + rel-pos MMAs, 1.0k bias bounce, 1.4k max pass, 3.7k
// exp pass, 1.6k PV, 2.4k epilogue = 14.5k clk, two CTAs per window-head (K / V loaded twice),
// two CTAs per SM: 14.5k clk per item per SM against a MUFU floor of 3.1k.  Here:
//   * a CTA loops over items; the TMA warp runs up to two items ahead (Q, K, V rings of 2), so
//     load latency and set-up are paid once per CTA, and K / V are loaded once per item;
//   * the item's two 128-query tiles (rows 0-127 / 128-195) are two INDEPENDENT pipelines, each a
//     softmax warpgroup plus its own MMA-issuing warp and its own 208-column TMEM region:
//       T = Q.[Rph;Rpw]^T -> bias registers -> S = Q.K^T -> max pass -> exp pass (P over S)
//       -> O = P.V (columns 112..191 of the region) -> O/l -> shared -> TMA store;
//     while one pipeline waits for its MMAs the other one computes;
//   * the bias values are picked straight out of TMEM: all rows of a warp span at most four image
//     rows mh, and T_h[row][mh + 13 - kh] is a 14-column window starting at column mh, so one
//     x16 load per distinct mh plus a predicated move replaces the shared-memory bounce;
//   * O is staged in the (dead) Q slot in the TMA swizzle layouts and written with a TMA store,
//     which also clips the rows beyond token 195.
// ===========================================================================================
constexpr int kWin3Threads = 384;   // warps 0-3 / 4-7: softmax WG of tile A / B, 8: TMA, 9 / 10: MMA of tile A / B

template <int HD>
struct W3Cfg {
  static constexpr int E = 14, S = 196, SP = 208;
  static constexpr int kTail = HD - 64;
  // The window's 196 tokens = 14 rows of 14.  Tile A stores rows 0-8 (tokens 0..125; its MMA rows
  // 126, 127 are computed and ignored), tile B rows 9-13 (tokens 126..195): both are rectangles of
  // the window, so each tile's O is ONE TMA store box {hd, 14, rows} -- into the windowed layout or
  // straight into image order (window_unpartition + crop for free: out-of-image elements of a box
  // are not written).
  static constexpr int kTokB = 126, kRowsA = 126, kValidB = S - kTokB;  // 70
  static constexpr int kRowsB = 72;                                   // Q rows loaded for tile B (>= 70, atoms of 8)
  static constexpr int kQAMain = 128 * 128, kQBMain = kRowsB * 128;
  static constexpr int kQATail = kTail ? 128 * 32 : 0, kQBTail = kTail ? kRowsB * 32 : 0;
  // [QA main | QA tail | QB main | QB tail]: a tile's main + tail are adjacent because its O is
  // staged over both as plain rows of hd fp16 (A: 126 x 2 hd <= 20480 / 16384 B, B: 70 x 2 hd)
  static constexpr int oQA = 0, oQAT = oQA + kQAMain, oQB = oQAT + kQATail, oQBT = oQB + kQBMain;
  static constexpr int kQStage = ((oQBT + kQBTail + 1023) / 1024) * 1024;
  static_assert(kRowsA * HD * 2 <= kQAMain + kQATail && kValidB * HD * 2 <= kQBMain + kQBTail, "O staging fits");
  static constexpr int kKVMain = SP * 128, kKVTail = kTail ? SP * 32 : 0;
  static constexpr int oKM = 0, oVM = kKVMain, oKT = 2 * kKVMain, oVT = oKT + kKVTail;
  static constexpr int kKVStage = ((oVT + kKVTail + 1023) / 1024) * 1024;
  static constexpr int kRpMain = 32 * 128, kRpTail = kTail ? 32 * 32 : 0;
  // [Rph main | Rpw main | Rph tail | Rpw tail]: the two tables form one 64-row B operand
  static constexpr int oRp = 0;
  static constexpr int kRpBytes = ((2 * kRpMain + 2 * kRpTail + 1023) / 1024) * 1024;
  static constexpr int oQ = oRp + kRpBytes;
  static constexpr int oKV = oQ + 2 * kQStage;
  static constexpr int oL = oKV + 2 * kKVStage;
  static constexpr int oBars = oL;
  static constexpr int kNumBars = 1 + 6 * 2 + 6 * 2;
  static constexpr int kSmemBytes = oBars + kNumBars * 8 + 16 + 1024;
  static constexpr int cO = 112;                                      // O columns inside a region
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int HD>
__global__ void __launch_bounds__(kWin3Threads, 1)
attn_win3_kernel(const __grid_constant__ CUtensorMap map_qa_main, const __grid_constant__ CUtensorMap map_qa_tail,
                 const __grid_constant__ CUtensorMap map_qb_main, const __grid_constant__ CUtensorMap map_qb_tail,
                 const __grid_constant__ CUtensorMap map_kv_main, const __grid_constant__ CUtensorMap map_kv_tail,
                 const __grid_constant__ CUtensorMap map_rph_main, const __grid_constant__ CUtensorMap map_rph_tail,
                 const __grid_constant__ CUtensorMap map_rpw_main, const __grid_constant__ CUtensorMap map_rpw_tail,
                 const __grid_constant__ CUtensorMap map_o_a, const __grid_constant__ CUtensorMap map_o_b,
                 int heads, int n_items, float scale, int relw_mode, int img_nh, int img_nw, int exact_max) {
  using C = W3Cfg<HD>;
  constexpr int E = C::E, SP = C::SP;
  PROF_DECL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sRp = smem + C::oRp;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBars);
  uint64_t* rp_full = bars;
  uint64_t* q_full = bars + 1;       // [2 stages]
  uint64_t* q_empty = q_full + 2;    // count 2: both tiles' O stores have read the slot
  uint64_t* k_full = q_empty + 2;
  uint64_t* k_empty = k_full + 2;    // count 2: both tiles' QK^T retired
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 2;    // count 2: both tiles' PV retired
  uint64_t* t_full = v_empty + 2;    // [2 tiles] from here on
  uint64_t* t_done = t_full + 2;     // count 4
  uint64_t* s_full = t_done + 2;
  uint64_t* p_full = s_full + 2;     // count 4
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;     // count 4: O has been read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;

  if (warp == 9 && lane == 0) {
    mbar_init(rp_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2);
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      mbar_init(&t_full[i], 1); mbar_init(&t_done[i], 4);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(rp_full, 2 * (C::kRpMain + C::kRpTail));
      tma_load_2d(sRp, &map_rph_main, rp_full, 0, 0);
      tma_load_2d(sRp + C::kRpMain, &map_rpw_main, rp_full, 0, 0);
      if (C::kTail) {
        tma_load_2d(sRp + 2 * C::kRpMain, &map_rph_tail, rp_full, 64, 0);
        tma_load_2d(sRp + 2 * C::kRpMain + C::kRpTail, &map_rpw_tail, rp_full, 64, 0);
      }
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        const int b = item / heads, head = item % heads;
        uint8_t* sQ = smem + C::oQ + st * C::kQStage;
        uint8_t* sKV = smem + C::oKV + st * C::kKVStage;
        mbar_wait(&q_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&q_full[st], C::kQAMain + C::kQBMain + C::kQATail + C::kQBTail);
        tma_load_3d(sQ + C::oQA, &map_qa_main, &q_full[st], head * HD, 0, b);
        tma_load_3d(sQ + C::oQB, &map_qb_main, &q_full[st], head * HD, C::kTokB, b);
        if (C::kTail) {
          tma_load_3d(sQ + C::oQAT, &map_qa_tail, &q_full[st], head * HD + 64, 0, b);
          tma_load_3d(sQ + C::oQBT, &map_qb_tail, &q_full[st], head * HD + 64, C::kTokB, b);
        }
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], C::kKVMain + C::kKVTail);
        tma_load_3d(sKV + C::oKM, &map_kv_main, &k_full[st], D + head * HD, 0, b);
        if (C::kTail) tma_load_3d(sKV + C::oKT, &map_kv_tail, &k_full[st], D + head * HD + 64, 0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], C::kKVMain + C::kKVTail);
        tma_load_3d(sKV + C::oVM, &map_kv_main, &v_full[st], 2 * D + head * HD, 0, b);
        if (C::kTail) tma_load_3d(sKV + C::oVT, &map_kv_tail, &v_full[st], 2 * D + head * HD + 64, 0, b);
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ============================ MMA issuer of tile X ============================
    const int X = warp - 9;
    const uint32_t region = tmem_base + X * SP;
    constexpr uint32_t idesc_t = make_idesc_f16(128, 64, 0);
    constexpr uint32_t idesc_qk = make_idesc_f16(128, SP, 0);
    constexpr uint32_t idesc_pv_main = make_idesc_f16(128, 64, 1);
    constexpr uint32_t idesc_pv_tail = make_idesc_f16(128, 16, 1);
    const uint64_t rp_main = make_smem_desc(smem_u32(sRp), 0, 1024, kLayoutSw128);
    const uint64_t rp_tail = make_smem_desc(smem_u32(sRp + 2 * C::kRpMain), 0, 256, kLayoutSw32);
    mbar_wait(rp_full, 0);
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int st = n & 1;
      const uint32_t ph = (n >> 1) & 1, pn = n & 1;
      const uint8_t* sQ = smem + C::oQ + st * C::kQStage;
      const uint8_t* sKV = smem + C::oKV + st * C::kKVStage;
      const uint64_t q_main = make_smem_desc(smem_u32(sQ + (X ? C::oQB : C::oQA)), 0, 1024, kLayoutSw128);
      const uint64_t q_tail = make_smem_desc(smem_u32(sQ + (X ? C::oQBT : C::oQAT)), 0, 256, kLayoutSw32);
      // D[128, N] = Q . B^T for a K-major B tile (rel-pos tables or K)
      auto mma_q_times = [&](uint64_t b_main, uint64_t b_tail, uint32_t idesc, uint64_t* bar0, uint64_t* bar1) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_ss(region, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc, k > 0);
          if (C::kTail) tc_mma_ss(region, q_tail, b_tail, idesc, 1);
          tc_commit(bar0);
          if (bar1) tc_commit(bar1);
        }
        __syncwarp();
      };
      // rel-pos tables into columns [0, 64) of the region (free once the previous O was read)
      PROF_BEGIN;
      mbar_wait(&q_full[st], ph);
      PROF_END(0);
      PROF_BEGIN;
      mbar_wait(&o_free[X], pn ^ 1);
      PROF_END(1);
      tc_fence_after();
      mma_q_times(rp_main, rp_tail, idesc_t, &t_full[X], nullptr);
#ifdef SAMQ_ATTN_PROFILE
      PROF_BEGIN;
      mbar_wait(&t_full[X], pn);
      PROF_END(2);
#endif
      // S = Q K^T over the whole region once the bias values have been read out
      PROF_BEGIN;
      mbar_wait(&k_full[st], ph);
      mbar_wait(&t_done[X], pn);
      PROF_END(3);
      tc_fence_after();
      mma_q_times(make_smem_desc(smem_u32(sKV + C::oKM), 0, 1024, kLayoutSw128),
                  make_smem_desc(smem_u32(sKV + C::oKT), 0, 256, kLayoutSw32), idesc_qk, &s_full[X], &k_empty[st]);
#ifdef SAMQ_ATTN_PROFILE
      PROF_BEGIN;
      mbar_wait(&s_full[X], pn);
      PROF_END(4);
#endif
      // O = P V
      PROF_BEGIN;
      mbar_wait(&v_full[st], ph);
      mbar_wait(&p_full[X], pn);
      PROF_END(5);
      tc_fence_after();
      const uint64_t v_main0 = make_smem_desc(smem_u32(sKV + C::oVM), C::kKVMain, 1024, kLayoutSw128);
      const uint64_t v_tail0 = make_smem_desc(smem_u32(sKV + C::oVT), 4096, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < SP / 16; ++ks) {
          tc_mma_ts(region + C::cO, region + ks * 8, v_main0 + (ks * 2048 >> 4), idesc_pv_main, ks > 0);
          if (C::kTail)
            tc_mma_ts(region + C::cO + 64, region + ks * 8, v_tail0 + (ks * 512 >> 4), idesc_pv_tail, ks > 0);
        }
        tc_commit(&o_full[X]);
        tc_commit(&v_empty[st]);
      }
      __syncwarp();
#ifdef SAMQ_ATTN_PROFILE
      PROF_BEGIN;
      mbar_wait(&o_full[X], pn);
      PROF_END(6);
#endif
    }
  } else if (warp < 8) {
    // ============================ softmax warpgroup of tile X ============================
    const int X = warp >> 2;
    const int e = warp & 3;
    const int row = e * 32 + lane;                    // row of the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(e * 32) << 16;
    const uint32_t region = tmem_base + X * SP + lane_off;
    const int m = X * C::kTokB + row;                 // token inside the window
    const int n_valid = X ? C::kValidB : C::kRowsA;   // rows of this tile that are stored
    const bool valid = row < n_valid;
    const bool warp_valid = e * 32 < n_valid;         // warp-uniform: any valid row in this warp
    const int mh = valid ? m / E : 0, mw = valid ? m % E : 0;
    // distinct table windows needed by this warp: image rows of its first / last valid token
    const int m_first = X * C::kTokB + e * 32, m_last = X * C::kTokB + min(e * 32 + 31, n_valid - 1);
    const int vh_lo = m_first / E, vh_hi = m_last / E;
    float c_scale = scale * kLog2e;
    asm volatile("mov.b32 %0, %0;" : "+f"(c_scale));
    // this row's O staging address inside the tile's Q slot: plain rows of hd fp16
    const uint32_t o_row_off = (X ? C::oQB : C::oQA) + row * (HD * 2);

    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int st = n & 1;
      const uint32_t pn = n & 1;
      const int b = item / heads, head = item % heads;
      uint8_t* sQ = smem + C::oQ + st * C::kQStage;

      // ---- bias values out of TMEM: bh[k] = T_h[row][mh + 13 - k], bw[k] = T_w[row][rw + 13 - k],
      // rounded through fp16 (the reference forms fp16 rel-pos products), times log2(e) ----
      float bh[E], bw[E];
#pragma unroll
      for (int k = 0; k < E; ++k) bh[k] = bw[k] = 0.f;
      float bw_hi = 0.f;
      bool bound_ok = false;
      PROF_BEGIN;
      mbar_wait(&t_full[X], pn);
      PROF_END(0);
      PROF_BEGIN;
      tc_fence_after();
      if (warp_valid) {
        if (relw_mode != SAMQ_RELW_UPSTREAM) {
          // both tables are indexed by the image row: one pass over the warp's (<= 4) image rows
#pragma unroll 1
          for (int v = vh_lo; v <= vh_hi; ++v) {
            uint32_t rh[16], rv[16];
            tmem_ld_x16(region + v, rh);
            tmem_ld_x16(region + 32 + v, rv);
            tmem_ld_wait();
            if (mh == v) {
#pragma unroll
              for (int k = 0; k < E; ++k) {
                bh[k] = __uint_as_float(rh[13 - k]);
                bw[k] = __uint_as_float(rv[13 - k]);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int v = vh_lo; v <= vh_hi; ++v) {
            uint32_t r[16];
            tmem_ld_x16(region + v, r);
            tmem_ld_wait();
            if (mh == v) {
#pragma unroll
              for (int k = 0; k < E; ++k) bh[k] = __uint_as_float(r[13 - k]);
            }
          }
#pragma unroll 1
          for (int v = 0; v < E; ++v) {
            uint32_t r[16];
            tmem_ld_x16(region + 32 + v, r);
            tmem_ld_wait();
            if (mw == v) {
#pragma unroll
              for (int k = 0; k < E; ++k) bw[k] = __uint_as_float(r[13 - k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) {
          bh[k] = kLog2e * __half2float(__float2half_rn(bh[k]));
          bw[k] = kLog2e * __half2float(__float2half_rn(bw[k]));
        }
        // see attn_glob3_kernel: where the 14 column biases of every row of this warp lie within
        // 15 (log2 units) of each other the row maximum is replaced by the bound
        // max_k(scale * s + bh) + max(bw) - min(spread, 7), which needs no per-element FMA
        float bw_lo = bw[0];
        bw_hi = bw[0];
#pragma unroll
        for (int k = 1; k < E; ++k) { bw_hi = fmaxf(bw_hi, bw[k]); bw_lo = fminf(bw_lo, bw[k]); }
        bound_ok = __all_sync(0xffffffffu, !valid || bw_hi - bw_lo <= 15.f) && !exact_max;
        bw_hi -= fminf(bw_hi - bw_lo, 7.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_done[X]);
      PROF_END(1);
      PROF_BEGIN;

      // ---- softmax over the 196 real keys: 7 steps of 2 key rows (28 keys) ----
      // one-time half-period skew: tile B starts its first softmax when tile A has finished its
      // first, so that afterwards one pipeline's exp pass (MUFU-bound) overlaps the other's
      // MMA waits / max pass / epilogue instead of colliding with its exp pass
      if (X == 1 && n == 0) mbar_wait(&p_full[0], 0);
      mbar_wait(&s_full[X], pn);
      PROF_END(2);
      PROF_BEGIN;
      tc_fence_after();
      float l = 0.f;
      if (warp_valid) {
        // Both passes loop over PAIRS of steps at run time (steps 2 ii, 2 ii + 1; the TMEM load of
        // the next step is in flight while a step is processed) instead of being unrolled seven
        // times: fully unrolled, the kernel's hot body was ~2.4k instructions and ran at an 82 %
        // instruction-cache hit rate.  The step's two bias values are picked with selects.
        uint32_t ra[32], rb[32];
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        tmem_ld_x32(region, ra);
        if (bound_ok) {
#pragma unroll 1
          for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
              const int i = 2 * ii + sb;
              if (i < 7) {
                uint32_t (&r)[32] = sb ? rb : ra;
                tmem_ld_wait();
                if (i < 6) tmem_ld_x32(region + 28 * (i + 1), sb ? ra : rb);
                const float ba = ii < 2 ? (ii == 0 ? bh[2 * sb] : bh[4 + 2 * sb]) : (ii == 2 ? bh[8 + 2 * sb] : bh[12]);
                const float bb = ii < 2 ? (ii == 0 ? bh[1 + 2 * sb] : bh[5 + 2 * sb]) : (ii == 2 ? bh[9 + 2 * sb] : bh[13]);
                float ua = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
                float ub = fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15]));
#pragma unroll
                for (int j = 2; j < E; j += 2) {
                  ua = fmaxf(fmaxf(ua, __uint_as_float(r[j])), __uint_as_float(r[j + 1]));
                  ub = fmaxf(fmaxf(ub, __uint_as_float(r[E + j])), __uint_as_float(r[E + j + 1]));
                }
                m0 = fmaxf(m0, fmaf(ua, c_scale, ba));
                m1 = fmaxf(m1, fmaf(ub, c_scale, bb));
              }
            }
          }
          m0 += bw_hi;
          m1 += bw_hi;
        } else {
  #pragma unroll 1
          for (int ii = 0; ii < 4; ++ii) {
  #pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
              const int i = 2 * ii + sb;
              if (i < 7) {
                uint32_t (&r)[32] = sb ? rb : ra;
                tmem_ld_wait();
                if (i < 6) tmem_ld_x32(region + 28 * (i + 1), sb ? ra : rb);
                const float ba = ii < 2 ? (ii == 0 ? bh[2 * sb] : bh[4 + 2 * sb]) : (ii == 2 ? bh[8 + 2 * sb] : bh[12]);   // ii == 3: only step 6 exists
                const float bb = ii < 2 ? (ii == 0 ? bh[1 + 2 * sb] : bh[5 + 2 * sb]) : (ii == 2 ? bh[9 + 2 * sb] : bh[13]);
  #pragma unroll
                for (int j = 0; j < 28; j += 4) {
                  m0 = fmaxf(m0, fmaf(__uint_as_float(r[j + 0]), c_scale, bw[(j + 0) % E]) + ((j + 0) >= E ? bb : ba));
                  m1 = fmaxf(m1, fmaf(__uint_as_float(r[j + 1]), c_scale, bw[(j + 1) % E]) + ((j + 1) >= E ? bb : ba));
                  m2 = fmaxf(m2, fmaf(__uint_as_float(r[j + 2]), c_scale, bw[(j + 2) % E]) + ((j + 2) >= E ? bb : ba));
                  m3 = fmaxf(m3, fmaf(__uint_as_float(r[j + 3]), c_scale, bw[(j + 3) % E]) + ((j + 3) >= E ? bb : ba));
                }
              }
            }
          }
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        PROF_END(3);
        PROF_BEGIN;
        if (n > 0 && e == 0 && lane == 0) {
          // the previous item's O store was queued ~2.5k clk ago: it has read its shared-memory
          // source by now, so its Q slot can go back to the TMA producer
          tma_store_wait_read<0>();
          mbar_arrive(&q_empty[st ^ 1]);
        }
#pragma unroll
        for (int k = 0; k < E; ++k) bh[k] -= mx;
        // P = 2^(x - max) as fp16 pairs, written behind the read pointer
        float l0 = 0.f, l1 = 0.f;
        tmem_ld_x32(region, ra);
#pragma unroll 1
        for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {
            const int i = 2 * ii + sb;
            if (i < 7) {
              uint32_t (&r)[32] = sb ? rb : ra;
              tmem_ld_wait();
              if (i < 6) tmem_ld_x32(region + 28 * (i + 1), sb ? ra : rb);
              const float ba = ii < 2 ? (ii == 0 ? bh[2 * sb] : bh[4 + 2 * sb]) : (ii == 2 ? bh[8 + 2 * sb] : bh[12]);   // ii == 3: only step 6 exists
              const float bb = ii < 2 ? (ii == 0 ? bh[1 + 2 * sb] : bh[5 + 2 * sb]) : (ii == 2 ? bh[9 + 2 * sb] : bh[13]);
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 28; j += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(r[j]), c_scale, bw[j % E]) + (j >= E ? bb : ba));
                const float p1 = ex2(fmaf(__uint_as_float(r[j + 1]), c_scale, bw[(j + 1) % E]) + (j + 1 >= E ? bb : ba));
                l0 += p0;
                l1 += p1;
                pk[j >> 1] = pack_h2(p0, p1);
              }
              pk[14] = 0;   // the two extra columns belong to the next step (rewritten there) or are
              pk[15] = 0;   // the zero padding after key 195
              // P columns [14i, 14i+16) lie behind both this step's and the prefetched step's S columns
              tmem_st_x16(region + 14 * i, pk);
            }
          }
        }
        l = l0 + l1;
        // padded keys 200..207 (P columns 100..103) must be exact zeros for the K = 208 PV MMA
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(region + 100), "r"(0u)
                     : "memory");
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[X]);
      PROF_END(4);
      PROF_BEGIN;
#ifdef SAMQ_ATTN_STAMPS
      if (n >= 5 && n <= 8) PROF_STAMP(2 * (n - 5) + 1);
#endif

      // ---- O / l -> fp16 -> the item's Q slot (dead since S was formed) -> TMA store ----
      mbar_wait(&o_full[X], pn);
      PROF_END(5);
      PROF_BEGIN;
      tc_fence_after();
      {
        const float inv_l = warp_valid ? 1.f / l : 0.f;
        const uint32_t o_tmem = region + C::cO;
        uint4* o_row = reinterpret_cast<uint4*>(sQ + o_row_off);
        auto pack8 = [&](const uint32_t* r) {
          uint4 o;
          o.x = pack_h2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
          o.y = pack_h2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
          o.z = pack_h2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
          o.w = pack_h2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
          return o;
        };
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_x32(o_tmem + c * 32, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int v = 0; v < 4; ++v) o_row[c * 4 + v] = pack8(r + 8 * v);
          }
        }
        if (C::kTail) {
          uint32_t r[16];
          tmem_ld_x16(o_tmem + 64, r);
          tmem_ld_wait();
          if (valid) {
            o_row[8] = pack8(r);
            o_row[9] = pack8(r + 8);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[X]);         // the region may be overwritten
      PROF_END(6);
      PROF_BEGIN;
      fence_proxy_async_smem();                       // staging writes -> visible to the TMA engine
      named_bar_sync(1 + X, 128);
      if (e == 0 && lane == 0) {
        // queued without waiting; the slot is released during the next item (see the max pass).
        // One box {hd, 14 columns, 9 | 5 rows} of the window: windowed layout (img_nw == 0) or the
        // window's place in the [B, H, W, D] image (out-of-image rows / columns are clipped).
        int c1 = 0, c2 = X ? 9 : 0, c3 = b;
        if (img_nw > 0) {
          const int ww = b % img_nw, t = b / img_nw;
          c1 = ww * E;
          c2 += (t % img_nh) * E;
          c3 = t / img_nh;
        }
        tma_store_4d(X ? &map_o_b : &map_o_a, sQ + (X ? C::oQB : C::oQA), head * HD, c1, c2, c3);
        tma_store_commit();
      }
      PROF_END(7);
    }
    if (e == 0 && lane == 0) tma_store_wait_all<0>();
  }

  PROF_FLUSH;
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// SAMQ_ATTN_MAX=exact (developer switch, A/B timing and tests): the softmax warps take the exact
// row maximum even where the bound on it would do.
static int exact_max_requested() {
  const char* v = getenv("SAMQ_ATTN_MAX");
  return v && !strcmp(v, "exact");
}

template <int HD>
int launch_attn_win3(const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads, float scale,
                     int relw_mode, int img_h, int img_w, cudaStream_t st) {
  using C = W3Cfg<HD>;
  const int D = heads * HD;
  const uint64_t row_bytes = static_cast<uint64_t>(3) * D * 2;
  uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(C::S), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {row_bytes, row_bytes * C::S};
  uint32_t qa_main[3] = {64, 128, 1}, qa_tail[3] = {16, 128, 1};
  uint32_t qb_main[3] = {64, C::kRowsB, 1}, qb_tail[3] = {16, C::kRowsB, 1};
  uint32_t kv_main[3] = {64, static_cast<uint32_t>(C::SP), 1}, kv_tail[3] = {16, static_cast<uint32_t>(C::SP), 1};
  const CUtensorMap* mqa = get_tensor_map_nd(qkv, 3, dims, strides, qa_main, 2, 3);
  const CUtensorMap* mqb = get_tensor_map_nd(qkv, 3, dims, strides, qb_main, 2, 3);
  const CUtensorMap* mkv = get_tensor_map_nd(qkv, 3, dims, strides, kv_main, 2, 3);
  const CUtensorMap* mh = get_tensor_map_2d(rph, 27, HD, HD * 2, 32, 64, 2, 3);
  const CUtensorMap* mw = get_tensor_map_2d(rpw, 27, HD, HD * 2, 32, 64, 2, 3);
  // O: 4-D view (d, column, row, window | image) of the windowed [B, 14, 14, D] or of the image-order
  // [B / (nH nW), img_h, img_w, D] output; tile A stores window rows 0-8, tile B rows 9-13
  const int img_nh = img_h > 0 ? (img_h + C::E - 1) / C::E : 0, img_nw = img_w > 0 ? (img_w + C::E - 1) / C::E : 0;
  const uint64_t ow = img_w > 0 ? img_w : C::E, oh = img_h > 0 ? img_h : C::E;
  const uint64_t ob = img_w > 0 ? static_cast<uint64_t>(B) / (img_nh * img_nw) : static_cast<uint64_t>(B);
  uint64_t odims[4] = {static_cast<uint64_t>(D), ow, oh, ob};
  uint64_t ostrides[3] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(D) * 2 * ow, static_cast<uint64_t>(D) * 2 * ow * oh};
  uint32_t box_a[4] = {static_cast<uint32_t>(HD), static_cast<uint32_t>(C::E), 9, 1};
  uint32_t box_b[4] = {static_cast<uint32_t>(HD), static_cast<uint32_t>(C::E), 5, 1};
  const CUtensorMap* moa = get_tensor_map_nd(out, 4, odims, ostrides, box_a, 2, 0);
  const CUtensorMap* mob = get_tensor_map_nd(out, 4, odims, ostrides, box_b, 2, 0);
  if (!mqa || !mqb || !mkv || !mh || !mw || !moa || !mob) return SAMQ_ERR_LAUNCH;
  const CUtensorMap *mqat = mqa, *mqbt = mqb, *mkvt = mkv, *mht = mh, *mwt = mw;
  if (C::kTail) {
    mqat = get_tensor_map_nd(qkv, 3, dims, strides, qa_tail, 2, 1);
    mqbt = get_tensor_map_nd(qkv, 3, dims, strides, qb_tail, 2, 1);
    mkvt = get_tensor_map_nd(qkv, 3, dims, strides, kv_tail, 2, 1);
    mht = get_tensor_map_2d(rph, 27, HD, HD * 2, 32, 16, 2, 1);
    mwt = get_tensor_map_2d(rpw, 27, HD, HD * 2, 32, 16, 2, 1);
    if (!mqat || !mqbt || !mkvt || !mht || !mwt) return SAMQ_ERR_LAUNCH;
  }
  auto kern = attn_win3_kernel<HD>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "attn_win3"); rc != SAMQ_OK) return rc;
  const int num_sms = device_sm_count();
  const int n_items = B * heads;
  dim3 grid(n_items < num_sms ? n_items : num_sms);
  kern<<<grid, kWin3Threads, C::kSmemBytes, st>>>(*mqa, *mqat, *mqb, *mqbt, *mkv, *mkvt, *mh, *mht, *mw, *mwt, *moa, *mob,
                                                 heads, n_items, scale, relw_mode, img_nh, img_nw, exact_max_requested());
  count_launch();
  return check_launch("attn_win3_kernel");
}

// ===========================================================================================
// Global (64x64) attention, second design: two softmax warpgroups, software-pipelined.
//
// What the first design (attn_relpos_kernel<HD, false>) lost, measured with the clock64()
// breakdown in tests/micro/attn_prof.cu: per 128-key tile the softmax warps spent ~900 clk in
// "TMEM load -> scale + bias -> max" and ~1300 clk in "ex2 -> pack -> TMEM store", strictly one
// after the other (every warp is in the same phase at the same time), against a MUFU floor of
// 1024 clk and 640 clk of MMA; with head_dim 80 only two K/V stages fitted and the MMA warp
// additionally waited ~1600 clk per tile for K.  This design:
//   * a 128-key tile is two key rows kh = 2j, 2j+1 of the image; warpgroup g (warps 0-3 / 4-7)
//     owns key row 2j+g, i.e. S columns [64g, 64g+64), for all 128 query rows, so a thread's
//     whole share of a tile (64 scores) lives in registers: S is read from TMEM ONCE;
//   * the scores of tile j+1 are fetched from TMEM before the ex2 phase of tile j and their
//     scale / bias / max arithmetic is interleaved with that phase's MUFU stream;
//   * S is triple-buffered in TMEM, so QK^T runs two tiles ahead of the softmax;
//   * K and V have separate 3-slot rings (a K slot is released as soon as its QK^T retires);
//     slot 2 of both aliases the rel-pos tables, which are dead after the prologue MMAs;
//   * the bias tables go TMEM -> shared exactly once: bh as fp32 [key row][query] and bw as fp32
//     [query][key col] in XOR-swizzled 16-byte chunks (conflict-free LDS.128).
// The two threads of a query row exchange partial maxima through shared memory once per tile
// (one 256-thread named barrier); partial row sums are combined at the end.
// ===========================================================================================
// tcgen05.ld 32x32b.x32 straight into a slice of a float array (the instruction is .b32-typed)
__device__ __forceinline__ void tmem_ld_x32f(uint32_t taddr, float (&r)[64], int o) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(r[o + 0]), "=f"(r[o + 1]), "=f"(r[o + 2]), "=f"(r[o + 3]), "=f"(r[o + 4]), "=f"(r[o + 5]),
        "=f"(r[o + 6]), "=f"(r[o + 7]), "=f"(r[o + 8]), "=f"(r[o + 9]), "=f"(r[o + 10]), "=f"(r[o + 11]),
        "=f"(r[o + 12]), "=f"(r[o + 13]), "=f"(r[o + 14]), "=f"(r[o + 15]), "=f"(r[o + 16]), "=f"(r[o + 17]),
        "=f"(r[o + 18]), "=f"(r[o + 19]), "=f"(r[o + 20]), "=f"(r[o + 21]), "=f"(r[o + 22]), "=f"(r[o + 23]),
        "=f"(r[o + 24]), "=f"(r[o + 25]), "=f"(r[o + 26]), "=f"(r[o + 27]), "=f"(r[o + 28]), "=f"(r[o + 29]),
        "=f"(r[o + 30]), "=f"(r[o + 31])
      : "r"(taddr)
      : "memory");
}
constexpr int kGlobThreads = 384;   // warps 0-3 / 4-7: softmax warpgroups, 8: TMA + TMEM alloc, 9: MMA, 10-11 idle

template <int HD>
struct GCfg {
  static constexpr int E = 64, S = E * E, kQTiles = S / 128, kKVTiles = S / 128;
  static constexpr int kTail = HD - 64;                      // 0 or 16
  static constexpr int kMainBytes = 128 * 128;               // 128 rows x 64 fp16, 128B swizzle
  static constexpr int kTailBytes = kTail ? 128 * 32 : 0;    // 128 rows x 16 fp16, 32B swizzle
  static constexpr int kTileBytes = kMainBytes + kTailBytes; // Q / K / V tile and one rel-pos table
  static constexpr int kSlots = 3;
  // shared memory carve (all tile bases 1024-aligned)
  static constexpr int oQ = 0;
  static constexpr int oRp = oQ + kTileBytes;                // Rph | Rpw, then K slot 2 | V slot 2
  static constexpr int oK = oRp + 2 * kTileBytes;            // K slots 0, 1
  static constexpr int oV = oK + 2 * kTileBytes;             // V slots 0, 1
  static constexpr int oBh = oV + 2 * kTileBytes;            // float [64 key rows][128 queries]
  static constexpr int oBw = oBh + 64 * 128 * 4;             // float [128 queries][64 key cols], swizzled
  static constexpr int oX = oBw + 128 * 64 * 4;              // float xmax[2][2][128], xsum[2][128]
  static constexpr int oBars = oX + 6 * 128 * 4;
  static constexpr int kNumBars = 1 + 4 * kSlots + 2 + 3 + 3 + 2;
  static constexpr int kSmemBytes = oBars + kNumBars * 8 + 16 + 1024;
  static constexpr int cO = 384;                             // TMEM: S buffers at 0 / 128 / 256, O at 384
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int HD>
__global__ void __launch_bounds__(kGlobThreads, 1)
attn_glob_kernel(const __grid_constant__ CUtensorMap map_qkv_main, const __grid_constant__ CUtensorMap map_qkv_tail,
                 const __grid_constant__ CUtensorMap map_rph_main, const __grid_constant__ CUtensorMap map_rph_tail,
                 const __grid_constant__ CUtensorMap map_rpw_main, const __grid_constant__ CUtensorMap map_rpw_tail,
                 __half* __restrict__ out, int heads, float scale, int relw_mode) {
  using C = GCfg<HD>;
  constexpr int E = C::E, S = C::S, T = C::kKVTiles;
  PROF_DECL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + C::oQ;
  uint8_t* sRph = smem + C::oRp;
  uint8_t* sRpw = sRph + C::kTileBytes;
  float* sBh = reinterpret_cast<float*>(smem + C::oBh);
  float* sBw = reinterpret_cast<float*>(smem + C::oBw);
  float* sX = reinterpret_cast<float*>(smem + C::oX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBars);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + 3;
  uint64_t* v_full = k_empty + 3;
  uint64_t* v_empty = v_full + 3;
  uint64_t* t_full = v_empty + 3;
  uint64_t* t_done = t_full + 1;
  uint64_t* s_full = t_done + 1;
  uint64_t* p_full = s_full + 3;
  uint64_t* pv_done = p_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  auto k_slot = [&](int i) -> uint8_t* { return i < 2 ? smem + C::oK + i * C::kTileBytes : sRph; };
  auto v_slot = [&](int i) -> uint8_t* { return i < 2 ? smem + C::oV + i * C::kTileBytes : sRpw; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int D = heads * HD;
  const int m0 = q_tile * 128;

  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
    }
    mbar_init(t_full, 1);
    mbar_init(t_done, 8);
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    // descriptor fetch overlaps barrier init / TMEM allocation (first-load latency is on the
    // critical path of this one-item CTA)
    tma_prefetch_desc(&map_qkv_main);
    tma_prefetch_desc(&map_rph_main);
    tma_prefetch_desc(&map_rpw_main);
    if (C::kTail) {
      tma_prefetch_desc(&map_qkv_tail);
      tma_prefetch_desc(&map_rph_tail);
      tma_prefetch_desc(&map_rpw_tail);
    }
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 384 threads start with 168 registers each; the softmax warpgroups need ~220 (64 scores + 64
  // prefetched scores + 32 packed probabilities), the third warpgroup needs almost none
  // (each setmaxnreg sits at the top of its role branch: ptxas budgets registers per branch)
  if (warp == 8) {
    reg_dealloc<72>();
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 3 * C::kTileBytes);
      tma_load_3d(sQ, &map_qkv_main, q_full, head * HD, m0, b);
      tma_load_2d(sRph, &map_rph_main, q_full, 0, 0);
      tma_load_2d(sRpw, &map_rpw_main, q_full, 0, 0);
      if (C::kTail) {
        tma_load_3d(sQ + C::kMainBytes, &map_qkv_tail, q_full, head * HD + 64, m0, b);
        tma_load_2d(sRph + C::kMainBytes, &map_rph_tail, q_full, 64, 0);
        tma_load_2d(sRpw + C::kMainBytes, &map_rpw_tail, q_full, 64, 0);
      }
      int slot = 0;
      uint32_t ph = 0;
      for (int j = 0; j < T; ++j) {
        if (j == 2) mbar_wait(t_full, 0);     // slot 2 aliases the rel-pos tables
        uint8_t* sK = k_slot(slot);
        uint8_t* sV = v_slot(slot);
        PROF_BEGIN;
        mbar_wait(&k_empty[slot], ph ^ 1);
        PROF_END(0);
        mbar_arrive_expect_tx(&k_full[slot], C::kTileBytes);
        tma_load_3d(sK, &map_qkv_main, &k_full[slot], D + head * HD, j * 128, b);
        if (C::kTail) tma_load_3d(sK + C::kMainBytes, &map_qkv_tail, &k_full[slot], D + head * HD + 64, j * 128, b);
        PROF_BEGIN;
        mbar_wait(&v_empty[slot], ph ^ 1);
        PROF_END(1);
        mbar_arrive_expect_tx(&v_full[slot], C::kTileBytes);
        tma_load_3d(sV, &map_qkv_main, &v_full[slot], 2 * D + head * HD, j * 128, b);
        if (C::kTail)
          tma_load_3d(sV + C::kMainBytes, &map_qkv_tail, &v_full[slot], 2 * D + head * HD + 64, j * 128, b);
        if (++slot == 3) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9) {
    reg_dealloc<72>();
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0);
    constexpr uint32_t idesc_pv_main = make_idesc_f16(128, 64, 1);
    constexpr uint32_t idesc_pv_tail = make_idesc_f16(128, 16, 1);
    const uint64_t q_main = make_smem_desc(smem_u32(sQ), 0, 1024, kLayoutSw128);
    const uint64_t q_tail = make_smem_desc(smem_u32(sQ + C::kMainBytes), 0, 256, kLayoutSw32);
    // D[128 queries, 128] = Q . B^T for a K-major 128-row tile B (K tile or rel-pos table)
    auto mma_q_times = [&](uint32_t d_tmem, const uint8_t* tile, uint64_t* bar0, uint64_t* bar1) {
      const uint64_t b_main = make_smem_desc(smem_u32(tile), 0, 1024, kLayoutSw128);
      const uint64_t b_tail = make_smem_desc(smem_u32(tile + C::kMainBytes), 0, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(d_tmem, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc_qk, k > 0);
        if (C::kTail) tc_mma_ss(d_tmem, q_tail, b_tail, idesc_qk, 1);
        if (bar0) tc_commit(bar0);
        if (bar1) tc_commit(bar1);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    mma_q_times(tmem_base + 0, sRph, nullptr, nullptr);
    mma_q_times(tmem_base + 128, sRpw, t_full, nullptr);
    mbar_wait(t_done, 0);                      // T_h / T_w have been copied out of TMEM
    tc_fence_after();
    for (int i = 0; i < 2; ++i) {
      mbar_wait(&k_full[i], 0);
      tc_fence_after();
      mma_q_times(tmem_base + 128 * i, k_slot(i), &s_full[i], &k_empty[i]);
    }
    int slot = 0, slot2 = 2;                   // slot of tile j / tile j + 2
    uint32_t ph = 0, ph2 = 0;
    for (int j = 0; j < T; ++j) {
      if (j + 2 < T) {
        PROF_BEGIN;
        mbar_wait(&k_full[slot2], ph2);
        PROF_END(0);
        tc_fence_after();
        mma_q_times(tmem_base + 128 * slot2, k_slot(slot2), &s_full[slot2], &k_empty[slot2]);
      }
      PROF_BEGIN;
      mbar_wait(&p_full[slot], ph);
      PROF_END(1);
      PROF_BEGIN;
      mbar_wait(&v_full[slot], ph);
      PROF_END(2);
      tc_fence_after();
      const uint8_t* sV = v_slot(slot);
      const uint32_t p_tmem = tmem_base + 128 * slot;
      const uint64_t v_main0 = make_smem_desc(smem_u32(sV), C::kMainBytes, 1024, kLayoutSw128);
      const uint64_t v_tail0 = make_smem_desc(smem_u32(sV + C::kMainBytes), C::kTailBytes, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          tc_mma_ts(tmem_base + C::cO, p_tmem + ks * 8, v_main0 + (ks * 2048 >> 4), idesc_pv_main, acc);
          if (C::kTail)
            tc_mma_ts(tmem_base + C::cO + 64, p_tmem + ks * 8, v_tail0 + (ks * 512 >> 4), idesc_pv_tail, acc);
        }
        tc_commit(&v_empty[slot]);
        tc_commit(&pv_done[j & 1]);
      }
      __syncwarp();
      if (++slot == 3) { slot = 0; ph ^= 1; }
      if (++slot2 == 3) { slot2 = 0; ph2 ^= 1; }
    }
  } else if (warp >= 10) {
    reg_dealloc<72>();
  } else {
    reg_alloc<216>();
    // ============================ softmax warpgroups ============================
    const int g = warp >> 2;                  // 0: key row 2j, 1: key row 2j+1
    const int e = warp & 3;                   // TMEM lane quadrant
    const int row = e * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(e * 32) << 16;
    const int m = m0 + row;
    const int mh = m / E, mw = m % E;
    const int swz = row & 7;                  // XOR swizzle of this row's 16-byte bw chunks

    // ---- bias tables, TMEM -> shared (rounded through fp16 like the reference's fp16 rel-pos
    // products): warpgroup 0 writes bh[kh][row] = T_h[row][mh - kh + 63], warpgroup 1 writes
    // bw[row][kw] = T_w[row][rw - kw + 63], both pre-multiplied by log2(e) ----
    PROF_STAMP(3);
    mbar_wait(t_full, 0);
    PROF_STAMP(4);
    tc_fence_after();
    {
      const int rw = (relw_mode == SAMQ_RELW_UPSTREAM) ? mw : mh;
      const uint32_t src = tmem_base + 128 * g + lane_off;
      auto put = [&](int kidx, float t) {
        const float v = kLog2e * __half2float(__float2half_rn(t));
        if (g == 0) sBh[kidx * 128 + row] = v;
        else sBw[row * 64 + ((((kidx >> 2) ^ swz) << 2) | (kidx & 3))] = v;
      };
      if (g == 0 || relw_mode != SAMQ_RELW_UPSTREAM) {
        // the 64-entry window starts at column mh for every row of the warp (32 | 64): entry
        // kidx = 63 - i sits in column mh + i, a compile-time register index
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_x32(src + mh + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) put(63 - (c * 32 + i), __uint_as_float(r[i]));
        }
      } else {
        // upstream rel_w semantics: the window start mw differs per row -> predicated scatter
        const int base = rw + E - 1;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_x32(src + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kidx = base - (c * 32 + i);
            if (static_cast<unsigned>(kidx) < static_cast<unsigned>(E)) put(kidx, __uint_as_float(r[i]));
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_done);
    named_bar_sync(1, 256);                   // both tables visible to both warpgroups
    // 16-byte chunk q of this row's bw lives at bw_addr ^ (q << 4) (+128 for the second half)
    uint32_t bw_addr = smem_u32(sBw + row * 64) | (static_cast<uint32_t>(swz) << 4);
    float c_scale = scale * kLog2e;
    // opaque moves: without them ptxas re-derives both values from scratch at every use
    asm volatile("mov.b32 %0, %0;" : "+r"(bw_addr));
    asm volatile("mov.b32 %0, %0;" : "+f"(c_scale));
    auto ld_bw = [&](int q, int hf) -> float4 {
      float4 w;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                   : "r"((bw_addr ^ (q << 4)) + hf * 128));
      return w;
    };

    // shared-memory scalars through explicit ld/st.shared (pointers captured by the lambdas below
    // would otherwise degrade to generic loads)
    const uint32_t bh_addr = smem_u32(sBh + g * 128 + row);            // + tile * 1024 bytes
    const uint32_t x_mine = smem_u32(sX + g * 128 + row), x_other = smem_u32(sX + (1 - g) * 128 + row);
    auto lds = [](uint32_t addr) -> float {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
      return v;
    };
    auto sts = [](uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); };

    // Two score arrays alternate between "current tile" (scaled + column-biased scores x) and
    // "prefetched next tile" (raw S from TMEM, turned into x in place).
    float xa[64], xb[64];
    float mx_raw;
    auto fetch = [&](float (&r)[64], int buf) {
      const uint32_t s_tmem = tmem_base + 128 * buf + lane_off + 64 * g;
      tmem_ld_x32f(s_tmem, r, 0);
      tmem_ld_x32f(s_tmem + 32, r, 32);
    };
    // r <- r * scale*log2e + bw (this thread's 64 key columns); returns the maximum
    auto bias_max = [&](float (&r)[64]) -> float {
      float a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int o = 32 * hf + 4 * q;
          const float4 w = ld_bw(q, hf);
          r[o + 0] = fmaf(r[o + 0], c_scale, w.x);
          r[o + 1] = fmaf(r[o + 1], c_scale, w.y);
          r[o + 2] = fmaf(r[o + 2], c_scale, w.z);
          r[o + 3] = fmaf(r[o + 3], c_scale, w.w);
          a0 = fmaxf(a0, fmaxf(r[o + 0], r[o + 2]));
          a1 = fmaxf(a1, fmaxf(r[o + 1], r[o + 3]));
        }
      }
      return fmaxf(a0, a1);
    };
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
    fetch(xa, 0);
    tmem_ld_wait();
    mx_raw = bias_max(xa);
    PROF_STAMP(5);

    float m_used = -INFINITY, l = 0.f;
    constexpr uint32_t o_cols = HD / 2;       // O columns rescaled / stored by this warpgroup
    int buf = 0, nbuf = 1;                    // S buffer of tile j / tile j + 1
    uint32_t nph = 0;                         // parity of s_full[nbuf] for tile j + 1
    int j = 0;
    // One key tile: x = scores of tile j, nx = landing zone of tile j + 1.  `more` = a next tile
    // exists (the last tile is peeled, so each body is branch-free and can be scheduled freely).
    auto tile_step = [&](float (&x)[64], float (&nx)[64], auto more_tag) {
      constexpr bool more = decltype(more_tag)::value;
      const float bh = lds(bh_addr + j * 1024);
      const float mx = mx_raw + bh;
      const uint32_t xoff = (j & 1) * 1024;
      sts(x_mine + xoff, mx);
      // prefetch the next tile's scores; they land while the maxima are exchanged
      if (more) {
        PROF_BEGIN;
        mbar_wait(&s_full[nbuf], nph);
        PROF_END(0);
        tc_fence_after();
        fetch(nx, nbuf);
      }
      PROF_BEGIN;
      named_bar_sync(1, 256);
      PROF_END(1);
      const float m_new = fmaxf(m_used, fmaxf(mx, lds(x_other + xoff)));
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, m_new > m_used + 8.f)) {
        // lazy rescale; both warpgroups take the same decision (same data), each rescales its
        // half of the O columns
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const float alpha = ex2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
        const uint32_t o_tmem = tmem_base + C::cO + lane_off + g * o_cols;
        {
          uint32_t r[32];
          tmem_ld_x32(o_tmem, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x32(o_tmem, r);
        }
        if (HD == 80) {
          uint32_t r[8];
          tmem_ld_x8(o_tmem + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x8(o_tmem + 32, r);
        }
        tmem_st_wait();
      }
      PROF_BEGIN;
      if (more) tmem_ld_wait();
      const float mm = m_used - bh;
      // ---- P = 2^(x - m) -> P columns [32g, 32g + 32) of the tile's S buffer.  Warpgroup 1's P
      // columns overlap warpgroup 0's S columns [32, 64): all S reads of this tile completed one
      // iteration ago (prefetch), before the barrier above. ----
      // Each group of four ex2 is followed in the source by the scale / bias / max arithmetic of
      // four scores of the NEXT tile: a warp issues in order and the MUFU pipe takes one
      // warp-instruction per 8 clk, so FMA work placed between the ex2 (instead of behind all 64)
      // overlaps with it.  ptxas keeps about half of the interleave (1076 -> 1046 us).
      uint32_t pk[32];
      float sum0 = 0.f, sum1 = 0.f, a0 = -INFINITY, a1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float p0 = ex2v(x[i + 0] - mm), p1 = ex2v(x[i + 1] - mm);
        const float p2 = ex2v(x[i + 2] - mm), p3 = ex2v(x[i + 3] - mm);
        sum0 += p0 + p2;
        sum1 += p1 + p3;
        pk[(i >> 1) + 0] = pack_h2(p0, p1);
        pk[(i >> 1) + 1] = pack_h2(p2, p3);
        if (more) {
          const float4 w = ld_bw((i >> 2) & 7, i >> 5);
          nx[i + 0] = fmaf(nx[i + 0], c_scale, w.x);
          nx[i + 1] = fmaf(nx[i + 1], c_scale, w.y);
          nx[i + 2] = fmaf(nx[i + 2], c_scale, w.z);
          nx[i + 3] = fmaf(nx[i + 3], c_scale, w.w);
          a0 = fmaxf(a0, fmaxf(nx[i + 0], nx[i + 2]));
          a1 = fmaxf(a1, fmaxf(nx[i + 1], nx[i + 3]));
        }
      }
      l += sum0 + sum1;
      if (more) mx_raw = fmaxf(a0, a1);
      tmem_st_x32(tmem_base + 128 * buf + lane_off + 32 * g, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
      PROF_END(2);
      buf = nbuf;
      if (++nbuf == 3) { nbuf = 0; nph ^= 1; }   // tile t lives in buffer t % 3, phase (t / 3) & 1
      ++j;
    };
    for (int jj = 0; jj < T / 2 - 1; ++jj) {
      tile_step(xa, xb, std::true_type{});
      tile_step(xb, xa, std::true_type{});
    }
    tile_step(xa, xb, std::true_type{});
    tile_step(xb, xa, std::false_type{});

    PROF_STAMP(6);
    // ---- epilogue: combine the two partial sums of each row, O / l ----
    sts(x_mine + 2048, l);
    named_bar_sync(1, 256);
    const float inv_l = 1.f / (l + lds(x_other + 2048));
    mbar_wait(&pv_done[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const uint32_t o_tmem = tmem_base + C::cO + lane_off + g * o_cols;
    __half* dst = out + (static_cast<size_t>(b) * S + m) * D + head * HD + g * o_cols;
    {
      uint32_t r[32];
      tmem_ld_x32(o_tmem, r);
      tmem_ld_wait();
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 o;
        o.x = pack_h2(__uint_as_float(r[8 * v + 0]) * inv_l, __uint_as_float(r[8 * v + 1]) * inv_l);
        o.y = pack_h2(__uint_as_float(r[8 * v + 2]) * inv_l, __uint_as_float(r[8 * v + 3]) * inv_l);
        o.z = pack_h2(__uint_as_float(r[8 * v + 4]) * inv_l, __uint_as_float(r[8 * v + 5]) * inv_l);
        o.w = pack_h2(__uint_as_float(r[8 * v + 6]) * inv_l, __uint_as_float(r[8 * v + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + v * 8) = o;
      }
    }
    if (HD == 80) {
      uint32_t r[8];
      tmem_ld_x8(o_tmem + 32, r);
      tmem_ld_wait();
      uint4 o;
      o.x = pack_h2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
      o.y = pack_h2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
      o.z = pack_h2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
      o.w = pack_h2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
      *reinterpret_cast<uint4*>(dst + 32) = o;
    }
  }

  PROF_STAMP(7);
  PROF_FLUSH;
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
int launch_attn_glob(const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads, float scale,
                     int relw_mode, cudaStream_t st) {
  using C = GCfg<HD>;
  const int D = heads * HD;
  const uint64_t row_bytes = static_cast<uint64_t>(3) * D * 2;
  uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(C::S), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {row_bytes, row_bytes * C::S};
  uint32_t box_main[3] = {64, 128, 1}, box_tail[3] = {16, 128, 1};
  const CUtensorMap* m_main = get_tensor_map_nd(qkv, 3, dims, strides, box_main, 2, 3);
  const int rp_rows = 2 * C::E - 1;
  const CUtensorMap* h_main = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, 128, 64, 2, 3);
  const CUtensorMap* w_main = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, 128, 64, 2, 3);
  if (!m_main || !h_main || !w_main) return SAMQ_ERR_LAUNCH;
  const CUtensorMap *m_tail = m_main, *h_tail = h_main, *w_tail = w_main;
  if (C::kTail) {
    m_tail = get_tensor_map_nd(qkv, 3, dims, strides, box_tail, 2, 1);
    h_tail = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, 128, 16, 2, 1);
    w_tail = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, 128, 16, 2, 1);
    if (!m_tail || !h_tail || !w_tail) return SAMQ_ERR_LAUNCH;
  }
  auto kern = attn_glob_kernel<HD>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "attn_glob"); rc != SAMQ_OK) return rc;
  dim3 grid(C::kQTiles, heads, B);
  kern<<<grid, kGlobThreads, C::kSmemBytes, st>>>(*m_main, *m_tail, *h_main, *h_tail, *w_main, *w_tail,
                                                 reinterpret_cast<__half*>(out), heads, scale, relw_mode);
  count_launch();
  return check_launch("attn_glob_kernel");
}

// ===========================================================================================
// Global (64x64) attention, third design: SMALL CTAs, two per SM.
//
// The second design keeps every unit of one SM in lock-step: all eight softmax warps are in the
// same phase at the same time and the CTA's prologue / epilogue (13 % + 4 % of its life) runs with
// the MUFU idle.  Here a CTA is one softmax warpgroup (thread = query row, all 128 keys of a tile,
// two TMEM passes), one TMA warp and one MMA warp, with 97 KB of shared memory, 256 TMEM columns
// and <= 168 registers, so that TWO CTAs share an SM and drift apart: one CTA's exp pass, MMA
// round trips, prologue and epilogue overlap the other's.
//   TMEM   : S [0,128) (P written over it as fp16, behind the read pointer), O [128, 128+hd);
//            the rel-pos tables T_h / T_w occupy [0,128) / [128,256) during the prologue.
//   shared : Q | K slot 0 | K slot 1 (= rel_pos_h until the tables exist) | V (= rel_pos_w) |
//            bh table fp16 [64 key rows][128 queries]
//   S is single-buffered: QK^T(j+1) is issued right behind P.V(j); the ~0.9k clk round trip is
//   hidden by the neighbour CTA.  bw (64 values) lives in registers, read straight from TMEM.
// ===========================================================================================
constexpr int kGlob3Threads = 192;   // warps 0-3: softmax, 4: TMA, 5: MMA + TMEM alloc

template <int HD>
struct G3Cfg {
  static constexpr int E = 64, S = E * E, kQTiles = S / 128, kKVTiles = S / 128;
  static constexpr int kTail = HD - 64;
  static constexpr int kMainBytes = 128 * 128;
  static constexpr int kTailBytes = kTail ? 128 * 32 : 0;
  static constexpr int kTileBytes = kMainBytes + kTailBytes;
  static constexpr int oQ = 0;
  static constexpr int oK0 = oQ + kTileBytes;
  static constexpr int oK1 = oK0 + kTileBytes;      // rel_pos_h first
  static constexpr int oV = oK1 + kTileBytes;       // rel_pos_w first
  static constexpr int oBh = oV + kTileBytes;       // __half [64][128]
  static constexpr int oBars = oBh + 64 * 128 * 2;
  static constexpr int kNumBars = 15;
  static constexpr int kSmemBytes = oBars + kNumBars * 8 + 16 + 1024;
  static constexpr int cO = 128;
  static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
};

template <int HD>
__global__ void __launch_bounds__(kGlob3Threads, 2)
attn_glob3_kernel(const __grid_constant__ CUtensorMap map_qkv_main, const __grid_constant__ CUtensorMap map_qkv_tail,
                  const __grid_constant__ CUtensorMap map_rph_main, const __grid_constant__ CUtensorMap map_rph_tail,
                  const __grid_constant__ CUtensorMap map_rpw_main, const __grid_constant__ CUtensorMap map_rpw_tail,
                  __half* __restrict__ out, int heads, float scale, int relw_mode, int exact_max) {
  using C = G3Cfg<HD>;
  constexpr int E = C::E, S = C::S, T = C::kKVTiles;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + C::oQ;
  uint8_t* sK0 = smem + C::oK0;
  uint8_t* sK1 = smem + C::oK1;
  uint8_t* sV = smem + C::oV;
  __half* sBh = reinterpret_cast<__half*>(smem + C::oBh);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBars);
  uint64_t* q_full = bars;          // Q + both rel-pos tables
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* t_full = bars + 7;
  uint64_t* t_done = bars + 8;      // count 4
  uint64_t* s_full = bars + 9;      // [2]: S buffer of the even / odd half-tiles
  uint64_t* p_full = bars + 11;     // [2], count 4
  uint64_t* o_done = bars + 13;
  uint64_t* pv_done = bars + 14;    // one completion per P.V (only the lazy rescale waits for it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int D = heads * HD;
  const int m0 = q_tile * 128;

  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
    mbar_init(&k_empty[0], 1); mbar_init(&k_empty[1], 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(t_full, 1); mbar_init(t_done, 4);
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&p_full[0], 4); mbar_init(&p_full[1], 4);
    mbar_init(o_done, 1); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 3 * C::kTileBytes);
      tma_load_3d(sQ, &map_qkv_main, q_full, head * HD, m0, b);
      tma_load_2d(sK1, &map_rph_main, q_full, 0, 0);
      tma_load_2d(sV, &map_rpw_main, q_full, 0, 0);
      if (C::kTail) {
        tma_load_3d(sQ + C::kMainBytes, &map_qkv_tail, q_full, head * HD + 64, m0, b);
        tma_load_2d(sK1 + C::kMainBytes, &map_rph_tail, q_full, 64, 0);
        tma_load_2d(sV + C::kMainBytes, &map_rpw_tail, q_full, 64, 0);
      }
      auto load_k = [&](int j) {
        uint8_t* dst = (j & 1) ? sK1 : sK0;
        mbar_arrive_expect_tx(&k_full[j & 1], C::kTileBytes);
        tma_load_3d(dst, &map_qkv_main, &k_full[j & 1], D + head * HD, j * 128, b);
        if (C::kTail) tma_load_3d(dst + C::kMainBytes, &map_qkv_tail, &k_full[j & 1], D + head * HD + 64, j * 128, b);
      };
      auto load_v = [&](int j) {
        mbar_arrive_expect_tx(v_full, C::kTileBytes);
        tma_load_3d(sV, &map_qkv_main, v_full, 2 * D + head * HD, j * 128, b);
        if (C::kTail) tma_load_3d(sV + C::kMainBytes, &map_qkv_tail, v_full, 2 * D + head * HD + 64, j * 128, b);
      };
      load_k(0);
      mbar_wait(t_done, 0);                    // tables (and the upstream-mode scratch) are dead
      load_k(1);
      load_v(0);
      for (int j = 0; j < T; ++j) {
        // K(j+2) into the slot of K(j) once QK^T(j) has retired; V(j+1) once P.V(j) has retired
        if (j + 2 < T) {
          mbar_wait(&k_empty[j & 1], (j >> 1) & 1);
          load_k(j + 2);
        }
        if (j + 1 < T) {
          mbar_wait(v_empty, j & 1);
          load_v(j + 1);
        }
      }
    }
  } else if (warp == 5) {
    // ============================ MMA issuer ============================
    // A 128-key K / V tile is processed as two HALF-TILES of 64 keys (= one key row of the image)
    // with their own S buffers (TMEM columns [0, 64) and [64, 128)): S(h+2) is formed while the
    // softmax warps work on S(h+1), so they go from one half-tile to the next without waiting for
    // a QK^T (with one 128-column S buffer a third of their time was that wait).
    constexpr uint32_t idesc_tab = make_idesc_f16(128, 128, 0);
    constexpr uint32_t idesc_qk = make_idesc_f16(128, 64, 0);
    constexpr uint32_t idesc_pv_main = make_idesc_f16(128, 64, 1);
    constexpr uint32_t idesc_pv_tail = make_idesc_f16(128, 16, 1);
    const uint64_t q_main = make_smem_desc(smem_u32(sQ), 0, 1024, kLayoutSw128);
    const uint64_t q_tail = make_smem_desc(smem_u32(sQ + C::kMainBytes), 0, 256, kLayoutSw32);
    auto mma_q_times = [&](uint32_t d_tmem, const uint8_t* tile, uint32_t idesc, uint64_t* bar0, uint64_t* bar1) {
      const uint64_t b_main = make_smem_desc(smem_u32(tile), 0, 1024, kLayoutSw128);
      const uint64_t b_tail = make_smem_desc(smem_u32(tile + C::kMainBytes), 0, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(d_tmem, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc, k > 0);
        if (C::kTail) tc_mma_ss(d_tmem, q_tail, b_tail, idesc, 1);
        if (bar0) tc_commit(bar0);
        if (bar1) tc_commit(bar1);
      }
      __syncwarp();
    };
    // S(h) = Q . K(tile h / 2, keys 64 (h & 1) ..)^T into S buffer h & 1; the K slot is released
    // by its second half
    auto mma_qk_half = [&](int h) {
      const int hh = h & 1;
      const uint8_t* tile = ((h >> 1) & 1) ? sK1 : sK0;
      const uint64_t b_main = make_smem_desc(smem_u32(tile + hh * 8192), 0, 1024, kLayoutSw128);
      const uint64_t b_tail = make_smem_desc(smem_u32(tile + C::kMainBytes + hh * 2048), 0, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(tmem_base + 64 * hh, q_main + (k * 32 >> 4), b_main + (k * 32 >> 4), idesc_qk, k > 0);
        if (C::kTail) tc_mma_ss(tmem_base + 64 * hh, q_tail, b_tail, idesc_qk, 1);
        tc_commit(&s_full[hh]);
        if (hh) tc_commit(&k_empty[(h >> 1) & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    mma_q_times(tmem_base + 0, sK1, idesc_tab, nullptr, nullptr);        // T_h = Q . rel_pos_h^T
    mma_q_times(tmem_base + 128, sV, idesc_tab, t_full, nullptr);        // T_w = Q . rel_pos_w^T
    mbar_wait(t_done, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    mma_qk_half(0);
    mma_qk_half(1);
    for (int h = 0; h < 2 * T; ++h) {
      const int hh = h & 1, j = h >> 1;
      mbar_wait(&p_full[hh], j & 1);
      if (hh == 0) mbar_wait(v_full, j & 1);
      tc_fence_after();
      const uint64_t v_main0 = make_smem_desc(smem_u32(sV), C::kMainBytes, 1024, kLayoutSw128);
      const uint64_t v_tail0 = make_smem_desc(smem_u32(sV + C::kMainBytes), C::kTailBytes, 256, kLayoutSw32);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int ks = 4 * hh + kk;                       // 16-key step inside the 128-key V tile
          const uint32_t acc = (h > 0 || kk > 0) ? 1u : 0u;
          tc_mma_ts(tmem_base + C::cO, tmem_base + 64 * hh + kk * 8, v_main0 + (ks * 2048 >> 4), idesc_pv_main, acc);
          if (C::kTail) tc_mma_ts(tmem_base + C::cO + 64, tmem_base + 64 * hh + kk * 8, v_tail0 + (ks * 512 >> 4), idesc_pv_tail, acc);
        }
        tc_commit(pv_done);
        if (hh) tc_commit(v_empty);
        if (h == 2 * T - 1) tc_commit(o_done);
      }
      __syncwarp();
      if (h + 2 < 2 * T) {
        // S(h+2) overwrites P(h): the tensor pipe retires in order, P.V(h) was issued above
        if (hh == 0) mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        mma_qk_half(h + 2);
      }
    }
  } else {
    // ============================ softmax warpgroup ============================
    const int e = warp;
    const int row = e * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(e * 32) << 16;
    const uint32_t tm = tmem_base + lane_off;
    const int m = m0 + row;
    const int mh = m / E, mw = m % E;
    float c_scale = scale * kLog2e;
    asm volatile("mov.b32 %0, %0;" : "+f"(c_scale));

    // ---- bias tables out of TMEM (values rounded through fp16 like the reference's fp16 rel-pos
    // products): bh -> shared fp16 [kh][row]; bw[kw] = log2e * T_w[row][rw - kw + 63] -> registers ----
    float bw[E];
    mbar_wait(t_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {           // window of 64 columns starting at mh (warp-uniform)
      uint32_t r[32];
      tmem_ld_x32(tm + mh + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sBh[(63 - (c * 32 + i)) * 128 + row] = __float2half_rn(__uint_as_float(r[i]));
    }
    if (relw_mode != SAMQ_RELW_UPSTREAM) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tm + 128 + mh + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          bw[63 - (c * 32 + i)] = kLog2e * __half2float(__float2half_rn(__uint_as_float(r[i])));
      }
    } else {
      // the window start mw differs per row: bounce the row through shared memory (the K1 | V slots
      // are free between the table MMAs and the first loads into them, which wait for t_done)
      __half* scratch = reinterpret_cast<__half*>(sK1) + row * 128;     // 128 rows x 256 B = 2 x 16 KB
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tm + 128 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          reinterpret_cast<uint32_t*>(scratch)[c * 16 + (i >> 1)] =
              pack_h2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      }
      __syncwarp();
#pragma unroll
      for (int kw = 0; kw < E; ++kw) bw[kw] = kLog2e * __half2float(scratch[mw - kw + E - 1]);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_done);
    named_bar_sync(1, 128);                 // every row's bh column is visible to its own thread: not
                                            // needed for correctness (own row only) but keeps the
                                            // warps together for the first tile
    float bw_max = bw[0], bw_min = bw[0];
#pragma unroll
    for (int kw = 1; kw < E; ++kw) { bw_max = fmaxf(bw_max, bw[kw]); bw_min = fminf(bw_min, bw[kw]); }
    // The true tile maximum lies in [bound - spread, bound], spread = max(bw) - min(bw).  Using
    // bound - min(spread, 7) as the maximum keeps every 2^(x - m) <= 2^(7 + 8 lazy-rescale lag) (fp16
    // holds 2^15) and the row's largest term >= 2^-(spread - 7): with spread <= 15 that is >= 2^-8, so
    // fp16 subnormal rounding (2^-25 absolute) stays below 2^-17 of the largest term.
    const float bw_spread = bw_max - bw_min;
    const bool bound_ok = __all_sync(0xffffffffu, bw_spread <= 15.f) && !exact_max;
    bw_max -= fminf(bw_spread, 7.f);
    const uint32_t bh_addr = smem_u32(sBh + row);
    auto lds_h = [](uint32_t addr) -> float {
      unsigned short v;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
      return __half2float(__ushort_as_half(v));
    };

    float m_used = -INFINITY, l = 0.f;
#pragma unroll 1
    for (int h = 0; h < 2 * T; ++h) {
      // half-tile h = key row h of the image: 64 scores per query row in S buffer h & 1
      const int hh = h & 1;
      const uint32_t ts = tm + 64 * hh;
      const float bh0 = kLog2e * lds_h(bh_addr + h * 256);
      mbar_wait(&s_full[hh], (h >> 1) & 1);
      tc_fence_after();
      // ---- pass 1.  Softmax is shift-invariant and the running maximum only has to keep
      // 2^(x - m) in fp16 range, so where the 64 column biases of every row of this warp lie within
      // 15 (log2 units) of each other the BOUND max(scale * s) + max(bw) + bh, shifted as explained
      // above, does and the 64 FMAs of the exact maximum are skipped; other warps take the exact
      // maximum. ----
      float a0 = -INFINITY, a1 = -INFINITY;
      uint32_t ra[32], rb[32];
      tmem_ld_x32(ts, ra);
      tmem_ld_x32(ts + 32, rb);
      tmem_ld_wait();
      if (bound_ok) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&r)[32] = c ? rb : ra;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            a0 = fmaxf(a0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
            a1 = fmaxf(a1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&r)[32] = c ? rb : ra;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            a0 = fmaxf(a0, fmaf(__uint_as_float(r[i]), c_scale, bw[c * 32 + i]));
            a1 = fmaxf(a1, fmaf(__uint_as_float(r[i + 1]), c_scale, bw[c * 32 + i + 1]));
          }
        }
      }
      const float m_tile = bound_ok ? fmaf(fmaxf(a0, a1), c_scale, bh0) + bw_max : fmaxf(a0, a1) + bh0;
      const float m_new = fmaxf(m_used, m_tile);
      if (h == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, m_new > m_used + 8.f)) {
        // lazy rescale.  P.V(h-1) was issued when this warpgroup finished half-tile h-1 and may
        // still be accumulating into O: wait for its completion first.
        mbar_wait(pv_done, (h - 1) & 1);
        tc_fence_after();
        const float alpha = ex2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
        const uint32_t o_tmem = tm + C::cO;
        // rare path: 8 columns at a time, so that it does not take registers from the common one
#pragma unroll 1
        for (int c = 0; c < HD / 8; ++c) {
          uint32_t r[8];
          tmem_ld_x8(o_tmem + c * 8, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x8(o_tmem + c * 8, r);
        }
        tmem_st_wait();
      }
      // ---- pass 2: P = 2^(x - m) as fp16 pairs over the half-tile's own S columns (the scores
      // are still in registers from pass 1) ----
      const float mm = m_used - bh0;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t (&r)[32] = c ? rb : ra;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(r[i]), c_scale, bw[c * 32 + i]) - mm);
          const float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), c_scale, bw[c * 32 + i + 1]) - mm);
          s0 += p0;
          s1 += p1;
          pk[i >> 1] = pack_h2(p0, p1);
        }
        tmem_st_x16(ts + 16 * c, pk);
      }
      l += s0 + s1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[hh]);
    }

    // ---- epilogue: O / l ----
    mbar_wait(o_done, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const uint32_t o_tmem = tm + C::cO;
    __half* dst = out + (static_cast<size_t>(b) * S + m) * D + head * HD;
    auto pack8 = [&](const uint32_t* r) {
      uint4 o;
      o.x = pack_h2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
      o.y = pack_h2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
      o.z = pack_h2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
      o.w = pack_h2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
      return o;
    };
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_x32(o_tmem + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int v = 0; v < 4; ++v) *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = pack8(r + 8 * v);
    }
    if (C::kTail) {
      uint32_t r[16];
      tmem_ld_x16(o_tmem + 64, r);
      tmem_ld_wait();
      *reinterpret_cast<uint4*>(dst + 64) = pack8(r);
      *reinterpret_cast<uint4*>(dst + 72) = pack8(r + 8);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int HD>
int launch_attn_glob3(const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads, float scale,
                      int relw_mode, cudaStream_t st) {
  using C = G3Cfg<HD>;
  const int D = heads * HD;
  const uint64_t row_bytes = static_cast<uint64_t>(3) * D * 2;
  uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(C::S), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {row_bytes, row_bytes * C::S};
  uint32_t box_main[3] = {64, 128, 1}, box_tail[3] = {16, 128, 1};
  const CUtensorMap* m_main = get_tensor_map_nd(qkv, 3, dims, strides, box_main, 2, 3);
  const int rp_rows = 2 * C::E - 1;
  const CUtensorMap* h_main = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, 128, 64, 2, 3);
  const CUtensorMap* w_main = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, 128, 64, 2, 3);
  if (!m_main || !h_main || !w_main) return SAMQ_ERR_LAUNCH;
  const CUtensorMap *m_tail = m_main, *h_tail = h_main, *w_tail = w_main;
  if (C::kTail) {
    m_tail = get_tensor_map_nd(qkv, 3, dims, strides, box_tail, 2, 1);
    h_tail = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, 128, 16, 2, 1);
    w_tail = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, 128, 16, 2, 1);
    if (!m_tail || !h_tail || !w_tail) return SAMQ_ERR_LAUNCH;
  }
  auto kern = attn_glob3_kernel<HD>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "attn_glob3"); rc != SAMQ_OK) return rc;
  dim3 grid(C::kQTiles, heads, B);
  kern<<<grid, kGlob3Threads, C::kSmemBytes, st>>>(*m_main, *m_tail, *h_main, *h_tail, *w_main, *w_tail,
                                                  reinterpret_cast<__half*>(out), heads, scale, relw_mode, exact_max_requested());
  count_launch();
  return check_launch("attn_glob3_kernel");
}

template <int HD, bool WIN>
int launch_attn(const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads,
                float scale, int relw_mode, cudaStream_t st) {
  using C = ACfg<HD, WIN>;
  const int D = heads * HD;
  const uint64_t row_bytes = static_cast<uint64_t>(3) * D * 2;
  uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(C::S), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {row_bytes, row_bytes * C::S};
  uint32_t box_main[3] = {64, 128, 1};
  uint32_t box_tail[3] = {16, 128, 1};
  const CUtensorMap* m_main = get_tensor_map_nd(qkv, 3, dims, strides, box_main, 2, 3);
  if (!m_main) return SAMQ_ERR_LAUNCH;
  const CUtensorMap* m_tail = m_main;
  const int rp_rows = 2 * C::E - 1;
  const CUtensorMap* h_main = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, C::kRpRows, 64, 2, 3);
  const CUtensorMap* w_main = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, C::kRpRows, 64, 2, 3);
  if (!h_main || !w_main) return SAMQ_ERR_LAUNCH;
  const CUtensorMap* h_tail = h_main;
  const CUtensorMap* w_tail = w_main;
  if (C::kTail) {
    m_tail = get_tensor_map_nd(qkv, 3, dims, strides, box_tail, 2, 1);
    h_tail = get_tensor_map_2d(rph, rp_rows, HD, HD * 2, C::kRpRows, 16, 2, 1);
    w_tail = get_tensor_map_2d(rpw, rp_rows, HD, HD * 2, C::kRpRows, 16, 2, 1);
    if (!m_tail || !h_tail || !w_tail) return SAMQ_ERR_LAUNCH;
  }
  auto kern = attn_relpos_kernel<HD, WIN>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "attn"); rc != SAMQ_OK) return rc;
  dim3 grid(C::kQTiles, heads, B);
  kern<<<grid, kAttThreads, C::kSmemBytes, st>>>(*m_main, *m_tail, *h_main, *h_tail, *w_main, *w_tail,
                                                reinterpret_cast<__half*>(out), heads, scale, relw_mode);
  count_launch();
  return check_launch("attn_relpos_kernel");
}

}  // namespace
}  // namespace samq

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
  const char* wv = getenv("SAMQ_ATTN_WIN");   // ablations: "v1" first design (two key tiles), "v2" one CTA per q-tile
  const bool win_v1 = wv && strcmp(wv, "v1") == 0;
  const bool win_v2 = wv && strcmp(wv, "v2") == 0;
  if (!glob && !win_v1 && !win_v2) {
    return hd == 64 ? launch_attn_win3<64>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, 0, 0, st)
                    : launch_attn_win3<80>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, 0, 0, st);
  }
  const char* gv = getenv("SAMQ_ATTN_GLOB");  // ablations: "v1" first design, "v2" two lock-step warpgroups
  const bool glob_v1 = gv && strcmp(gv, "v1") == 0;
  const bool glob_v2 = gv && strcmp(gv, "v2") == 0;
  if (glob && glob_v2) {
    return hd == 64 ? launch_attn_glob<64>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st)
                    : launch_attn_glob<80>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
  }
  if (glob && !glob_v1) {
    return hd == 64 ? launch_attn_glob3<64>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st)
                    : launch_attn_glob3<80>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
  }
  if (hd == 64) {
    if (glob) return launch_attn<64, false>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
    return win_v1 ? launch_attn<64, true>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st)
                  : launch_attn_win<64>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
  }
  if (glob) return launch_attn<80, false>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
  return win_v1 ? launch_attn<80, true>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st)
                : launch_attn_win<80>(qkv, rel_pos_h, rel_pos_w, out, B, heads, scale, relw_mode, st);
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
  return hd == 64 ? launch_attn_win3<64>(qkv, rel_pos_h, rel_pos_w, out, static_cast<int>(windows), heads, scale,
                                         relw_mode, H, W, st)
                  : launch_attn_win3<80>(qkv, rel_pos_h, rel_pos_w, out, static_cast<int>(windows), heads, scale,
                                         relw_mode, H, W, st);
}
