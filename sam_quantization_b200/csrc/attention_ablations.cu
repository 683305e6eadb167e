// Earlier generations of the attention kernels (first design: one CTA per 128-query tile with
// online softmax; second designs: small windowed CTA, two lock-step softmax warpgroups for global).
// ABLATIONS ONLY: compiled with `make ABLATIONS=1` (-DSAMQ_ABLATIONS) and selected with
// SAMQ_ATTN_WIN=v1|v2 / SAMQ_ATTN_GLOB=v1|v2; the shipped libsamq.so does not contain them.
#ifdef SAMQ_ABLATIONS
#include "attention_common.cuh"

namespace samq {
namespace {

constexpr int kAttThreads = 256;

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

int attn_ablation_dispatch(int generation, bool glob, int hd, const void* qkv, const void* rph, const void* rpw,
                           void* out, int B, int heads, float scale, int relw_mode, cudaStream_t st) {
  if (generation == 2 && glob)
    return hd == 64 ? launch_attn_glob<64>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st)
                    : launch_attn_glob<80>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st);
  if (generation == 2)
    return hd == 64 ? launch_attn_win<64>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st)
                    : launch_attn_win<80>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st);
  if (hd == 64)
    return glob ? launch_attn<64, false>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st)
                : launch_attn<64, true>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st);
  return glob ? launch_attn<80, false>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st)
              : launch_attn<80, true>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st);
}

}  // namespace samq
#endif  // SAMQ_ABLATIONS
