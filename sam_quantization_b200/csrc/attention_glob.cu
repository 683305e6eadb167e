// Global (64x64) attention with SAM's decomposed relative-position bias: the product kernel.
// See attention.cu for the math and the reference citations.
#include "attention_common.cuh"

namespace samq {
namespace {

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
  // images (and heads) from the END of the batch: the qkv GEMM wrote its output front to back, so the
  // last images are what the L2 still holds, and the proj GEMM reads this kernel's output from the
  // front, which is then written last (same reasoning as in attention_win.cu / layernorm.cu)
  const int q_tile = blockIdx.x, head = gridDim.y - 1 - blockIdx.y, b = gridDim.z - 1 - blockIdx.z;
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
                                                  reinterpret_cast<__half*>(out), heads, scale, relw_mode, config().attn_exact_max);
  count_launch();
  return check_launch("attn_glob3_kernel");
}

}  // namespace

int attn_glob3_dispatch(int hd, const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads,
                        float scale, int relw_mode, cudaStream_t st) {
  return hd == 64 ? launch_attn_glob3<64>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st)
                  : launch_attn_glob3<80>(qkv, rph, rpw, out, B, heads, scale, relw_mode, st);
}

}  // namespace samq
