// Windowed (14x14) attention with SAM's decomposed relative-position bias: the product kernel.
// See attention.cu for the math and the reference citations.
#include "attention_common.cuh"

namespace samq {
namespace {

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
        const int ri = n_items - 1 - item;   // items from the end: see the note at the softmax warps
        const int b = ri / heads, head = ri % heads;
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
      // (window, head) items are taken from the END of the qkv tensor: the qkv GEMM wrote it front to
      // back, so its tail is what the L2 still holds; and the proj GEMM reads this kernel's output
      // from the front, which is then written last
      const int ri = n_items - 1 - item;
      const int b = ri / heads, head = ri % heads;
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
          for (int v = vh_lo; v <= vh_hi; v += 2) {
            // two image rows per wait (the second one clamped: a duplicate load when the count is odd)
            const int v2 = min(v + 1, vh_hi);
            uint32_t rh[16], rv[16], rh2[16], rv2[16];
            tmem_ld_x16(region + v, rh);
            tmem_ld_x16(region + 32 + v, rv);
            tmem_ld_x16(region + v2, rh2);
            tmem_ld_x16(region + 32 + v2, rv2);
            tmem_ld_wait();
            if (mh == v) {
#pragma unroll
              for (int k = 0; k < E; ++k) {
                bh[k] = __uint_as_float(rh[13 - k]);
                bw[k] = __uint_as_float(rv[13 - k]);
              }
            }
            if (mh == v2) {
#pragma unroll
              for (int k = 0; k < E; ++k) {
                bh[k] = __uint_as_float(rh2[13 - k]);
                bw[k] = __uint_as_float(rv2[13 - k]);
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
        // all of O in flight at once, ONE wait (tcgen05.wait::ld waits for every outstanding load, so
        // a load / wait pair per chunk paid the TMEM latency three times)
        uint32_t r0[32], r1[32], r2[16];
        tmem_ld_x32(o_tmem, r0);
        tmem_ld_x32(o_tmem + 32, r1);
        if (C::kTail) tmem_ld_x16(o_tmem + 64, r2);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            o_row[v] = pack8(r0 + 8 * v);
            o_row[4 + v] = pack8(r1 + 8 * v);
          }
          if (C::kTail) {
            o_row[8] = pack8(r2);
            o_row[9] = pack8(r2 + 8);
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
                                                 heads, n_items, scale, relw_mode, img_nh, img_nw, config().attn_exact_max);
  count_launch();
  return check_launch("attn_win3_kernel");
}

}  // namespace

int attn_win3_dispatch(int hd, const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads,
                       float scale, int relw_mode, int img_h, int img_w, cudaStream_t st) {
  return hd == 64 ? launch_attn_win3<64>(qkv, rph, rpw, out, B, heads, scale, relw_mode, img_h, img_w, st)
                  : launch_attn_win3<80>(qkv, rph, rpw, out, B, heads, scale, relw_mode, img_h, img_w, st);
}

}  // namespace samq
