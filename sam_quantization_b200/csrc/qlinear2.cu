// ABLATION: compiled only with `make ABLATIONS=1` (-DSAMQ_ABLATIONS); not in the shipped libsamq.so.
#ifdef SAMQ_ABLATIONS
// 2-CTA (cta_group::2) variant of the fused int4 dequant-GEMM (see qlinear.cu for the
// single-CTA kernel, the operand orientation and the dequant arithmetic).
//
// A CTA pair (cluster 2x1x1, two SMs of one TPC) computes a 256-feature x BM-token tile:
//   * each CTA dequantises ITS 128 features into ITS TMEM (A operand, 128 lanes),
//   * each CTA TMA-loads HALF of the x tile (BM/2 tokens) into its own shared memory,
//   * the leader CTA's single issuing thread drives both tensor cores with
//     tcgen05.mma.cta_group::2 (M = 256); the hardware feeds each SM both x halves,
//   * each CTA drains its own 128 x BM accumulator.
// Versus the 1-CTA kernel this halves, per SM: the x bytes read from L2 and written to /
// read from shared memory (shared-memory bandwidth was the structural ceiling of the
// 1-CTA tile: 159 B/clk needed vs 128 B/clk available), and the MMA-issue overhead (one
// issuer per two SMs).
//
// Cross-CTA protocol (pattern of DeepGEMM / CUTLASS 2-SM kernels):
//   x_full[s], a_full[s], acc_empty[b]  live in the LEADER; the peer signals them remotely
//       (TMA complete_tx with the peer bit cleared / mbarrier.arrive on a mapa address),
//   x_empty[s], a_empty[s], acc_full[b] are arrived in BOTH CTAs by a multicast
//       tcgen05.commit, each CTA's producers / epilogue wait on their local copy,
//   w_full / w_empty (packed-weight ring) are CTA-local.
#include "qlinear_common.cuh"

namespace samq {
namespace {

constexpr int kBN2 = 128;   // features per CTA (UMMA M = 256 over the pair)
constexpr int kBK2 = 64;
constexpr int kThreads2 = 448;
constexpr int kWarpTma2 = 12, kWarpMma2 = 13;
constexpr int kWStageBytes2 = 8 * kBN2 * 4;

template <int BM>
struct Cfg2 {
  static constexpr int kXHalfRows = BM / 2;
  static constexpr int kXStageBytes = kXHalfRows * kBK2 * 2;   // this CTA's half of the x tile
  static constexpr int kXStages = 8;
  static constexpr int kWStages = 8;
  static constexpr int kAStages = (512 - 2 * BM) / 32;
  static constexpr int kTmemABase = 2 * BM;
  static constexpr int kSmemData = kXStages * kXStageBytes + kWStages * kWStageBytes2;
  static constexpr int kEpiBytes = 4 * 2048;
  static constexpr int kNumBars = 2 * kXStages + 2 * kWStages + 2 * 8 + 4;
  static constexpr int kSmemBytes = kSmemData + kEpiBytes + kNumBars * 8 + 16 + 1024;
  static_assert(kXStageBytes % 1024 == 0, "x half tile must keep 1024-byte alignment (128B swizzle)");
  static_assert((kAStages & (kAStages - 1)) == 0 && kAStages >= 2, "TMEM A stages");
  static_assert(BM % 32 == 0 && BM <= 256 && (BM / 2) % 8 == 0, "BM");
};

template <int BM, bool GELU>
__global__ void __launch_bounds__(kThreads2, 1)
qlinear2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const __half* __restrict__ scales, const int32_t* __restrict__ qzeros,
                const __half* __restrict__ bias, const __half* residual, __half* y, int M, int N,
                int K, int groupsize, const RowMap rowmap) {
  using C = Cfg2<BM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sx = smem;
  uint8_t* sw = sx + C::kXStages * C::kXStageBytes;
  uint8_t* sepi = smem + C::kSmemData;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemData + C::kEpiBytes);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + C::kXStages;
  uint64_t* w_full = x_empty + C::kXStages;
  uint64_t* w_empty = w_full + C::kWStages;
  uint64_t* a_full = w_empty + C::kWStages;
  uint64_t* a_empty = a_full + 8;
  uint64_t* acc_full = a_empty + 8;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int num_kb = K / kBK2;
  const int NT = N / (2 * kBN2);           // 256-feature tiles
  const int MT = (M + BM - 1) / BM;
  const int num_tiles = NT * MT;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int my_tiles = (num_tiles - pair + npairs - 1) / npairs;
  const int total_kb = my_tiles * num_kb;

  if (warp == kWarpMma2 && lane == 0) {
    for (int i = 0; i < C::kXStages; ++i) {
      mbar_init(&x_full[i], 2);     // leader's expect_tx arrival + the peer's remote arrival
      mbar_init(&x_empty[i], 1);    // multicast commit
    }
    for (int i = 0; i < C::kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 4);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 8);     // 4 dequant warps of each CTA
      mbar_init(&a_empty[i], 1);    // multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);   // multicast commit
      mbar_init(&acc_empty[i], 8);  // 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == kWarpTma2 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  cluster_sync_all();   // both CTAs are resident before the paired TMEM allocation
  if (warp == kWarpMma2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();   // barriers initialised and TMEM allocated in BOTH CTAs before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // remote (leader) addresses of the barriers the peer signals
  const uint32_t lead_x_full = mapa_u32(smem_u32(x_full), 0);
  const uint32_t lead_a_full = mapa_u32(smem_u32(a_full), 0);
  const uint32_t lead_acc_empty = mapa_u32(smem_u32(acc_empty), 0);

  if (warp == kWarpTma2) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int xs = 0, ws = 0;
      uint32_t xph = 0, wph = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int n_tile = t % NT, m_tile = t / NT;
        const int n0 = n_tile * 2 * kBN2 + static_cast<int>(rank) * kBN2;
        const int m0 = m_tile * BM + static_cast<int>(rank) * C::kXHalfRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&w_full[ws], kWStageBytes2);
          tma_load_2d(sw + ws * kWStageBytes2, &map_w, &w_full[ws], n0, kb * 8);
          if (++ws == C::kWStages) { ws = 0; wph ^= 1; }

          mbar_wait(&x_empty[xs], xph ^ 1);
          if (leader) mbar_arrive_expect_tx(&x_full[xs], 2 * C::kXStageBytes);
          else mbar_arrive_cluster(lead_x_full + xs * 8);
          tma_load_2d_pair(sx + xs * C::kXStageBytes, &map_x, &x_full[xs], kb * kBK2, m0);
          if (++xs == C::kXStages) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma2) {
    // ===================== MMA issuer (leader CTA only) =====================
    // early non-blocking probes of the next k-block's barriers overlap the MMA issue
    // (see the 1-CTA kernel for the measurement behind this)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(2 * kBN2, BM, 0);
      int xs = 0, as = 0, kb = 0, lt = 0;
      uint32_t xph = 0, aph = 0;
      bool a_rdy = total_kb > 0 && mbar_test(&a_full[0], 0);
      bool x_rdy = total_kb > 0 && mbar_test(&x_full[0], 0);
      for (int kbc = 0; kbc < total_kb; ++kbc) {
        const int ab = lt & 1;
        if (kb == 0) mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);
        if (!a_rdy) mbar_wait(&a_full[as], aph);
        if (!x_rdy) mbar_wait(&x_full[xs], xph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * BM;
        const uint64_t b_desc =
            make_smem_desc(smem_u32(sx + xs * C::kXStageBytes), 0, 1024, kLayoutSw128);
        const uint32_t a_tmem = tmem_base + C::kTmemABase + as * 32;
        int xs_n = xs + 1, as_n = as + 1;
        uint32_t xph_n = xph, aph_n = aph;
        if (xs_n == C::kXStages) { xs_n = 0; xph_n ^= 1; }
        if (as_n == C::kAStages) { as_n = 0; aph_n ^= 1; }
        const bool more = kbc + 1 < total_kb;
        const bool a_rdy_n = more && mbar_test(&a_full[as_n], aph_n);
        const bool x_rdy_n = more && mbar_test(&x_full[xs_n], xph_n);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK2 / 16; ++k)
            tc_mma_ts_pair(d_tmem, a_tmem + k * 8, b_desc + (k * 32 >> 4), idesc, (kb | k) != 0);
          tc_commit_pair(&x_empty[xs], 3);
          tc_commit_pair(&a_empty[as], 3);
          if (kb == num_kb - 1) tc_commit_pair(&acc_full[ab], 3);
        }
        __syncwarp();
        xs = xs_n; xph = xph_n;
        as = as_n; aph = aph_n;
        a_rdy = a_rdy_n;
        x_rdy = x_rdy_n;
        if (++kb == num_kb) { kb = 0; ++lt; }
      }
    }
  } else if (warp < 4 || (warp >= 8 && warp < 12)) {
    // ===================== dequant warps (both CTAs, own 128 features) =====================
    dequant_warp_loop<C::kWStages, C::kAStages, kWStageBytes2>(
        warp >> 3, warp & 3, lane, total_kb, num_kb, N, groupsize, scales, qzeros, sw, w_full, w_empty,
        a_empty, tmem_base + C::kTmemABase,
        [&](int tl) { return ((pair + tl * npairs) % NT) * 2 * kBN2 + static_cast<int>(rank) * kBN2; },
        [&](int as) {
          if (leader) mbar_arrive(&a_full[as]);
          else mbar_arrive_cluster(lead_a_full + as * 8);
        });
  } else if (warp < 8) {
    // ===================== epilogue warps (both CTAs, own accumulator) =====================
    const int e = warp - 4;
    int lt = 0;
    __half* stage = reinterpret_cast<__half*>(sepi + e * 2048);
    for (int t = pair; t < num_tiles; t += npairs, ++lt) {
      const int n_tile = t % NT, m_tile = t / NT;
      const int nb = n_tile * 2 * kBN2 + static_cast<int>(rank) * kBN2 + e * 32;
      const int ab = lt & 1;
      const uint32_t acc_ph = (lt >> 1) & 1;
      const float bv = bias ? __half2float(bias[nb + lane]) : 0.f;
      mbar_wait(&acc_full[ab], acc_ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * BM + (static_cast<uint32_t>(e * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BM / 32; ++c) {
        const int m0 = m_tile * BM + c * 32;
        EpiBlock<GELU> blk;
        blk.prefetch(m0, M, N, nb, lane, residual, rowmap);
        uint32_t r[32];
        tmem_ld_x32(d_tmem + c * 32, r);
        tmem_ld_wait();
        if (c == BM / 32 - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&acc_empty[ab]);
            else mbar_arrive_cluster(lead_acc_empty + ab * 8);
          }
        }
        blk.finish(r, bv, stage, N, nb, lane, residual != nullptr, y);
      }
    }
  }

  // neither CTA may exit (or free TMEM) while its partner can still touch its shared
  // memory, tensor memory or barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == kWarpMma2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

// Launcher used by samq_qlinear_fwd when N % 256 == 0 (declared in qlinear.cu).
int launch_qlinear_pair(const void* x, const void* qweight, const __half* scales,
                        const int32_t* qzeros, const __half* bias, const __half* residual,
                        __half* y, int64_t M, int K, int N, int groupsize, int epilogue,
                        const RowMap& rowmap, int num_sms, cudaStream_t st) {
  constexpr int BM = 192;
  using C = Cfg2<BM>;
  const CUtensorMap* mx = get_tensor_map_2d(x, static_cast<uint64_t>(M), K, static_cast<uint64_t>(K) * 2,
                                            C::kXHalfRows, kBK2, 2, 3);
  const CUtensorMap* mw = get_tensor_map_2d(qweight, K / 8, N, static_cast<uint64_t>(N) * 4, 8, kBN2, 4, 0);
  if (!mx || !mw) return SAMQ_ERR_LAUNCH;
  const bool gelu = epilogue == SAMQ_EPI_GELU;
  auto kern = gelu ? qlinear2_kernel<BM, true> : qlinear2_kernel<BM, false>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "qlinear2_kernel"); rc != SAMQ_OK) return rc;
  const int NT = N / (2 * kBN2);
  const int64_t MT = (M + BM - 1) / BM;
  const int64_t tiles = NT * MT;
  const int max_pairs = num_sms / 2;
  const int pairs = static_cast<int>(tiles < max_pairs ? tiles : max_pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *mx, *mw, scales, qzeros, bias, residual, y,
                                     static_cast<int>(M), N, K, groupsize, rowmap);
  count_launch();
  if (e != cudaSuccess) {
    set_error("qlinear2_kernel launch: %s", cudaGetErrorString(e));
    return SAMQ_ERR_LAUNCH;
  }
  return check_launch("qlinear2_kernel");
}

}  // namespace samq

#endif  // SAMQ_ABLATIONS
