// Dense fp16 GEMM on a CTA pair (cta_group::2): y = epi(x . Wt^T + bias) + residual.
//
// This is the second stage of the "unpack once, then GEMM" path that samq_qlinear_fwd takes for
// long M (see qlinear.cu), and the kernel behind samq_dense_linear_fwd when N % 256 == 0.
// Same transposed orientation as the other GEMM kernels (features are the UMMA M dimension):
//   pair tile = 256 features x 256 tokens x 64 k per stage
//   each CTA : A = its 128 features of Wt (TMA, 128B swizzle)  +  HALF of the x tile (128 tokens)
//   leader   : one thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256), SS operands
//   TMEM     : 2 x 256 fp32 columns per CTA (double-buffered accumulator)
// Per CTA and k-block that is 32 KB written by TMA and 32 KB read by the tensor core in 512
// tensor clocks = 128 B/clk of shared-memory traffic; the single-CTA 128x192 tile needs
// 208 B/clk against the ~128 B/clk an SM provides, which capped it at ~65 % of the tensor peak.
#include "qlinear_common.cuh"

#include <cstdlib>

namespace samq {
namespace {

// Warps 0 .. kEpiWarps-1 drain the accumulator (groups of four = TMEM lane quadrants; group g owns
// token columns [256 g / G, 256 (g+1) / G)), then one TMA warp and one MMA + TMEM-alloc warp.
// (16 epilogue warps for the GELU kernel were tried: each warp's share halves but the kernel
// time does not move -- 316 us for 32768 x 1280 x 5120 either way, 283 us without GELU -- so
// the GELU cost is not an issue-slot or latency problem of the epilogue warps; 8 it stays.)
template <bool GELU>
struct DCfg {
  static constexpr int kEpiWarps = 8;
  static constexpr int kWarpTma = kEpiWarps, kWarpMma = kEpiWarps + 1;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;
  static constexpr int kEpiBytes = kEpiWarps * 2048;
};
constexpr int kDBM = 256;                 // tokens per pair tile (UMMA N)
constexpr int kDBN = 128;                 // features per CTA   (UMMA M = 256 over the pair)
constexpr int kDBK = 64;
constexpr int kDStages = 6;
constexpr int kDABytes = kDBN * kDBK * 2;          // 16 KB
constexpr int kDXBytes = (kDBM / 2) * kDBK * 2;    // 16 KB (this CTA's half of the x tile)
constexpr int kDStageBytes = kDABytes + kDXBytes;
constexpr int kDSmemData = kDStages * kDStageBytes;
template <bool GELU>
constexpr int kDSmemBytes = kDSmemData + DCfg<GELU>::kEpiBytes + (2 * kDStages + 4) * 8 + 16 + 1024;
static_assert(kDSmemBytes<true> <= 232448, "shared memory budget");

#ifdef SAMQ_GEMM_PROFILE
// developer-only clock64 breakdown (tests/micro/gemm_prof.cu); never compiled into libsamq.so
__device__ long long g_gemm_prof[2][18][8];
#define GP_DECL long long gt0 = 0, gacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define GP_BEGIN gt0 = clock64()
#define GP_END(i) gacc[i] += clock64() - gt0
#define GP_MAX(i) { long long d_ = clock64() - gt0; gacc[i] = d_ > gacc[i] ? d_ : gacc[i]; }
#define GP_FLUSH                                                                     \
  if (lane == 0 && (blockIdx.x >> 1) == 5)                                           \
    for (int i_ = 0; i_ < 8; ++i_) g_gemm_prof[blockIdx.x & 1][warp][i_] = gacc[i_]
#else
#define GP_DECL
#define GP_BEGIN
#define GP_END(i)
#define GP_MAX(i)
#define GP_FLUSH
#endif

template <bool GELU, bool RES>
__global__ void __launch_bounds__(DCfg<GELU>::kThreads, 1)
dense2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const __half* __restrict__ bias, const __half* residual, __half* y, int M, int N, int K,
              const RowMap rowmap) {
  GP_DECL;
  pdl_trigger();   // the next kernel may be scheduled; it waits for this grid's completion itself
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sepi = smem + kDSmemData;
  using DC = DCfg<GELU>;
  constexpr int kDWarpTma = DC::kWarpTma, kDWarpMma = DC::kWarpMma;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDSmemData + DC::kEpiBytes);
  uint64_t* full = bars;                       // leader's is the live one
  uint64_t* empty = full + kDStages;           // multicast commit -> both CTAs
  uint64_t* acc_full = empty + kDStages;       // multicast commit -> both CTAs
  uint64_t* acc_empty = acc_full + 2;          // leader's: 4 epilogue warps of each CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_kb = K / kDBK;
  const int NT = N / (2 * kDBN);
  const int MT = (M + kDBM - 1) / kDBM;
  const int num_tiles = NT * MT;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int my_tiles = (num_tiles - pair + npairs - 1) / npairs;
  const int total_kb = my_tiles * num_kb;

  if (warp == kDWarpMma && lane == 0) {
    for (int i = 0; i < kDStages; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 2 * DC::kEpiWarps);   // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == kDWarpTma && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  cluster_sync_all();
  if (warp == kDWarpMma) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  // launched with programmatic stream serialization: everything above overlapped the tail of the
  // preceding kernel (the weight unpack); its results are visible after this wait
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lead_full = mapa_u32(smem_u32(full), 0);
  const uint32_t lead_acc_empty = mapa_u32(smem_u32(acc_empty), 0);

  if (warp == kDWarpTma) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int n_tile = t % NT, m_tile = t / NT;
        const int n0 = n_tile * 2 * kDBN + static_cast<int>(rank) * kDBN;
        const int m0 = m_tile * kDBM + static_cast<int>(rank) * (kDBM / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(&empty[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[s], 2 * kDStageBytes);
          else mbar_arrive_cluster(lead_full + s * 8);
          uint8_t* stage = smem + s * kDStageBytes;
          tma_load_2d_pair(stage, &map_w, &full[s], kb * kDBK, n0);
          tma_load_2d_pair(stage + kDABytes, &map_x, &full[s], kb * kDBK, m0);
          if (++s == kDStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == kDWarpMma) {
    // ===================== MMA issuer (leader only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(2 * kDBN, kDBM, 0);
      int s = 0, kb = 0, lt = 0;
      uint32_t ph = 0;
      bool rdy = total_kb > 0 && mbar_test(&full[0], 0);
      for (int kbc = 0; kbc < total_kb; ++kbc) {
        const int ab = lt & 1;
        GP_BEGIN;
        if (kb == 0) mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);
        GP_END(0);
        GP_BEGIN;
        if (!rdy) mbar_wait(&full[s], ph);
        GP_END(1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kDBM;
        const uint64_t a_desc = make_smem_desc(smem_u32(smem + s * kDStageBytes), 0, 1024, kLayoutSw128);
        const uint64_t b_desc = make_smem_desc(smem_u32(smem + s * kDStageBytes + kDABytes), 0, 1024, kLayoutSw128);
        int s_n = s + 1;
        uint32_t ph_n = ph;
        if (s_n == kDStages) { s_n = 0; ph_n ^= 1; }
        const bool rdy_n = (kbc + 1 < total_kb) && mbar_test(&full[s_n], ph_n);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kDBK / 16; ++k)
            tc_mma_ss_pair(d_tmem, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc, (kb | k) != 0);
          tc_commit_pair(&empty[s], 3);
          if (kb == num_kb - 1) tc_commit_pair(&acc_full[ab], 3);
        }
        __syncwarp();
        s = s_n; ph = ph_n; rdy = rdy_n;
        if (++kb == num_kb) { kb = 0; ++lt; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 features) =====================
    // two groups of four warps (TMEM lane quadrant = warp % 4); group 0 drains token columns
    // [0, 128), group 1 drains [128, 256) of the accumulator
    const int e = warp & 3;
    const int grp = warp >> 2;
    int lt = 0;
    __half* stage = reinterpret_cast<__half*>(sepi + warp * 2048);
    constexpr int kChunks = kDBM / 32 / (DC::kEpiWarps / 4);   // 32-token chunks per group and tile
    const int c0 = grp * kChunks;
    auto feature_base = [&](int t) { return (t % NT) * 2 * kDBN + static_cast<int>(rank) * kDBN + e * 32; };
    // The residual rows of a chunk are requested one chunk ahead (across tile boundaries too):
    // with the request issued right before the chunk's own TMEM load, the ~1.5k clk of DRAM
    // latency was exposed in every chunk and the proj GEMM was epilogue-bound (12.0k clk of
    // epilogue per tile against a 10.2k clk main loop; 5.9k without the residual).
    EpiBlock<GELU, RES ? 1 : 0> blk_a, blk_b;
    if (pair < num_tiles) blk_a.prefetch((pair / NT) * kDBM + c0 * 32, M, N, feature_base(pair), lane, residual, rowmap);
    for (int t = pair; t < num_tiles; t += npairs, ++lt) {
      const int m_tile = t / NT;
      const int nb = feature_base(t);
      const int ab = lt & 1;
      const float bv = bias ? __half2float(bias[nb + lane]) : 0.f;
      const int t_next = t + npairs;
      GP_BEGIN;
      mbar_wait(&acc_full[ab], (lt >> 1) & 1);
      GP_END(0);
      GP_BEGIN;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * kDBM + (static_cast<uint32_t>(e * 32) << 16);
      // one chunk: request the residual of the chunk after it, drain this one
      auto chunk = [&](int c, EpiBlock<GELU, RES ? 1 : 0>& cur, EpiBlock<GELU, RES ? 1 : 0>& nxt, bool last) {
        // (one call site: two would merge through register moves that wait for the loads)
        const int m_next = last ? (t_next / NT) * kDBM + c0 * 32 : m_tile * kDBM + (c + 1) * 32;
        const int nb_next = last ? feature_base(t_next) : nb;
        nxt.prefetch(m_next, (last && t_next >= num_tiles) ? 0 : M, N, nb_next, lane, residual, rowmap);
        uint32_t r[32];
        tmem_ld_x32(d_tmem + c * 32, r);
        tmem_ld_wait();
        if (last) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&acc_empty[ab]);
            else mbar_arrive_cluster(lead_acc_empty + ab * 8);
          }
          GP_END(1);
          GP_MAX(3);
        }
        cur.finish(r, bv, stage, N, nb, lane, residual != nullptr, y);
      };
#pragma unroll 1
      for (int cc = 0; cc < kChunks; cc += 2) {
        chunk(c0 + cc, blk_a, blk_b, false);
        chunk(c0 + cc + 1, blk_b, blk_a, cc + 2 == kChunks);
      }
      GP_END(2);
      GP_MAX(4);
    }
  }
  GP_FLUSH;

  tc_fence_before();
  cluster_sync_all();
  if (warp == kDWarpMma) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

int launch_dense_pair(const void* x, const void* wt, const __half* bias, const __half* residual, __half* y,
                      int64_t M, int K, int N, int epilogue, const RowMap& rowmap, int num_sms, cudaStream_t st) {
  const CUtensorMap* mx = get_tensor_map_2d(x, static_cast<uint64_t>(M), K, static_cast<uint64_t>(K) * 2,
                                            kDBM / 2, kDBK, 2, 3);
  const CUtensorMap* mw = get_tensor_map_2d(wt, N, K, static_cast<uint64_t>(K) * 2, kDBN, kDBK, 2, 3);
  if (!mx || !mw) return SAMQ_ERR_LAUNCH;
  const bool gelu = epilogue == SAMQ_EPI_GELU;
  const bool res = residual != nullptr;
  auto kern = gelu ? (res ? dense2_kernel<true, true> : dense2_kernel<true, false>)
                   : (res ? dense2_kernel<false, true> : dense2_kernel<false, false>);
  const int smem_bytes = gelu ? kDSmemBytes<true> : kDSmemBytes<false>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_bytes, "dense2_kernel"); rc != SAMQ_OK) return rc;
  const int NT = N / (2 * kDBN);
  const int64_t MT = (M + kDBM - 1) / kDBM;
  const int64_t tiles = NT * MT;
  const int max_pairs = num_sms / 2;
  const int pairs = static_cast<int>(tiles < max_pairs ? tiles : max_pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(gelu ? DCfg<true>::kThreads : DCfg<false>::kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *mx, *mw, bias, residual, y, static_cast<int>(M), N, K, rowmap);
  count_launch();
  if (e != cudaSuccess) {
    set_error("dense2_kernel launch: %s", cudaGetErrorString(e));
    return SAMQ_ERR_LAUNCH;
  }
  return check_launch("dense2_kernel");
}

}  // namespace samq
