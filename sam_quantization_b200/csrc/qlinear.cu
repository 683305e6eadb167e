// Dequant-GEMM for GPTQ QuantLinear on sm_100a:  y = epi(x . W + bias) + residual.
//
// Replaces triton_matmul4 / matmul4_kernel (gptq_triton/quant_linear.py:231-437).
//
// Orientation.  The kernel computes the TRANSPOSED product D[n, m] = sum_k Wt[n,k] x[m,k]
// so that the (dequantised) weight is the tcgen05 A operand and can live in TMEM:
//   A = 128 output features x 64 k   (fused mode: written to TMEM by the dequant
//                                     warps with tcgen05.st; dense mode: fp16 Wt
//                                     tile in shared memory, TMA-staged)
//   B = BM tokens x 64 k             (x tile, K-major, 128-byte swizzle, TMA-staged)
//   D = 128 lanes (n) x BM fp32 columns (m) in TMEM, two buffers so the epilogue of
//       tile i overlaps the main loop of tile i+1.
// Each dequant thread owns ONE output feature n (= its TMEM lane), so scale and
// zero are thread-constant inside a quantisation group and the packed words
// qweight[k/8, n] are read coalesced (128 consecutive n per row).
//
// Warp roles (448 threads, persistent CTA, one per SM):
//   warps 0-3   dequant set 0 (even k-blocks): packed int4 (smem) -> fp16 -> TMEM A stage
//   warps 4-7   epilogue: TMEM D -> +bias -> GELU -> fp16 -> smem transpose -> +residual -> global
//   warps 8-11  dequant set 1 (odd k-blocks)                                   [fused mode]
//   warp 12     TMA producer: packed-weight tiles + x tiles (+ Wt tiles in dense mode)
//   warp 13     TMEM allocator; one thread issues tcgen05.mma kind::f16 and commits -> mbarriers
//
// Dequant arithmetic is bit-identical to dequant.cu / oracle "fma" form, which is what the
// reference's Triton kernel computes on this GPU (tests/golden/dequant_triton_b4.npz):
//   q  = (1024 + q) - 1024                     (exact)
//   w  = fma(q, s, -fp16((z+1)*s))             (one fp16 rounding)
#include "qlinear_common.cuh"

#include <cstdlib>
#include <cstring>

namespace samq {

int launch_qlinear_pair(const void* x, const void* qweight, const __half* scales,
                        const int32_t* qzeros, const __half* bias, const __half* residual,
                        __half* y, int64_t M, int K, int N, int groupsize, int epilogue,
                        const RowMap& rowmap, int num_sms, cudaStream_t st);

int launch_dense_pair(const void* x, const void* wt, const __half* bias, const __half* residual, __half* y,
                      int64_t M, int K, int N, int epilogue, const RowMap& rowmap, int num_sms, cudaStream_t st);

int prefetch_dequant4(const int32_t* qweight, const int32_t* qzeros, const void* scales, void* w_out, int K, int N,
                      int groupsize, int num_sms, cudaStream_t st);
int unpack_dequant(const int32_t* qweight, const int32_t* qzeros, const void* scales,
                   const int32_t* g_idx, void* w_out, int K, int N, int bits, int groupsize,
                   int transposed, cudaStream_t st);

namespace {

constexpr int kBN = 128;        // output features per tile (UMMA M)
constexpr int kBK = 64;         // k per pipeline stage (one 128-byte swizzle row of fp16)
constexpr int kThreads = 448;   // 14 warps, see the role table above
constexpr int kWarpTma = 12, kWarpMma = 13;
constexpr int kAStageBytes = kBN * kBK * 2; // dense fp16 Wt tile

template <int BM, bool FUSED, int BITS = 4>
struct Cfg {
  static constexpr int kWStageBytes = 2 * BITS * kBN * 4;   // packed tile of one k-block: 2*BITS words x 128 n
  static constexpr int kXStageBytes = BM * kBK * 2;
  static constexpr int kXStages = FUSED ? 6 : 5;
  static constexpr int kWStages = BITS == 8 ? 4 : 8;        // fused only (8-bit tiles are 8 KB each)
  static constexpr int kAStages = (512 - 2 * BM) / 32;      // TMEM A stages (fused only)
  static constexpr int kTmemABase = 2 * BM;                 // column offset of A stages
  static constexpr int kSmemData =
      kXStages * kXStageBytes + (FUSED ? kWStages * kWStageBytes : kXStages * kAStageBytes);
  static_assert(kSmemData + 4 * 2048 + 1024 + 512 <= 232448, "shared memory budget");
  static constexpr int kEpiBytes = 4 * 2048;  // one 32x32 fp16 transpose block per epilogue warp
  static constexpr int kNumBars = 2 * kXStages + 2 * kWStages + 2 * 8 + 4;
  static constexpr int kSmemBytes = kSmemData + kEpiBytes + kNumBars * 8 + 16 + 1024;  // + alignment slack
  static_assert(kAStages >= 2 || !FUSED, "need at least two TMEM A stages");
  static_assert(BM % 32 == 0 && BM <= 256, "BM");
};

// F32ACC (dense mode only): the epilogue writes  y32[m, n] = alpha * acc + beta * y32[m, n]  in fp32
// instead of the fp16 bias / GELU / residual epilogue -- the Hessian accumulation of the GPTQ solver
// (samq_syrk_f32_fwd below); bias / residual / rowmap are ignored.
template <int BM, bool FUSED, bool GELU, int BITS = 4, bool F32ACC = false>
__global__ void __launch_bounds__(kThreads, 1)
qlinear_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const __half* __restrict__ scales, const int32_t* __restrict__ qzeros,
               const __half* __restrict__ bias, const __half* residual, __half* y, int M, int N,
               int K, int groupsize, const RowMap rowmap, float alpha, float beta) {
  using C = Cfg<BM, FUSED, BITS>;
  constexpr int kWStageBytes = C::kWStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sx = smem;
  uint8_t* sw = sx + C::kXStages * C::kXStageBytes;  // fused: packed W ring; dense: Wt ring
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

  const int num_kb = (K + kBK - 1) / kBK;   // a ragged last k-block is zero-filled by TMA (dense mode)
  const int NT = (N + kBN - 1) / kBN;       // ragged only for the fp32 Hessian kernel (rows beyond N: TMA zero fill)
  const int MT = (M + BM - 1) / BM;
  const int num_tiles = NT * MT;
  // k-blocks this CTA walks through, over all of its tiles (tile i of this CTA is
  // blockIdx.x + i * gridDim.x)
  const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int total_kb = my_tiles * num_kb;

  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < C::kXStages; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < C::kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 4);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kWarpTma) {
    // ===================== TMA producer: x tiles + weight tiles =====================
    if (lane == 0) {
      int xs = 0, ws = 0;
      uint32_t xph = 0, wph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % NT, m_tile = t / NT;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (FUSED) {
            // packed weights first: the dequant warps are the longer leg of the pipeline
            mbar_wait(&w_empty[ws], wph ^ 1);
            mbar_arrive_expect_tx(&w_full[ws], kWStageBytes);
            tma_load_2d(sw + ws * kWStageBytes, &map_w, &w_full[ws], n_tile * kBN, kb * 2 * BITS);
            if (++ws == C::kWStages) { ws = 0; wph ^= 1; }
          }
          mbar_wait(&x_empty[xs], xph ^ 1);
          mbar_arrive_expect_tx(&x_full[xs], C::kXStageBytes + (FUSED ? 0 : kAStageBytes));
          tma_load_2d(sx + xs * C::kXStageBytes, &map_x, &x_full[xs], kb * kBK, m_tile * BM);
          if (!FUSED)
            tma_load_2d(sw + xs * kAStageBytes, &map_w, &x_full[xs], kb * kBK, n_tile * kBN);
          if (++xs == C::kXStages) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop convergently (so stage indices, descriptors and TMEM
    // addresses are warp-uniform and live in uniform registers); one elected lane issues.
    // A blocking mbarrier wait costs ~160 clk even when the phase is already complete
    // (tests/micro/umma_rate.cu), and two of them plus the ~155 clk of MMA issue exceed the
    // 384 clk the tensor core needs per k-block.  So the barriers of k-block i+1 are PROBED
    // (non-blocking test_wait) before the MMAs of k-block i are issued; the probe latency
    // overlaps the issue, and the blocking wait only runs when the data really is late.
    {
      constexpr uint32_t idesc = make_idesc_f16(kBN, BM, 0);
      int xs = 0, as = 0, kb = 0, lt = 0;
      uint32_t xph = 0, aph = 0;
      bool a_rdy = FUSED ? (total_kb > 0 && mbar_test(&a_full[0], 0)) : true;
      bool x_rdy = total_kb > 0 && mbar_test(&x_full[0], 0);
      for (int kbc = 0; kbc < total_kb; ++kbc) {
        const int ab = lt & 1;
        if (kb == 0) {   // new tile: the epilogue must have drained this accumulator buffer
          mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);
        }
        if (FUSED && !a_rdy) mbar_wait(&a_full[as], aph);
        if (!x_rdy) mbar_wait(&x_full[xs], xph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * BM;
        const uint64_t b_desc =
            make_smem_desc(smem_u32(sx + xs * C::kXStageBytes), 0, 1024, kLayoutSw128);
        const uint32_t a_tmem = tmem_base + C::kTmemABase + as * 32;
        const uint64_t a_desc =
            make_smem_desc(smem_u32(sw + xs * kAStageBytes), 0, 1024, kLayoutSw128);
        // next k-block's stage indices + early probes
        int xs_n = xs + 1, as_n = as + 1;
        uint32_t xph_n = xph, aph_n = aph;
        if (xs_n == C::kXStages) { xs_n = 0; xph_n ^= 1; }
        if (as_n == C::kAStages) { as_n = 0; aph_n ^= 1; }
        const bool more = kbc + 1 < total_kb;
        const bool a_rdy_n = FUSED ? (more && mbar_test(&a_full[as_n], aph_n)) : true;
        const bool x_rdy_n = more && mbar_test(&x_full[xs_n], xph_n);
        if (elect_one()) {
          if (FUSED) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma_ts(d_tmem, a_tmem + k * 8, b_desc + (k * 32 >> 4), idesc, (kb | k) != 0);
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma_ss(d_tmem, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc, (kb | k) != 0);
          }
          tc_commit(&x_empty[xs]);
          if (FUSED) tc_commit(&a_empty[as]);
          if (kb == num_kb - 1) tc_commit(&acc_full[ab]);
        }
        __syncwarp();
        xs = xs_n; xph = xph_n;
        if (FUSED) { as = as_n; aph = aph_n; }
        a_rdy = a_rdy_n;
        x_rdy = x_rdy_n;
        if (++kb == num_kb) { kb = 0; ++lt; }
      }
    }
  } else if (warp < 4 || (warp >= 8 && warp < 12)) {
    // ===================== dequant warps (fused mode) =====================
    // Two sets of four warps (set 0 = warps 0-3, set 1 = warps 8-11; warp % 4 = TMEM lane
    // quadrant) take alternate k-blocks, so one set's unpack + tcgen05.st latency (~500 clk
    // per k-block for a single warp) hides behind the other's and behind the 384-clk MMA.
    if (FUSED) {
      dequant_warp_loop<C::kWStages, C::kAStages, kWStageBytes, BITS>(
          warp >> 3, warp & 3, lane, total_kb, num_kb, N, groupsize, scales, qzeros, sw, w_full, w_empty,
          a_empty, tmem_base + C::kTmemABase,
          [&](int tl) { return ((static_cast<int>(blockIdx.x) + tl * static_cast<int>(gridDim.x)) % NT) * kBN; },
          [&](int as) { mbar_arrive(&a_full[as]); });
    }
  } else if (warp < 8) {
    // ===================== epilogue warps =====================
    const int e = warp - 4;
    int lt = 0;
    // 32(m) x 32(n) staging block of this warp: lanes own n, so the block is transposed
    // through shared memory and written as 16-byte row segments (2-byte scattered global
    // stores ran at <2 B/clk/SM).
    __half* stage = reinterpret_cast<__half*>(sepi + e * 2048);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int n_tile = t % NT, m_tile = t / NT;
      const int n = n_tile * kBN + e * 32 + lane;
      const int ab = lt & 1;
      const uint32_t acc_ph = (lt >> 1) & 1;
      const float bv = bias ? __half2float(bias[n]) : 0.f;
      mbar_wait(&acc_full[ab], acc_ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * BM + (static_cast<uint32_t>(e * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BM / 32; ++c) {
        const int m0 = m_tile * BM + c * 32;
        if constexpr (F32ACC) {
          // lanes own n (consecutive), registers own m: every register is one coalesced 128-byte row segment
          uint32_t r[32];
          tmem_ld_x32(d_tmem + c * 32, r);
          tmem_ld_wait();
          if (c == BM / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
          }
          float* y32 = reinterpret_cast<float*>(y);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = m0 + j;
            if (m < M && n < N) {
              float* dst = y32 + static_cast<size_t>(m) * N + n;
              const float v = alpha * __uint_as_float(r[j]);
              *dst = beta != 0.f ? fmaf(beta, *dst, v) : v;
            }
          }
          continue;
        }
        EpiBlock<GELU> blk;
        blk.prefetch(m0, M, N, n_tile * kBN + e * 32, lane, residual, rowmap);
        uint32_t r[32];
        tmem_ld_x32(d_tmem + c * 32, r);
        tmem_ld_wait();
        if (c == BM / 32 - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        blk.finish(r, bv, stage, N, n_tile * kBN + e * 32, lane, residual != nullptr, y);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int num_sms() { return device_sm_count(); }

template <int BM, bool FUSED, int BITS = 4>
int launch_qlinear(const void* x, const void* w, const __half* scales, const int32_t* qzeros,
                   const __half* bias, const __half* residual, __half* y, int64_t M, int K, int N,
                   int groupsize, int epilogue, const RowMap& rowmap, cudaStream_t st) {
  using C = Cfg<BM, FUSED, BITS>;
  const CUtensorMap* mx = get_tensor_map_2d(x, static_cast<uint64_t>(M), K, static_cast<uint64_t>(K) * 2, BM, kBK, 2, 3);
  if (!mx) return SAMQ_ERR_LAUNCH;
  const CUtensorMap* mw =
      FUSED ? get_tensor_map_2d(w, static_cast<uint64_t>(K) * BITS / 32, N, static_cast<uint64_t>(N) * 4, 2 * BITS, kBN, 4, 0)
            : get_tensor_map_2d(w, N, K, static_cast<uint64_t>(K) * 2, kBN, kBK, 2, 3);
  if (!mw) return SAMQ_ERR_LAUNCH;
  auto kern = epilogue == SAMQ_EPI_GELU ? qlinear_kernel<BM, FUSED, true, BITS> : qlinear_kernel<BM, FUSED, false, BITS>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "qlinear_kernel"); rc != SAMQ_OK) return rc;
  const int NT = N / kBN;
  const int64_t MT = (M + BM - 1) / BM;
  const int64_t tiles = NT * MT;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(*mx, *mw, scales, qzeros, bias, residual, y,
                                              static_cast<int>(M), N, K, groupsize, rowmap, 1.f, 0.f);
  count_launch();
  return check_launch("qlinear_kernel");
}

// dense fp16 GEMM: the CTA-pair 256x256 kernel when the feature count tiles by 256 and there is
// enough work for it, else the single-CTA 128x192 kernel (SAMQ_DENSE=1cta forces the latter)
int launch_dense(const void* x, const void* wt, const __half* bias, const __half* residual, __half* y,
                 int64_t M, int K, int N, int epilogue, const RowMap& rowmap, cudaStream_t st) {
  const bool force_1cta = config().dense_1cta;
  if (N % 256 == 0 && M >= 2048 && !force_1cta)
    return launch_dense_pair(x, wt, bias, residual, y, M, K, N, epilogue, rowmap, num_sms(), st);
  return launch_qlinear<192, false>(x, wt, nullptr, nullptr, bias, residual, y, M, K, N, K, epilogue, rowmap, st);
}

int check_common(const void* x, const void* y, int64_t M, int K, int N, int epilogue,
                 const char* who) {
  SAMQ_REQUIRE(x && y, SAMQ_ERR_BAD_ARG, "%s: null pointer", who);
  SAMQ_REQUIRE(M > 0 && M < (1ll << 31) / 2, SAMQ_ERR_BAD_SHAPE, "%s: M=%lld out of range", who, (long long)M);
  SAMQ_REQUIRE(K > 0 && K % kBK == 0, SAMQ_ERR_BAD_SHAPE, "%s: K=%d must be a positive multiple of %d", who, K, kBK);
  SAMQ_REQUIRE(N > 0 && N % kBN == 0, SAMQ_ERR_BAD_SHAPE, "%s: N=%d must be a positive multiple of %d", who, N, kBN);
  SAMQ_REQUIRE(epilogue == SAMQ_EPI_NONE || epilogue == SAMQ_EPI_GELU, SAMQ_ERR_BAD_ARG, "%s: bad epilogue %d", who, epilogue);
  SAMQ_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, SAMQ_ERR_BAD_ARG, "%s: x must be 16-byte aligned", who);
  return SAMQ_OK;
}

}  // namespace

}  // namespace samq

namespace samq {
namespace {

// common implementation of samq_qlinear_fwd / samq_qlinear_unpartition_fwd
int qlinear_impl(const char* who, const void* x, const int32_t* qweight, const int32_t* qzeros,
                 const void* scales, const int32_t* g_idx, const void* bias, const void* residual, void* y,
                 void* workspace, int64_t M, int K, int N, int bits, int groupsize, int epilogue,
                 const RowMap& rowmap, cudaStream_t st) {
  SAMQ_REQUIRE(bits == 2 || bits == 3 || bits == 4 || bits == 8, SAMQ_ERR_UNSUPPORTED_BITS,
               "%s: bits must be 2, 3, 4 or 8 (got %d)", who, bits);
  int rc = check_common(x, y, M, K, N, epilogue, who);
  if (rc != SAMQ_OK) return rc;
  if (!qweight && !qzeros && !scales && !g_idx) {
    // the weight was unpacked ahead of time (samq_qlinear_prefetch): `workspace` holds fp16 Wt[N, K]
    SAMQ_REQUIRE(workspace && reinterpret_cast<uintptr_t>(workspace) % 16 == 0, SAMQ_ERR_BAD_ARG,
                 "%s: no packed weight and no 16-byte aligned prefetched workspace", who);
    return launch_dense(x, workspace, reinterpret_cast<const __half*>(bias), reinterpret_cast<const __half*>(residual),
                        reinterpret_cast<__half*>(y), M, K, N, epilogue, rowmap, st);
  }
  SAMQ_REQUIRE(qweight && qzeros && scales, SAMQ_ERR_BAD_ARG, "%s: null weight pointer", who);
  if (groupsize == -1) groupsize = K;
  SAMQ_REQUIRE(groupsize > 0 && K % groupsize == 0, SAMQ_ERR_BAD_SHAPE,
               "%s: groupsize=%d must divide K=%d", who, groupsize, K);
  const __half* b = reinterpret_cast<const __half*>(bias);
  const __half* r = reinterpret_cast<const __half*>(residual);
  __half* out = reinterpret_cast<__half*>(y);
  // The fused kernel (unpack in registers -> TMEM A operand -> MMA) exists for every packing
  // (template parameter BITS of qlinear_kernel / unpack_kblock); it needs contiguous groups of a
  // multiple of 64 k.  Act-order checkpoints (g_idx) reach it through the host, which sorts the
  // rows of qweight by group once and gathers the columns of x (ops.py / QuantLinear.sorted_pack).
  const bool fused = g_idx == nullptr && groupsize % kBK == 0 && reinterpret_cast<uintptr_t>(qweight) % 16 == 0 &&
                     (bits != 3 || (K % 32 == 0 && N % 32 == 0)) && N % (bits == 3 ? 32 : 32 / bits) == 0;
  if (fused) {
    // Kernel choice (all sm_100a, all tcgen05):
    //   fused 1-CTA  (default for M < kTwoKernelMinM): unpack in registers -> TMEM -> MMA.  Weight
    //                bytes stay packed all the way, which is what matters while the GEMM is short.
    //   unpack-once + dense GEMM (default for M >= kTwoKernelMinM when a workspace is given):
    //                every 128x192 tile of the fused kernel re-dequantises its weight tile, i.e.
    //                M/192 times per weight; for long M that redundant ALU work (and its power)
    //                costs more than reading fp16 weights from L2, and the dense GEMM can use the
    //                256x256 CTA-pair tile.  Measured sustained at M = 32768 (tests/gemm_power.py):
    //                fused 870, unpack + pair-dense 1199 TFLOP/s (cuBLAS fp16: 1193).
    //   fused 2-CTA  (SAMQ_GEMM=2cta, `make ABLATIONS=1` builds only, int4): cta_group::2 pair kernel, qlinear2.cu.
    // SAMQ_GEMM = fused | dense forces one of the product paths (tests: both must be bit-identical);
    // resolved once at load (Config, common.cuh), not per call.
    // Threshold measured on the B200 (round 2): the SAM encoder at batch 1 (every GEMM M = 4096) runs
    // 107.7 images/s on the fused kernel and 135.3 on unpack-once + pair GEMM; batch 2: 128.6 vs
    // 157.8; the fused kernel wins where the pair kernel's 256x256 tiles cannot fill the GPU
    // (M = 196: 21 vs 33 us per call).
    constexpr int64_t kTwoKernelMinM = 2048;
    const int variant = config().gemm;
    const bool force_fused = variant == 1, force_dense = variant == 2;
#ifdef SAMQ_ABLATIONS
    if (variant == 3 && N % 256 == 0 && bits == 4)
      return launch_qlinear_pair(x, qweight, reinterpret_cast<const __half*>(scales), qzeros, b, r, out, M, K, N,
                                 groupsize, epilogue, rowmap, num_sms(), st);
#endif
    if (workspace && reinterpret_cast<uintptr_t>(workspace) % 16 == 0 && !force_fused &&
        (force_dense || M >= kTwoKernelMinM)) {
      rc = unpack_dequant(qweight, qzeros, scales, nullptr, workspace, K, N, bits, groupsize, 1, st);
      if (rc != SAMQ_OK) return rc;
      return launch_dense(x, workspace, b, r, out, M, K, N, epilogue, rowmap, st);
    }
    const __half* sc = reinterpret_cast<const __half*>(scales);
    switch (bits) {
      case 2: return launch_qlinear<192, true, 2>(x, qweight, sc, qzeros, b, r, out, M, K, N, groupsize, epilogue, rowmap, st);
      case 3: return launch_qlinear<192, true, 3>(x, qweight, sc, qzeros, b, r, out, M, K, N, groupsize, epilogue, rowmap, st);
      case 8: return launch_qlinear<192, true, 8>(x, qweight, sc, qzeros, b, r, out, M, K, N, groupsize, epilogue, rowmap, st);
      default: return launch_qlinear<192, true, 4>(x, qweight, sc, qzeros, b, r, out, M, K, N, groupsize, epilogue, rowmap, st);
    }
  }
  SAMQ_REQUIRE(workspace && reinterpret_cast<uintptr_t>(workspace) % 16 == 0, SAMQ_ERR_BAD_ARG,
               "%s: bits=%d%s needs a 16-byte aligned K*N fp16 workspace", who, bits, g_idx ? " with g_idx" : "");
  rc = unpack_dequant(qweight, qzeros, scales, g_idx, workspace, K, N, bits, groupsize, 1, st);
  if (rc != SAMQ_OK) return rc;
  return launch_dense(x, workspace, b, r, out, M, K, N, epilogue, rowmap, st);
}

}  // namespace
}  // namespace samq

extern "C" int samq_dense_linear_fwd(const void* x, const void* wt, const void* bias,
                                     const void* residual, void* y, int64_t M, int K, int N,
                                     int epilogue, void* stream) {
  using namespace samq;
  int rc = check_common(x, y, M, K, N, epilogue, "samq_dense_linear_fwd");
  if (rc != SAMQ_OK) return rc;
  SAMQ_REQUIRE(wt && reinterpret_cast<uintptr_t>(wt) % 16 == 0, SAMQ_ERR_BAD_ARG,
               "samq_dense_linear_fwd: wt must be non-null and 16-byte aligned");
  const RowMap identity = {0, 0, 0, 0, 0, 0};
  return launch_dense(x, wt, reinterpret_cast<const __half*>(bias), reinterpret_cast<const __half*>(residual),
                      reinterpret_cast<__half*>(y), M, K, N, epilogue, identity,
                      reinterpret_cast<cudaStream_t>(stream));
}

// Hessian accumulation of the GPTQ solver on the tensor cores: H = beta H + alpha At Bt^T
extern "C" int samq_syrk_f32_fwd(const void* at, const void* bt, void* h, int n_feat, int64_t tokens, float alpha,
                                 float beta, void* stream) {
  using namespace samq;
  const void* xt = at;
  SAMQ_REQUIRE(at && bt && h, SAMQ_ERR_BAD_ARG, "samq_syrk_f32_fwd: null pointer");
  SAMQ_REQUIRE(n_feat > 0 && n_feat <= 65536, SAMQ_ERR_BAD_SHAPE, "samq_syrk_f32_fwd: n_feat=%d out of range", n_feat);
  SAMQ_REQUIRE(tokens > 0 && tokens % kBK == 0 && tokens < (1ll << 30), SAMQ_ERR_BAD_SHAPE,
               "samq_syrk_f32_fwd: tokens=%lld must be a positive multiple of %d (pad with zero columns)",
               (long long)tokens, kBK);
  SAMQ_REQUIRE(reinterpret_cast<uintptr_t>(at) % 16 == 0 && reinterpret_cast<uintptr_t>(bt) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(h) % 16 == 0,
               SAMQ_ERR_BAD_ARG, "samq_syrk_f32_fwd: pointers must be 16-byte aligned");
  constexpr int BM = 192;
  using C = Cfg<BM, false>;
  const int K = static_cast<int>(tokens);
  const CUtensorMap* mx = get_tensor_map_2d(xt, n_feat, K, static_cast<uint64_t>(K) * 2, BM, kBK, 2, 3);
  const CUtensorMap* mw = get_tensor_map_2d(bt, n_feat, K, static_cast<uint64_t>(K) * 2, kBN, kBK, 2, 3);
  if (!mx || !mw) return SAMQ_ERR_LAUNCH;
  auto kern = qlinear_kernel<BM, false, false, 4, true>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes, "syrk"); rc != SAMQ_OK) return rc;
  const int64_t tiles = static_cast<int64_t>((n_feat + kBN - 1) / kBN) * ((n_feat + BM - 1) / BM);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  const RowMap identity = {0, 0, 0, 0, 0, 0};
  kern<<<grid, kThreads, C::kSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *mx, *mw, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<__half*>(h), n_feat, n_feat, K, K, identity,
      alpha, beta);
  count_launch();
  return check_launch("qlinear_kernel<syrk>");
}

extern "C" int samq_qlinear_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                                const void* scales, const int32_t* g_idx, const void* bias,
                                const void* residual, void* y, void* workspace, int64_t M, int K,
                                int N, int bits, int groupsize, int epilogue, void* stream) {
  const samq::RowMap identity = {0, 0, 0, 0, 0, 0};
  return samq::qlinear_impl("samq_qlinear_fwd", x, qweight, qzeros, scales, g_idx, bias, residual, y, workspace,
                            M, K, N, bits, groupsize, epilogue, identity, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int samq_qlinear_unpartition_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                                            const void* scales, const int32_t* g_idx, const void* bias,
                                            const void* shortcut, void* y, void* workspace, int B, int H,
                                            int W, int ws, int K, int N, int bits, int groupsize,
                                            void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && ws > 0, SAMQ_ERR_BAD_SHAPE,
               "samq_qlinear_unpartition_fwd: B=%d H=%d W=%d ws=%d", B, H, W, ws);
  SAMQ_REQUIRE(shortcut != nullptr, SAMQ_ERR_BAD_ARG, "samq_qlinear_unpartition_fwd: shortcut is required");
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  const int64_t M = static_cast<int64_t>(B) * nH * nW * ws * ws;
  const RowMap rm = {ws, H, W, nH, nW, 0};
  return qlinear_impl("samq_qlinear_unpartition_fwd", x, qweight, qzeros, scales, g_idx, bias, shortcut, y,
                      workspace, M, K, N, bits, groupsize, SAMQ_EPI_NONE, rm, reinterpret_cast<cudaStream_t>(stream));
}

namespace samq {
namespace {

// qkv rows of the zero-padding tokens of the window layout: x is exactly 0 there, so the reference's
// GEMM yields fp16(0 + bias) = bias (0 without a bias).  One warp per PAD row, enumerated in closed
// form -- an image has H * padW pad tokens in its last window column (every image row h, columns
// j >= rW of window ww = nW - 1) and padH * nW * ws in its last window row (rows i >= rH of the
// windows wh = nH - 1) -- so that all warps do the same amount of work (the earlier version gave a
// warp 32 consecutive windowed tokens and let it copy whichever of them were padding: most warps
// had none, a few had 32; 57 us per call for 197 MB of stores).
__device__ __forceinline__ void fill_pad_row(__half* __restrict__ y, const __half* __restrict__ bias, int N,
                                             const RowMap& rm, int64_t wid, int per_image, int colA, int rH, int rW,
                                             int padW, int lane) {
  const int b = static_cast<int>(wid / per_image);
  int p = static_cast<int>(wid - static_cast<int64_t>(b) * per_image);
  int wh, ww, i, j;
  if (p < colA) {
    const int h = p / padW;
    wh = h / rm.ws; i = h - wh * rm.ws; ww = rm.nW - 1; j = rW + (p - h * padW);
  } else {
    p -= colA;
    const int span = rm.nW * rm.ws;
    const int ii = p / span, rest = p - ii * span;
    wh = rm.nH - 1; i = rH + ii; ww = rest / rm.ws; j = rest - ww * rm.ws;
  }
  const int64_t row = ((static_cast<int64_t>(b) * rm.nH + wh) * rm.nW + ww) * rm.ws * rm.ws + i * rm.ws + j;
  const uint4* src = reinterpret_cast<const uint4*>(bias);
  uint4* dst = reinterpret_cast<uint4*>(y + row * N);
  for (int c = lane; c < N / 8; c += 32) dst[c] = bias ? src[c] : make_uint4(0, 0, 0, 0);
}

// Launched programmatically BEHIND the qkv GEMM, one block per SM (no shared memory, a few registers:
// it fits beside the GEMM's persistent CTA): the pad rows are disjoint from the rows the GEMM stores, so
// the 197 MB of stores (batch 32) run next to the GEMM instead of as a 34-63 us kernel of their own.
// The output buffer is free for early writes: the GEMM's CTAs have started, i.e. every kernel before it
// -- every possible reader of the memory the buffer re-uses -- has finished.  The final wait keeps
// "this grid is complete" implying "the GEMM is complete" for the kernel launched behind it.
__global__ void __launch_bounds__(256)
fill_pad_rows_kernel(__half* __restrict__ y, const __half* __restrict__ bias, int B, int N, RowMap rm) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int rH = rm.H - (rm.nH - 1) * rm.ws, rW = rm.W - (rm.nW - 1) * rm.ws;   // valid rows / columns of the last windows
  const int padH = rm.ws - rH, padW = rm.ws - rW;
  const int colA = rm.H * padW, colB = padH * rm.nW * rm.ws;
  const int per_image = colA + colB;
  const int64_t total = static_cast<int64_t>(B) * per_image;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t wid = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); wid < total; wid += stride)
    fill_pad_row(y, bias, N, rm, wid, per_image, colA, rH, rW, padW, lane);
  pdl_wait();
}

}  // namespace
}  // namespace samq

extern "C" int samq_qlinear_partition_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                                          const void* scales, const int32_t* g_idx, const void* bias, void* y,
                                          void* workspace, int B, int H, int W, int ws, int K, int N, int bits,
                                          int groupsize, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(B > 0 && H > 0 && W > 0 && ws > 0, SAMQ_ERR_BAD_SHAPE,
               "samq_qlinear_partition_fwd: B=%d H=%d W=%d ws=%d", B, H, W, ws);
  SAMQ_REQUIRE(N % 8 == 0, SAMQ_ERR_BAD_SHAPE, "samq_qlinear_partition_fwd: N=%d must be a multiple of 8", N);
  const int nH = (H + ws - 1) / ws, nW = (W + ws - 1) / ws;
  const int64_t M_img = static_cast<int64_t>(B) * H * W;
  const RowMap to_win = {ws, H, W, nH, nW, 1};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SAMQ_REQUIRE(y != nullptr, SAMQ_ERR_BAD_ARG, "samq_qlinear_partition_fwd: null output");
  int rc = qlinear_impl("samq_qlinear_partition_fwd", x, qweight, qzeros, scales, g_idx, bias, nullptr, y, workspace,
                        M_img, K, N, bits, groupsize, SAMQ_EPI_NONE, to_win, st);
  if (rc != SAMQ_OK) return rc;
  // pad rows (disjoint from the rows the GEMM stores): launched behind the GEMM, runs next to it
  if (nH * ws != H || nW * ws != W) {
    const int64_t pad_rows = static_cast<int64_t>(B) * (H * (nW * ws - W) + (nH * ws - H) * nW * ws);
    const int64_t blocks = (pad_rows + 7) / 8;
    const int grid = static_cast<int>(blocks < num_sms() ? blocks : num_sms());
    launch_pdl(1, fill_pad_rows_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<__half*>(y),
               reinterpret_cast<const __half*>(bias), B, N, to_win);
    count_launch();
    return check_launch("fill_pad_rows_kernel");
  }
  return SAMQ_OK;
}

extern "C" int samq_qlinear_prefetch(const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                     void* workspace, int K, int N, int bits, int groupsize, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(bits == 4, SAMQ_ERR_UNSUPPORTED_BITS, "samq_qlinear_prefetch: int4 only (got %d bits)", bits);
  return prefetch_dequant4(qweight, qzeros, scales, workspace, K, N, groupsize, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}
