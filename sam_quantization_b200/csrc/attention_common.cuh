// Shared pieces of the attention translation units (attention_win.cu, attention_glob.cu,
// attention.cu and -- only with -DSAMQ_ABLATIONS -- attention_ablations.cu).
#pragma once
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace samq {

constexpr float kLog2e = 1.4426950408889634f;

// product kernels (one translation unit each); hd = 64 | 80
int attn_win3_dispatch(int hd, const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads,
                       float scale, int relw_mode, int img_h, int img_w, cudaStream_t st);
int attn_glob3_dispatch(int hd, const void* qkv, const void* rph, const void* rpw, void* out, int B, int heads,
                        float scale, int relw_mode, cudaStream_t st);
#ifdef SAMQ_ABLATIONS
// earlier designs, kept for A/B timing only: `make ABLATIONS=1` (never in the shipped library)
int attn_ablation_dispatch(int generation, bool glob, int hd, const void* qkv, const void* rph, const void* rpw,
                           void* out, int B, int heads, float scale, int relw_mode, cudaStream_t st);
#endif

namespace {
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

}  // namespace
}  // namespace samq
