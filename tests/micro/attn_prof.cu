// Developer microbenchmark: wait-time breakdown of the global attention kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSAMQ_ATTN_PROFILE --expt-relaxed-constexpr \
//        -I include tests/micro/attn_prof.cu sam_quantization_b200/csrc/runtime.cu -o tests/micro/attn_prof -lcuda
// (one translation unit on purpose: the profile counters live in an anonymous namespace)
#include "../../sam_quantization_b200/csrc/attention.cu"
#include "../../sam_quantization_b200/csrc/attention_win.cu"
#include "../../sam_quantization_b200/csrc/attention_glob.cu"
#include <cstdio>
#include <vector>
__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}
// SM clock right now: one block spinning for a fixed number of clock64 ticks, timed with events
static double sm_mhz() {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  spin_kernel<<<1, 32>>>(2000000);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return 2000000.0 / (ms * 1e3);
}
int main(int argc, char** argv) {
  const int hd = argc > 1 ? atoi(argv[1]) : 80;
  const int B = argc > 2 ? atoi(argv[2]) : 8;
  const int E = argc > 3 ? atoi(argv[3]) : 64;
  const int heads = 16, S = E * E, D = heads * hd;
  std::vector<__half> h(static_cast<size_t>(B) * S * 3 * D);
  unsigned x = 12345u;
  for (auto& v : h) { x = x * 1664525u + 1013904223u; v = __float2half(((x >> 8) & 0xffff) / 65536.f - 0.5f); }
  std::vector<__half> rp(static_cast<size_t>(2 * E - 1) * hd);
  for (auto& v : rp) { x = x * 1664525u + 1013904223u; v = __float2half((((x >> 8) & 0xffff) / 65536.f - 0.5f) * 0.2f); }
  __half *qkv, *rph, *rpw, *out;
  cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&rph, rp.size() * 2); cudaMalloc(&rpw, rp.size() * 2);
  cudaMalloc(&out, static_cast<size_t>(B) * S * D * 2);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(rph, rp.data(), rp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(rpw, rp.data(), rp.size() * 2, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = argc > 4 ? atoi(argv[4]) : 3;
  for (int it = 0; it < iters; ++it) {
    cudaEventRecord(e0);
    int rc = samq_attn_relpos_fwd(qkv, rph, rpw, out, B, E, E, heads, hd, 1.f / sqrtf(hd), 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it < 3 || it == iters - 1) printf("rc %d  %.1f us  (%s)\n", rc, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  printf("SM clock after the runs: %.0f MHz\n", sm_mhz());
#ifdef SAMQ_ATTN_PROFILE
  long long prof[12][8];
  cudaMemcpyFromSymbol(prof, samq::g_attn_prof, sizeof(prof));
  if (E == 14) {
    printf("windowed v4, one CTA, totals over its units (clk): wait s_full | max pass | exchange barrier | (wait t_full) | (wait o_full) | drain total | gather total | exp pass\n");
    for (int w = 0; w < 8; ++w)
      printf("wg %c warp %d: %7lld %7lld %7lld %7lld %7lld %7lld %7lld %7lld\n", w < 4 ? 'A' : 'B', w & 3, prof[w][0], prof[w][1],
             prof[w][2], prof[w][3], prof[w][4], prof[w][5], prof[w][6], prof[w][7]);
    printf("MMA warps (clk totals): wait q_full | wait o_free | T issue->done | wait k_full+t_done | QK issue->done | wait v_full+p_full | PV issue->done\n");
    for (int w = 9; w < 11; ++w)
      printf("MMA tile %c: %7lld %7lld %7lld %7lld %7lld %7lld %7lld\n", w == 9 ? 'A' : 'B', prof[w][0], prof[w][1], prof[w][2],
             prof[w][3], prof[w][4], prof[w][5], prof[w][6]);
    return 0;
  }
  const char* names[10] = {"sm g0 e0", "sm g0 e1", "sm g0 e2", "sm g0 e3", "sm g1 e0", "sm g1 e1", "sm g1 e2", "sm g1 e3", "TMA", "MMA"};
  printf("per-CTA totals over 32 key tiles (clk): softmax = wait s_full | max-exchange barrier | ex2+next-tile phase; "
         "TMA = wait k_empty | v_empty; MMA = wait k_full | p_full | v_full\n");
  for (int w = 0; w < 10; ++w)
    printf("%-9s %8lld %8lld %8lld | stamps: start-of-role %6lld  t_full %6lld  loop-start %6lld  loop-end %6lld  end %6lld\n", names[w],
           prof[w][0], prof[w][1], prof[w][2], prof[w][3], prof[w][4], prof[w][5], prof[w][6], prof[w][7]);
#endif
  return 0;
}
