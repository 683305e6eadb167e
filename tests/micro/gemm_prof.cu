// Developer microbenchmark: clock64 breakdown of the CTA-pair dense GEMM (dense2_kernel).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSAMQ_GEMM_PROFILE --expt-relaxed-constexpr \
//        -I include tests/micro/gemm_prof.cu sam_quantization_b200/csrc/runtime.cu -o tests/micro/gemm_prof -lcuda
#include "../../sam_quantization_b200/csrc/qlinear_dense2.cu"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 32768, K = argc > 2 ? atoi(argv[2]) : 1280, N = argc > 3 ? atoi(argv[3]) : 5120;
  const int epi = argc > 4 ? atoi(argv[4]) : 1, with_res = argc > 5 ? atoi(argv[5]) : 0;
  std::vector<__half> hx(static_cast<size_t>(M) * K), hw(static_cast<size_t>(N) * K), hb(N);
  unsigned s = 1234567u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.f - 0.5f; };
  for (auto& v : hx) v = __float2half(rnd());
  for (auto& v : hw) v = __float2half(rnd() * 0.05f);
  for (auto& v : hb) v = __float2half(rnd());
  __half *x, *w, *b, *y, *res;
  cudaMalloc(&x, hx.size() * 2); cudaMalloc(&w, hw.size() * 2); cudaMalloc(&b, hb.size() * 2);
  cudaMalloc(&y, static_cast<size_t>(M) * N * 2); cudaMalloc(&res, static_cast<size_t>(M) * N * 2);
  cudaMemcpy(x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(b, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(res, 0, static_cast<size_t>(M) * N * 2);
  samq::RowMap rm{0, 0, 0, 0, 0, 0};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    int rc = samq::launch_dense_pair(x, w, b, with_res ? res : nullptr, y, M, K, N, epi, rm, 148, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d  %.1f us  %.1f TFLOP/s (%s)\n", rc, ms * 1e3, 2.0 * M * K * N / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
#ifdef SAMQ_GEMM_PROFILE
  long long prof[2][18][8];
  cudaMemcpyFromSymbol(prof, samq::g_gemm_prof, sizeof(prof));
  printf("pair 5, per warp totals (clk): epilogue: wait acc_full | acc_full->release | total busy | max release | max tile ;  MMA (last warp): wait acc_empty | wait full\n");
  for (int r = 0; r < 2; ++r)
    for (int wp = 0; wp < (epi ? 18 : 10); ++wp)
      printf("cta %d warp %d: %9lld %9lld %9lld %9lld %9lld\n", r, wp, prof[r][wp][0], prof[r][wp][1], prof[r][wp][2], prof[r][wp][3], prof[r][wp][4]);
#endif
  return 0;
}
