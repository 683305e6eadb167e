// Microbenchmark: tcgen05.ld throughput per SM with 1 .. 8 warps loading concurrently (each warp reads its
// own 32-lane quarter; warps w and w + 4 share a quarter), x32 loads of fp32 columns, one CTA per SM.
// Answers whether the TMEM read port is per SM or per sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I include \
//        tests/micro/tmem_bw.cu sam_quantization_b200/csrc/runtime.cu -o tests/micro/tmem_bw -lcuda
#include "../../sam_quantization_b200/csrc/common.cuh"
#include <cstdio>
using namespace samq;

__global__ void __launch_bounds__(256, 1) ld_kernel(int active_warps, int iters, long long* out, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < active_warps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32], q[32];
      tmem_ld_x32(tm + ((i & 3) * 64), r);
      tmem_ld_x32(tm + ((i & 3) * 64) + 32, q);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += r[j] ^ q[j];
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < active_warps) out[blockIdx.x * 8 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; unsigned* sink;
  cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int aw : {1, 2, 4, 8}) {
    cudaMemset(d, 0, 148 * 8 * 8);
    ld_kernel<<<148, 256>>>(aw, iters, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148 * 8; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = double(aw) * iters * 2 * 32 * 32 * 4;   // per SM
    printf("%d warps: %.0f clk for %d x 2 x32 loads each -> %.1f B/clk/SM, %.1f clk per x32 load per warp  (%s)\n", aw, mx, iters,
           bytes / mx, mx / (2.0 * iters), cudaGetErrorString(e));
  }
  return 0;
}
