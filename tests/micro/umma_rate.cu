// Microbenchmark: issue rate of tcgen05.mma kind::f16 (M=128) for SS vs TS operands and
// several N, one CTA per SM, no data dependencies on memory (operands are whatever is in
// smem / TMEM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include "../../sam_quantization_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace samq;

template <bool TS, bool PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int rounds, int per_commit, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (PAIR) cluster_sync_all();
  if (warp == 0) { if (PAIR) tmem_alloc_pair(&slot, 512); else tmem_alloc(&slot, 512); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const bool issuer = warp == 1 && (!PAIR || cluster_ctarank() == 0);
  if (issuer) {
    const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, n, 0);
    const uint64_t a_desc = make_smem_desc(smem_u32(smem), 0, 1024, kLayoutSw128);
    const uint64_t b_desc = make_smem_desc(smem_u32(smem + 16384), 0, 1024, kLayoutSw128);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
        for (int i = 0; i < per_commit; ++i) {
          const int k = i & 3;
          if (TS) {
            if (PAIR) tc_mma_ts_pair(tmem, tmem + 384 + k * 8, b_desc + k * 2, idesc, 1);
            else tc_mma_ts(tmem, tmem + 384 + k * 8, b_desc + k * 2, idesc, 1);
          } else {
            tc_mma_ss(tmem, a_desc + k * 2, b_desc + k * 2, idesc, 1);
          }
        }
        if (PAIR) tc_commit_pair(&bar, 1); else tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);   // one batch in flight at a time (a lagging wait could be lapped)
      ph ^= 1;
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// issue batches of 4 TS MMAs separated by a `delay`-cycle busy wait; one commit at the very end.
// If the tensor pipe queues MMAs deeply, clk/MMA stays ~N/2 until delay approaches 4*N/2.
__global__ void __launch_bounds__(128, 1) delay_kernel(int n, int batches, int delay, int waits, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar, dummy;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, n, 0);
    const uint64_t b_desc = make_smem_desc(smem_u32(smem + 16384), 0, 1024, kLayoutSw128);
    long long t0 = clock64();
    for (int r = 0; r < batches; ++r) {
      // `waits` already-complete barrier waits (parity 1 of a fresh barrier passes immediately)
      for (int w = 0; w < waits; ++w) mbar_wait(&dummy, 1);
      tc_fence_after();
      if (elect_one()) {
        for (int k = 0; k < 4; ++k) tc_mma_ts(tmem, tmem + 384 + k * 8, b_desc + k * 2, idesc, 1);
      }
      __syncwarp();
      const long long t = clock64();
      while (clock64() - t < delay) {}
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

void run_delay(int n, int delay, int waits) {
  long long* d; cudaMalloc(&d, 148 * 8); cudaMemset(d, 0, 148 * 8);
  const int batches = 2000;
  cudaFuncSetAttribute(delay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  delay_kernel<<<148, 128, 100 * 1024>>>(n, batches, delay, waits, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0; for (int i = 0; i < 148; ++i) sum += h[i];
  printf("delay N=%3d delay=%4d waits=%d: %7.1f clk per 4-MMA batch (ideal %d)  %s\n", n, delay, waits, sum / 148 / batches, 2 * n, cudaGetErrorString(e2));
  cudaFree(d);
}

template <bool TS, bool PAIR>
void run(const char* name, int n, int per_commit) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaMemset(d, 0, 148 * 8);
  const int rounds = 200;
  auto k = rate_kernel<TS, PAIR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, n, rounds, per_commit, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0; int cnt = 0;
  for (int i = 0; i < 148; ++i) if (h[i] > 0) { sum += h[i]; ++cnt; }
  double cyc = cnt ? sum / cnt / (double(rounds) * per_commit) : 0;
  printf("%-10s N=%3d per_commit=%2d: %7.1f clk/MMA  (ideal %5.1f)  %s %s\n", name, n, per_commit, cyc, n / 2.0,
         cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  if (getenv("UMMA_DELAY")) {
    for (int delay : {0, 100, 200, 300, 400, 600}) run_delay(192, delay, 0);
    for (int waits : {1, 2, 3}) run_delay(192, 0, waits);
  }
  for (int n : {16, 32, 48, 80, 96}) {
    run<false, false>("SS 1cta", n, 64);
    run<true, false>("TS 1cta", n, 64);
  }
  for (int pc : {64}) {
    for (int n : {64, 128, 192, 256}) {
      run<false, false>("SS 1cta", n, pc);
      run<true, false>("TS 1cta", n, pc);
      run<true, true>("TS 2cta", n, pc);
    }
  }
  return 0;
}
