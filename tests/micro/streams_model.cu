// Model of the global attention kernel's softmax streams: W independent warps per SM sub-partition, each
// alternating a LATENCY phase (a dependent chain of global loads: no issue slots, no MUFU -- stands for
// the barrier / MMA round trips, TMEM loads, fences of a half-tile) and an EXP phase (64 x {FFMA, FADD,
// EX2, FADD} + 32 packs: the instruction mix of the kernel's pass 2).  Prints half-tiles per kclk per
// sub-partition and the implied MUFU utilisation for W = 1 .. 6, for several latency lengths.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tests/micro/streams_model tests/micro/streams_model.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

// 2^x for x in [-100, 30] on the FMA / ALU pipes: Cody-Waite split + degree-4 polynomial (rel. error 4e-5)
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -100.f);
  const float t = x + 12582912.f;                     // 1.5 * 2^23: the low mantissa bits hold round(x)
  const float f = x - (t - 12582912.f);               // [-0.5, 0.5]
  float p = fmaf(0.0096181291f, f, 0.055504109f);
  p = fmaf(p, f, 0.24022651f);
  p = fmaf(p, f, 0.69314718f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool STAGED, int POLY>
__global__ void __launch_bounds__(768, 1) model(const int* __restrict__ chain, int hops, int iters, float c, float mm,
                                                long long* out, unsigned* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[64], bw[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) { s[i] = (lane + i) * 0.01f; bw[i] = i * 0.001f; }
  int p = (blockIdx.x * 24 + warp) * 32 % 4096;
  float l = 0.f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int h = 0; h < hops; ++h) p = chain[p];          // latency phase: dependent L2 hits
    const float shift = mm + (p & 1) * 1e-6f;
    float s0 = 0.f, s1 = 0.f;
    if (STAGED) {
      // stage A: all arguments; stage B: all exponentials back to back; stage C: sums and packs
      float x[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("{ .reg .f32 t; fma.rn.f32 t, %1, %2, %3; sub.f32 %0, t, %4; }" : "=f"(x[i]) : "f"(s[i]), "f"(c), "f"(bw[i]), "f"(shift));
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = ex2f(x[i]);
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        asm volatile("add.f32 %0, %0, %1;" : "+f"(s0) : "f"(x[i]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(s1) : "f"(x[i + 1]));
        const __half2 hh = __floats2half2_rn(x[i], x[i + 1]);
        acc ^= *reinterpret_cast<const unsigned*>(&hh);
        s[i] += 1e-3f * x[i + 1]; s[i + 1] += 1e-3f * x[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float x0 = fmaf(s[i], c, bw[i]) - shift, x1 = fmaf(s[i + 1], c, bw[i + 1]) - shift;
        const float p0 = (POLY && (i / 2) % POLY == 0) ? ex2_poly(x0) : ex2f(x0);    // one exponential in 2 * POLY
        const float p1 = ex2f(x1);
        s0 += p0; s1 += p1;
        const __half2 hh = __floats2half2_rn(p0, p1);
        acc ^= *reinterpret_cast<const unsigned*>(&hh);
        s[i] += 1e-3f * p1; s[i + 1] += 1e-3f * p0;        // keeps the next iteration's inputs live (2 more FMAs per pair)
      }
    }
    l += s0 + s1;
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 24 + warp] = t1 - t0;
  if (acc == 0x1234567u || l == 1.2345f) sink[0] = acc;
}

int main() {
  int* chain; long long* out; unsigned* sink;
  int h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (i * 1237 + 331) % 4096;
  cudaMalloc(&chain, sizeof(h)); cudaMemcpy(chain, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 24 * 8); cudaMalloc(&sink, 4);
  const int iters = 400;
  for (int staged = 0; staged < 5; ++staged)
  for (int hops : {0, 2}) {
    if (staged == 1) continue;
    for (int W : {1, 2, 3, 4}) {
      cudaMemset(out, 0, 148 * 24 * 8);
      if (staged == 1) model<true, 0><<<148, 128 * W>>>(chain, hops, iters, 0.16f, 3.f, out, sink);
      else if (staged == 0) model<false, 0><<<148, 128 * W>>>(chain, hops, iters, 0.16f, 3.f, out, sink);
      else if (staged == 2) model<false, 4><<<148, 128 * W>>>(chain, hops, iters, 0.16f, 3.f, out, sink);   // 1/8 on the FMA pipes
      else if (staged == 3) model<false, 2><<<148, 128 * W>>>(chain, hops, iters, 0.16f, 3.f, out, sink);   // 1/4
      else model<false, 1><<<148, 128 * W>>>(chain, hops, iters, 0.16f, 3.f, out, sink);                    // 1/2
      cudaError_t e = cudaDeviceSynchronize();
      long long r[148 * 24]; cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148 * 24; ++i) mx = r[i] > mx ? r[i] : mx;
      const double per_iter = mx / iters;                       // clk per half-tile of one warp
      const double rate = W * 1000.0 / per_iter;                // half-tiles per kclk per sub-partition
      printf("%s latency hops %d, %d warps/sub-partition: %7.0f clk per half-tile per warp, %5.2f half-tiles/kclk/sub-partition, MUFU busy %4.1f %%  (%s)\n",
             staged == 0 ? "interleaved" : staged == 1 ? "STAGED     " : staged == 2 ? "poly 1/8   " : staged == 3 ? "poly 1/4   " : "poly 1/2   ", hops, W, per_iter, rate, rate * 64 * 8 / 10.0, cudaGetErrorString(e));
    }
  }
  return 0;
}
