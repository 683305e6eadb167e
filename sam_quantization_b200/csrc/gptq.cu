// GPTQ solver, the sequential part: one block of columns (gptq.py:108-151).
//
// Within a block the reference walks the columns one at a time: round column i to its group's
// grid, divide the rounding error by the diagonal of the inverse-Hessian Cholesky factor and
// spread it over the block's remaining columns (Wb[:, i:] -= e (x) Ub[i, i:]).  The weight ROWS
// never interact, so one thread owns one row: the row's block columns live in shared memory
// (column-major: conflict-free), the factor's block in shared memory too (broadcast reads).  What
// crosses blocks -- W[:, b1:] -= Eb . U[b0:b1, b1:] -- is a plain GEMM and stays a library call on
// the host side (sam_quantization_b200/gptq.py).  The Hessian itself is accumulated on the tensor
// cores (samq_syrk_f32_fwd, qlinear.cu).
//
// Arithmetic mirrors the torch expressions of the reference step by step (true division, round to
// nearest even, product and subtraction rounded separately) so that the rounded weights match the
// CPU solver's: tests/golden/gptq_*.npz.
#include "common.cuh"

namespace samq {
namespace {

constexpr int kGptqRows = 128;     // rows (threads) per CTA
constexpr int kGptqMaxCols = 128;  // block size limit (the reference's blocksize)

__global__ void __launch_bounds__(kGptqRows)
gptq_block_kernel(float* __restrict__ W, int ldw, int rows, int col0, int ncols, const float* __restrict__ U, int ldu,
                  const float* __restrict__ scale, const float* __restrict__ zero, int nparams,
                  const int32_t* __restrict__ colmap, int maxq, float* __restrict__ Q, int ldq,
                  float* __restrict__ E, float* __restrict__ loss) {
  extern __shared__ __align__(16) float gsm_f[];
  float* sU = gsm_f;                              // [ncols][ncols]
  float* sW = sU + ncols * ncols;                 // [ncols][kGptqRows]
  int* sMap = reinterpret_cast<int*>(sW + ncols * kGptqRows);
  const int tid = threadIdx.x;
  const int r = blockIdx.x * kGptqRows + tid;
  for (int i = tid; i < ncols * ncols; i += kGptqRows)
    sU[i] = U[static_cast<size_t>(col0 + i / ncols) * ldu + col0 + i % ncols];
  for (int i = tid; i < ncols; i += kGptqRows) sMap[i] = colmap[i];
  // the row's block columns (coalesced over the CTA: threads read consecutive columns of one row)
  for (int rr = 0; rr < kGptqRows; ++rr) {
    const int row = blockIdx.x * kGptqRows + rr;
    if (row < rows)
      for (int c = tid; c < ncols; c += kGptqRows) sW[c * kGptqRows + rr] = W[static_cast<size_t>(row) * ldw + col0 + c];
  }
  __syncthreads();
  if (r >= rows) return;
  float acc = 0.f;
  const float fmaxq = static_cast<float>(maxq);
  for (int i = 0; i < ncols; ++i) {
    const int g = sMap[i];
    const float s = scale[static_cast<size_t>(r) * nparams + g], z = zero[static_cast<size_t>(r) * nparams + g];
    const float w = sW[i * kGptqRows + tid];
    // scale * (clamp(round(w / scale) + zero, 0, maxq) - zero)   (gptq.py:183-187)
    const float qi = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(w, s)), z), 0.f), fmaxq);
    const float q = __fmul_rn(s, __fsub_rn(qi, z));
    const float d = sU[i * ncols + i];
    const float diff = __fsub_rn(w, q);
    const float e = __fdiv_rn(diff, d);
    acc += __fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d)) * 0.5f;
    Q[static_cast<size_t>(r) * ldq + col0 + i] = q;
    E[static_cast<size_t>(r) * ncols + i] = e;
    for (int j = i; j < ncols; ++j)
      sW[j * kGptqRows + tid] = __fsub_rn(sW[j * kGptqRows + tid], __fmul_rn(e, sU[i * ncols + j]));
  }
  // warp-reduce the loss, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) atomicAdd(loss, acc);
}

}  // namespace
}  // namespace samq

extern "C" int samq_gptq_block_fwd(void* W, int ldw, int rows, int col0, int ncols, const void* U, int ldu,
                                   const void* scale, const void* zero, int nparams, const int32_t* colmap, int maxq,
                                   void* Q, int ldq, void* E, void* loss, void* stream) {
  using namespace samq;
  SAMQ_REQUIRE(W && U && scale && zero && colmap && Q && E && loss, SAMQ_ERR_BAD_ARG, "samq_gptq_block_fwd: null pointer");
  SAMQ_REQUIRE(rows > 0 && ncols > 0 && ncols <= kGptqMaxCols && col0 >= 0 && col0 + ncols <= ldw && col0 + ncols <= ldu &&
                   ldq >= col0 + ncols && nparams > 0 && maxq > 0,
               SAMQ_ERR_BAD_SHAPE, "samq_gptq_block_fwd: rows=%d col0=%d ncols=%d (<= %d) ldw=%d ldu=%d ldq=%d nparams=%d maxq=%d",
               rows, col0, ncols, kGptqMaxCols, ldw, ldu, ldq, nparams, maxq);
  const int smem = (ncols * ncols + ncols * kGptqRows) * 4 + ncols * 4;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gptq_block_kernel), smem, "gptq_block"); rc != SAMQ_OK)
    return rc;
  gptq_block_kernel<<<(rows + kGptqRows - 1) / kGptqRows, kGptqRows, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float*>(W), ldw, rows, col0, ncols, reinterpret_cast<const float*>(U), ldu,
      reinterpret_cast<const float*>(scale), reinterpret_cast<const float*>(zero), nparams, colmap, maxq,
      reinterpret_cast<float*>(Q), ldq, reinterpret_cast<float*>(E), reinterpret_cast<float*>(loss));
  count_launch();
  return check_launch("gptq_block_kernel");
}
