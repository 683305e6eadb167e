/*
 * samq.h -- C ABI of libsamq.so: the B200 (sm_100a) implementation of the
 * GPTQ-quantized SAM image-encoder hot path of zhanglei1172/sam-quantization.
 *
 * The reference exposes no FFI: its boundary is a Python module-swap API
 * (gptq_triton/__init__.py:8-12).  Each entry point below replaces one Triton
 * launch or one chain of eager ATen ops of that path; the reference location it
 * stands in for is cited per function.  The Python host package
 * (sam_quantization_b200/) validates shapes/dtypes with the reference's own
 * exception types and then calls these through ctypes.
 *
 * Conventions
 *   - plain pointers (device memory unless stated), sizes, and a cudaStream_t
 *     passed as void*; no torch types.
 *   - every call is asynchronous on `stream`, allocates nothing, frees nothing
 *     and keeps no mutable global state except a small cache of TMA descriptors
 *     (immutable once built) and the thread-local last-error string.
 *   - return value: SAMQ_OK (0) or a negative samq_status; samq_last_error()
 *     gives a human-readable reason for the calling thread.
 *   - fp16 is IEEE binary16 (the reference runs model.half(),
 *     gptq4sam_infer.py:218); all tensors are dense row-major.
 */
#ifndef SAMQ_H_
#define SAMQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum samq_status {
  SAMQ_OK = 0,
  SAMQ_ERR_BAD_SHAPE = -1,       /* reference: AssertionError, quant_linear.py:378-399 */
  SAMQ_ERR_UNSUPPORTED_BITS = -2, /* reference: NotImplementedError, quant_linear.py:72-73 */
  SAMQ_ERR_UNSUPPORTED_ARCH = -3, /* reference: RuntimeError, fused_attention.py:314-318 */
  SAMQ_ERR_LAUNCH = -4,          /* CUDA launch / driver error */
  SAMQ_ERR_BAD_ARG = -5          /* null pointer, misaligned pointer, bad enum */
} samq_status;

/* epilogue selector of samq_qlinear_fwd */
typedef enum samq_epilogue {
  SAMQ_EPI_NONE = 0, /* y = x.W + bias                         (quant_linear.py:431-435) */
  SAMQ_EPI_GELU = 1  /* y = gelu_erf(x.W + bias)  lin1 + act   (common.py:25-26)          */
} samq_epilogue;

/* relw_mode of samq_attn_relpos_fwd: "reference" reproduces the fork's
 * torch.matmul broadcasting (fused_attention.py:76-78, rel_w indexed by the
 * query ROW), "upstream" is Meta's einsum("bhwc,wkc->bhwk"). */
typedef enum samq_relw_mode {
  SAMQ_RELW_REFERENCE = 0,
  SAMQ_RELW_UPSTREAM = 1
} samq_relw_mode;

/* Library / device ---------------------------------------------------------- */

/* ABI version of this header (bumped on any signature change or new entry point).
 * 1: round 1.  2: single-FMA dequant (the reference's Triton rounding), samq_config_reload /
 * samq_has_ablations.  3: samq_gather_cols_fwd, samq_syrk_f32_fwd, samq_gptq_block_fwd,
 * samq_attn_small_fwd, samq_small_linear_fwd, samq_gelu_fwd.  4: samq_qlinear_prefetch; the
 * samq_qlinear_* entry points accept NULL packed pointers + a prefetched workspace. */
int samq_abi_version(void);

/* Reason for the last non-OK status returned to the calling thread. */
const char* samq_last_error(void);

/* SAMQ_OK iff the current CUDA device is compute capability 10.x (B200).
 * Replaces the capability guard at fused_attention.py:314-318. */
int samq_device_check(void);

/* Number of kernels this library has launched in this process (all threads);
 * bench.py reads it around the timed region for "gpu_launches". */
uint64_t samq_launch_count(void);

/* The library's developer switches (SAMQ_GEMM, SAMQ_DENSE, SAMQ_ATTN_MAX, SAMQ_PDL, and in
 * ablation builds SAMQ_ATTN_WIN / SAMQ_ATTN_GLOB) are read from the environment ONCE, at first
 * use -- not per call.  samq_config_reload() re-reads them (tests, A/B timing scripts).  The
 * reference's counterpart is its per-module autotune cache (quant_linear.py:122-229). */
void samq_config_reload(void);

/* 1 iff this build contains the earlier kernel generations (`make ABLATIONS=1`); the shipped
 * library returns 0. */
int samq_has_ablations(void);

/* Packed-weight unpack + dequantise ------------------------------------------
 * Replaces the in-kernel unpack of matmul4_kernel (quant_linear.py:291-301,
 * 312-313, 334-339) as a standalone pass:
 *   q[k,n] = field k%(32/bits) of qweight[k/(32/bits), n]      (2/4/8 bit)
 *   z[g,n] = field n%(32/bits) of qzeros[g, n/(32/bits)]
 *   W[k,n] = fp16( q*s[g,n] - fp16((z+1)*s[g,n]) ),  g = g_idx[k] or k/groupsize
 *            (ONE fused multiply-add after the separately rounded zero term: the rounding of the
 *            reference's Triton kernel on this GPU -- fma.rn.f16x2 -- pinned by the
 *            identity-matrix extraction, tests/golden/dequant_triton_b4.npz)
 * 3-bit uses the 32-values-in-3-words layout of quant.py:160-180.
 *   qweight int32 [K*bits/32, N], qzeros int32 [G, N*bits/32], scales fp16 [G, N],
 *   g_idx int32 [K] or NULL.
 * transposed == 0: w_out is fp16 [K, N];  transposed != 0: w_out is fp16 [N, K]. */
int samq_unpack_dequant(const int32_t* qweight, const int32_t* qzeros,
                        const void* scales, const int32_t* g_idx, void* w_out,
                        int K, int N, int bits, int groupsize, int transposed,
                        void* stream);

/* Mask decoder, prompt side ----------------------------------
 * Attention of the two-way transformer, replaces Attention.forward's
 * `softmax(q k^T / sqrt(c_per_head)) v` (segment_anything/modeling/transformer.py:
 * 225-238) between the prompt tokens and the image tokens, either direction:
 *   out[b, i, h*hd:(h+1)*hd] = sum_j softmax_j(scale q[b,i,h].k[b,j,h]) v[b,j,h]
 * q, out fp16 [B, Nq, heads*hd]; k, v fp16 [B, Nk, heads*hd]; hd in {16, 32, 64};
 * fp32 softmax and accumulation. */
int samq_attn_small_fwd(const void* q, const void* k, const void* v, void* out, int B, int heads,
                        int Nq, int Nk, int hd, float scale, void* stream);

/* Skinny linear for the decoder's prompt tokens and heads (nn.Linear on a few rows:
 * transformer.py:151-153 MLPBlock, mask_decoder.py:155-178 MLP, the q/k/v/out
 * projections of the token side):  y = act(x w^T + bias) + residual, x fp16 [M, K],
 * w fp16 [N, K], bias fp16 [N] or NULL, residual fp16 [M, N] or NULL (fp16 add after the
 * rounding, like the reference's `queries + mlp_out`), act 0 none / 1 exact-erf GELU /
 * 2 ReLU; any N, K % 8 == 0.  The decoder's 4096-token linears use samq_dense_linear_fwd. */
int samq_small_linear_fwd(const void* x, const void* w, const void* bias, const void* residual,
                          void* y, int64_t M, int N, int K, int act, void* stream);

/* Elementwise exact-erf GELU on fp16 (nn.GELU after the LayerNorm2d of
 * MaskDecoder.output_upscaling, mask_decoder.py:55-59); n even; y may alias x. */
int samq_gelu_fwd(const void* x, void* y, int64_t n, void* stream);

/* GPTQ calibration (offline solver) ---------------------------
 * Hessian accumulation, replaces GPTQ.add_batch (gptq.py:29-60: `self.H *= n/(n+1);
 * inp = sqrt(2/(n+1)) inp.float(); self.H += inp.matmul(inp.t())`) on the tensor cores:
 *   h[m, n] = beta h[m, n] + alpha sum_t at[m, t] bt[n, t]
 * at, bt fp16 [n_feat, tokens] row-major (the layer input TRANSPOSED, zero-padded to
 * tokens % 64 == 0; bt may equal at; any n_feat), h fp32 [n_feat, n_feat], fp32 accumulation.
 * fp32 inputs are fed as an fp16 split x = hi + 2^-11 lo (three calls: hi.hi, hi.lo,
 * lo.hi), which restores fp32 accuracy; see sam_quantization_b200/gptq.py. */
int samq_syrk_f32_fwd(const void* at, const void* bt, void* h, int n_feat, int64_t tokens,
                      float alpha, float beta, void* stream);

/* One column block of GPTQ.fasterquant's inner loop (gptq.py:108-141), all rows in
 * parallel, columns sequential.  For i in [0, ncols):  with (s, z) = scale/zero[r, colmap[i]]
 *   q = s (clamp(round(w_i / s) + z, 0, maxq) - z);  Q[r, col0+i] = q;
 *   e = (w_i - q) / U[col0+i, col0+i];  E[r, i] = e;  loss += (w_i - q)^2 / U_ii^2 / 2;
 *   w_j -= e U[col0+i, col0+j]  for j >= i (inside the block only)
 * W fp32 [rows, ldw] (read only: the block-entry values), U fp32 [ldu, ldu] upper Cholesky
 * factor of H^-1, scale / zero fp32 [rows, nparams], colmap int32 [ncols], Q fp32
 * [rows, ldq], E fp32 [rows, ncols], loss fp32 scalar (accumulated).  ncols <= 128. */
int samq_gptq_block_fwd(void* W, int ldw, int rows, int col0, int ncols, const void* U, int ldu,
                        const void* scale, const void* zero, int nparams, const int32_t* colmap,
                        int maxq, void* Q, int ldq, void* E, void* loss, void* stream);

/* Column gather  y[m, j] = x[m, perm[j]]  (x, y fp16 [M, K], perm int32 [K]).
 * No reference counterpart as a kernel: the reference's QuantLinear has no g_idx
 * (quant_linear.py:66-116); GPTQ act-order checkpoints (gptq.py:88-96 permute the
 * columns before rounding) store g_idx[k] = group of input feature k.  The host
 * sorts qweight's rows by group once (QuantLinear.sorted_pack) and gathers x with
 * this kernel, so that the fused in-SM dequant GEMM sees contiguous groups. */
int samq_gather_cols_fwd(const void* x, const int32_t* perm, void* y, int64_t M, int K,
                         void* stream);

/* Dequant-GEMM ---------------------------------------------------------------
 * Replaces triton_matmul4 + matmul4_kernel + the separate bias add
 * (quant_linear.py:231-352, 355-437), with optional fused GELU (common.py:26)
 * and optional fused residual add (image_encoder.py:204-205):
 *   y[M,N] = epi( x[M,K] . W[K,N] + bias[N] ) + residual[M,N]
 * x, y, residual fp16 row-major; bias fp16 [N] or NULL; residual NULL or [M,N]
 * (may alias y).  fp32 accumulation on tcgen05 tensor cores.
 * bits 2/3/4/8 with g_idx == NULL and groupsize % 64 == 0 run the fused
 * unpack-in-registers->TMEM->tcgen05 kernel when M < 2048 or workspace == NULL; for
 * longer M (where re-dequantising the weight tile for every 192-token tile costs more
 * than reading fp16 weights from L2) and whenever g_idx != NULL it runs unpack_dequant
 * into `workspace` (K*N fp16 device scratch, may be reused between calls on the same
 * stream) followed by the dense tcgen05 kernel.
 * Requires K % 64 == 0, N % 128 == 0 (the reference asserts K % 128 == 0 and
 * N % 256 == 0, quant_linear.py:389-395). */
int samq_qlinear_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                     const void* scales, const int32_t* g_idx, const void* bias,
                     const void* residual, void* y, void* workspace,
                     int64_t M, int K, int N, int bits, int groupsize,
                     int epilogue, void* stream);

/* Weight prefetch for the unpack-once path (int4, contiguous groups): unpacks the packed weight of
 * the NEXT QuantLinear into `workspace` (fp16 Wt[N, K]) with a small persistent grid that is launched
 * programmatically behind the kernel enqueued before it -- normally the current layer's GEMM -- and
 * runs NEXT TO it instead of between two GEMMs.  `workspace` must not be read by any kernel still in
 * flight on the stream (the host rotates three buffers).  The prefetched weight is consumed by calling
 * samq_qlinear_fwd / samq_qlinear_unpartition_fwd / samq_qlinear_partition_fwd with
 * qweight == qzeros == scales == g_idx == NULL and the same `workspace`: they then run only the dense
 * tcgen05 GEMM.  Same arithmetic and bits as samq_unpack_dequant (quant_linear.py:291-301, 334-339). */
int samq_qlinear_prefetch(const int32_t* qweight, const int32_t* qzeros, const void* scales,
                          void* workspace, int K, int N, int bits, int groupsize, void* stream);

/* proj GEMM fused with window_unpartition + crop + residual add
 * (QuantAttention.forward's o_proj, fused_attention.py:147, followed by
 * image_encoder.py:201-204, 309-333):  x holds the attention output of the WINDOWED
 * tokens, fp16 [B*nH*nW*ws*ws, K] (nH = ceil(H/ws)); the result row of window token
 * (b, wh, ww, i, j) is written to image token (b, wh*ws+i, ww*ws+j) of y fp16 [B, H, W, N]
 * as  shortcut[b,h,w,:] + (x.W + bias); padding tokens are dropped.  Other arguments as
 * samq_qlinear_fwd.  y may alias shortcut. */
int samq_qlinear_unpartition_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                                 const void* scales, const int32_t* g_idx, const void* bias,
                                 const void* shortcut, void* y, void* workspace, int B, int H,
                                 int W, int ws, int K, int N, int bits, int groupsize,
                                 void* stream);

/* qkv GEMM fused with window_partition (QuantAttention.forward's qkv_proj,
 * fused_attention.py:110, applied to the output of image_encoder.py:196-198, 282-306):
 * x holds the normalised tokens in IMAGE order, fp16 [B, H, W, K]; the result row of image token
 * (b, h, w) is written to window token (b, h/ws, w/ws, h%ws, w%ws) of y fp16
 * [B*nH*nW, ws, ws, N].  The rows of the zero-padding tokens, for which the reference multiplies
 * zero rows, are set to `bias` (0 without a bias) -- the value fp16(0 + bias) that GEMM produces --
 * without being multiplied.  Other arguments as samq_qlinear_fwd. */
int samq_qlinear_partition_fwd(const void* x, const int32_t* qweight, const int32_t* qzeros,
                               const void* scales, const int32_t* g_idx, const void* bias, void* y,
                               void* workspace, int B, int H, int W, int ws, int K, int N, int bits,
                               int groupsize, void* stream);

/* Dense fp16 GEMM on the same tcgen05 kernel: y = epi(x . Wt^T + bias) + residual
 * with Wt fp16 [N, K] (already dequantised).  Used for the ablation "dequantise
 * once, then GEMM" and by samq_qlinear_fwd for the non-int4 formats. */
int samq_dense_linear_fwd(const void* x, const void* wt, const void* bias,
                          const void* residual, void* y, int64_t M, int K, int N,
                          int epilogue, void* stream);

/* Attention with decomposed relative-position bias ---------------------------
 * Replaces the q-slice copy, add_decomposed_rel_pos (2 x get_rel_pos + 2 batched
 * matmuls) and the Triton flash kernel _fwd_kernel1 of QuantAttention.forward
 * (fused_attention.py:46-80, 107-133, 159-358):
 *   qkv  fp16 [B, H*W, 3, heads, hd]   (the qkv QuantLinear output, untouched)
 *   rel_pos_h fp16 [2H-1, hd], rel_pos_w fp16 [2W-1, hd]
 *   out  fp16 [B, H*W, heads*hd]
 *   out = softmax(scale * q k^T + rel_h[m, kh] + rel_w[m, kw]) v
 * rel_h / rel_w are rounded to fp16 as the reference's torch.matmul does.
 * Supported: hd in {64, 80}; (H, W) = (64, 64) global or (14, 14) windowed. */
int samq_attn_relpos_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                         void* out, int B, int H, int W, int heads, int hd,
                         float scale, int relw_mode, void* stream);

/* Windowed attention with window_unpartition + crop fused into the store
 * (QuantAttention.forward on the partitioned tokens, fused_attention.py:107-149, followed by
 * image_encoder.py:201-203, 309-333): qkv is the packed qkv GEMM output of the WINDOWED tokens,
 * fp16 [B*nH*nW, ws, ws, 3*heads*hd] (nH = ceil(H/ws), the zero-padding tokens included as keys
 * exactly like the reference); out is fp16 [B, H, W, heads*hd] in IMAGE order -- token (i, j) of
 * window (b, wh, ww) is written to (b, wh*ws+i, ww*ws+j), tokens outside the image are dropped
 * (each window tile leaves as one TMA box; out-of-bounds elements of a box are not written).
 * ws must be 14; other arguments as samq_attn_relpos_fwd. */
int samq_attn_relpos_unpartition_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                                     void* out, int B, int H, int W, int ws, int heads, int hd,
                                     float scale, int relw_mode, void* stream);

/* LayerNorm / window partition ------------------------------------------------
 * y = LayerNorm(x; gamma, beta, eps) over the last dim, fp32 statistics
 * (nn.LayerNorm(eps=1e-6), image_encoder.py:173,183; build_sam.py:72). */
int samq_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y,
                       int64_t rows, int C, float eps, void* stream);

/* LayerNorm fused with window_partition (image_encoder.py:191-195, 282-306;
 * generic form fq_vit/models/sam/image_encoder.py:481-507):
 * x fp16 [B, H, W, C] -> y fp16 [B*nH*nW, ws, ws, C], nH = ceil(H/ws); padded
 * tokens are written as zeros (the pad is applied AFTER the norm). */
int samq_layernorm_partition_fwd(const void* x, const void* gamma, const void* beta,
                                 void* y, int B, int H, int W, int C, int ws,
                                 float eps, void* stream);

/* window_unpartition + crop + residual add (image_encoder.py:201-204, 309-333):
 * out[b,h,w,:] = shortcut[b,h,w,:] + windows[window(b,h,w), h%ws, w%ws, :].
 * out may alias shortcut. */
int samq_unpartition_residual(const void* windows, const void* shortcut, void* out,
                              int B, int H, int W, int C, int ws, void* stream);

/* Patch extraction for PatchEmbed (image_encoder.py:411-442): the conv16x16/stride16 is the
 * GEMM  rows[(b,ph,pw), (c,i,j)] . Wconv[D, (c,i,j)]^T, so the image is re-laid-out once
 * (x fp16 [B,C,H,W] -> out fp16 [B*(H/P)*(W/P), C*P*P]) and fed to samq_dense_linear_fwd with
 * bias and the positional embedding as the residual. */
int samq_patchify_fwd(const void* x, void* out, int B, int C, int H, int W, int P, void* stream);

/* The neck's 3x3 convolution (segment_anything image_encoder.py:96-103: Conv2d(256, 256, 3,
 * padding=1, bias=False)) as a GEMM: x fp16 NHWC [B,H,W,C] -> out fp16 [B*H*W, 9*C] with
 * out[(b,h,w), (ky*3+kx)*C + c] = x[b, h+ky-1, w+kx-1, c] (zero outside the image); feed it to
 * samq_dense_linear_fwd with the weight laid out [O, (ky, kx, c)].  The result is NHWC. */
int samq_im2col3x3_fwd(const void* x, void* out, int B, int H, int W, int C, void* stream);

/* out = a + b over n fp16 elements (image_encoder.py:204; global-attention blocks). */
int samq_add(const void* a, const void* b, void* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAMQ_H_ */
