"""Tensor-level wrappers over the C ABI (include/samq.h).

Each function validates dtype / contiguity / device on the host (raising the
exception types the reference raises), allocates the output with torch
and launches the CUDA kernel on torch's current stream.  No eager fallback exists.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib


#: M from which samq_qlinear_fwd switches int4 from the fused kernel to unpack-once + dense GEMM
TWO_KERNEL_MIN_M = 2048   # keep in sync with kTwoKernelMinM in csrc/qlinear.cu

#: M from which a linked QuantLinear prefetches the next layer's weight next to its own GEMM
#: (quant_linear.py::WeightPrefetchChain).  The prefetch grid is one small block per SM -- what fits
#: beside the GEMM's CTA -- so it needs ~8x the 6 us the stand-alone unpack takes, and it pays only
#: where the GEMM it hides under is long.  Measured on one B200 (ViT-H int4, images/s without -> with):
#: batch 1 138.7 -> 131.5, 2 167.1 -> 163.1, 4 179.7 -> 177.5, 8 186.5 -> 187.2, 16 190.0 -> 190.6,
#: 32 181.3 -> 183.7 (+1.3 %; another box 188.6 -> 190.8).
PREFETCH_MIN_M = 32768


def _dev_ctx(t: torch.Tensor):
    _lib.require_cuda(t, "input")
    _lib.device_check(t.device)
    return torch.cuda.device(t.device)


def _check_half(t: torch.Tensor, name: str) -> None:
    assert t.dtype == torch.float16, f"{name} must be float16 (got {t.dtype})"
    assert t.is_contiguous(), f"{name} must be contiguous"


def num_groups(infeatures: int, groupsize: int) -> int:
    gs = infeatures if groupsize == -1 else groupsize
    return math.ceil(infeatures / gs)


def unpack_dequant(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, bits: int,
                   groupsize: int, g_idx: Optional[torch.Tensor] = None,
                   transposed: bool = False) -> torch.Tensor:
    """Dequantised fp16 weight ``[K, N]`` (or ``[N, K]`` if transposed)."""
    if bits not in (2, 3, 4, 8):
        raise NotImplementedError("Only 2, 3, 4 and 8 bits are supported.")
    assert qweight.dtype == torch.int32 and qzeros.dtype == torch.int32, "qweight/qzeros must be int32"
    assert qweight.is_contiguous() and qzeros.is_contiguous(), "qweight/qzeros must be contiguous"
    _check_half(scales, "scales")
    N = qweight.shape[1]
    K = qweight.shape[0] * 32 // bits
    if g_idx is not None:
        assert g_idx.dtype == torch.int32 and g_idx.numel() == K and g_idx.is_contiguous()
    with _dev_ctx(qweight):
        out = torch.empty((N, K) if transposed else (K, N), dtype=torch.float16, device=qweight.device)
        _lib.check(_lib.load().samq_unpack_dequant(
            _lib.ptr(qweight), _lib.ptr(qzeros), _lib.ptr(scales), _lib.ptr(g_idx), _lib.ptr(out),
            K, N, bits, groupsize, 1 if transposed else 0, _lib.stream_ptr(qweight.device)))
    return out


def _check_packed(x: torch.Tensor, qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                  g_idx: Optional[torch.Tensor], bias: Optional[torch.Tensor], bits: int, groupsize: int,
                  K: int, N: int) -> None:
    """One validation of the packed operands for every QuantLinear entry point: dtype, device,
    contiguity and the reference's checkpoint shapes (quant_linear.py:96-110): a wrong buffer must
    be an AssertionError here, never an out-of-bounds device read."""
    assert qweight.dtype == torch.int32 and qzeros.dtype == torch.int32, "qweight/qzeros must be int32"
    assert qweight.is_contiguous() and qzeros.is_contiguous(), "qweight/qzeros must be contiguous"
    _check_half(scales, "scales")
    assert qweight.dim() == 2 and tuple(qweight.shape) == (K * bits // 32, N), \
        f"qweight must be [{K * bits // 32}, {N}] (got {tuple(qweight.shape)})"
    gs = K if groupsize == -1 else groupsize
    assert gs > 0 and K % gs == 0, "infeatures must be a multiple of groupsize"
    G = K // gs
    assert tuple(scales.shape) == (G, N), f"scales must be [{G}, {N}] (got {tuple(scales.shape)})"
    assert tuple(qzeros.shape) == (G, N * bits // 32), \
        f"qzeros must be [{G}, {N * bits // 32}] (got {tuple(qzeros.shape)})"
    tensors = [("qweight", qweight), ("qzeros", qzeros), ("scales", scales)]
    if g_idx is not None:
        assert g_idx.dtype == torch.int32 and g_idx.numel() == K and g_idx.is_contiguous(), \
            "g_idx must be a contiguous int32 vector of infeatures entries"
        tensors.append(("g_idx", g_idx))
    if bias is not None:
        _check_half(bias, "bias")
        assert bias.numel() == N, "bias must have outfeatures entries"
        tensors.append(("bias", bias))
    for name, t in tensors:
        assert t.device == x.device, f"{name} is on {t.device}, x on {x.device}"


ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def small_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                 residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(x @ weight.T + bias) + residual`` for a few rows (mask-decoder prompt tokens and heads):
    ``x[..., K]`` fp16, ``weight[N, K]`` fp16 (``nn.Linear`` layout) -> ``[..., N]`` fp16."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x"); _check_half(weight, "weight")
    N, K = weight.shape
    assert x.shape[-1] == K and K % 8 == 0 and weight.device == x.device
    M = x.numel() // K
    if bias is not None:
        _check_half(bias, "bias")
        assert bias.numel() == N and bias.device == x.device
    if residual is not None:
        _check_half(residual, "residual")
        assert residual.numel() == M * N and residual.device == x.device
    with _dev_ctx(x):
        y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float16, device=x.device)
        if M == 0:
            return y
        _lib.check(_lib.load().samq_small_linear_fwd(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(residual),
                                                     _lib.ptr(y), M, N, K, act, _lib.stream_ptr(x.device)))
    return y


def gelu(x: torch.Tensor) -> torch.Tensor:
    """Exact-erf GELU, elementwise, fp16."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    assert x.numel() % 2 == 0
    with _dev_ctx(x):
        y = torch.empty_like(x)
        if x.numel():
            _lib.check(_lib.load().samq_gelu_fwd(_lib.ptr(x), _lib.ptr(y), x.numel(), _lib.stream_ptr(x.device)))
    return y


def attn_small(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """``softmax(q k^T / sqrt(hd)) v`` per head without positional bias: ``q[B, Nq, C]``,
    ``k, v[B, Nk, C]`` fp16, ``C = heads * hd`` with ``hd`` in {16, 32, 64} (transformer.py:225-238)."""
    _lib.require_cuda(q, "q")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _check_half(t, n)
    B, Nq, C = q.shape
    Nk = k.shape[1]
    assert k.shape == v.shape == (B, Nk, C) and C % heads == 0 and k.device == q.device == v.device
    hd = C // heads
    if hd not in (16, 32, 64):
        raise NotImplementedError(f"head dim {hd}: the decoder kernel serves 16, 32 and 64")
    with _dev_ctx(q):
        out = torch.empty_like(q)
        _lib.check(_lib.load().samq_attn_small_fwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(out), B, heads, Nq,
                                                   Nk, hd, 1.0 / math.sqrt(hd), _lib.stream_ptr(q.device)))
    return out


def hessian_accumulate(H: torch.Tensor, x: torch.Tensor, alpha: float, beta: float) -> None:
    """``H = beta H + alpha x^T x`` in place for the GPTQ solver (``GPTQ.add_batch``): ``H`` fp32
    ``[C, C]``, ``x`` ``[tokens, C]`` fp16 or fp32, on the tensor cores (``samq_syrk_f32_fwd``).
    fp32 ``x`` is split as ``hi + 2^-11 lo`` with fp16 ``hi, lo`` and accumulated as
    ``hi hi^T + 2^-11 (hi lo^T + lo hi^T)`` -- the dropped ``lo lo^T`` term is 2^-22 relative."""
    _lib.require_cuda(x, "x")
    assert H.dtype == torch.float32 and H.is_contiguous() and H.dim() == 2 and H.shape[0] == H.shape[1]
    C = H.shape[0]
    assert x.dim() == 2 and x.shape[1] == C and H.device == x.device
    tokens = x.shape[0]
    pad = (tokens + 63) // 64 * 64
    lib = _lib.load()
    with _dev_ctx(x):
        def transposed(v):          # [C, pad] fp16, zero columns beyond `tokens`
            t = torch.zeros((C, pad), dtype=torch.float16, device=x.device)
            t[:, :tokens] = v.t()
            return t
        st = _lib.stream_ptr(x.device)
        if x.dtype == torch.float16:
            xt = transposed(x)
            _lib.check(lib.samq_syrk_f32_fwd(_lib.ptr(xt), _lib.ptr(xt), _lib.ptr(H), C, pad, alpha, beta, st))
            return
        xf = x.float()
        hi = xf.half()
        lo = ((xf - hi.float()) * 2048.0).half()
        hit, lot = transposed(hi), transposed(lo)
        _lib.check(lib.samq_syrk_f32_fwd(_lib.ptr(hit), _lib.ptr(hit), _lib.ptr(H), C, pad, alpha, beta, st))
        _lib.check(lib.samq_syrk_f32_fwd(_lib.ptr(hit), _lib.ptr(lot), _lib.ptr(H), C, pad, alpha / 2048.0, 1.0, st))
        _lib.check(lib.samq_syrk_f32_fwd(_lib.ptr(lot), _lib.ptr(hit), _lib.ptr(H), C, pad, alpha / 2048.0, 1.0, st))


def gptq_block(W: torch.Tensor, col0: int, ncols: int, U: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor,
               colmap: torch.Tensor, maxq: int, Q: torch.Tensor, loss: torch.Tensor) -> torch.Tensor:
    """One column block of the GPTQ rounding loop on the device (``samq_gptq_block_fwd``); returns the
    scaled errors ``E[rows, ncols]`` that the caller propagates to the later blocks."""
    _lib.require_cuda(W, "W")
    for t in (W, U, scale, zero, Q, loss):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.device == W.device
    assert colmap.dtype == torch.int32 and colmap.numel() == ncols and colmap.device == W.device
    rows, ldw = W.shape
    assert scale.shape == zero.shape and scale.shape[0] == rows and Q.shape == W.shape
    with _dev_ctx(W):
        E = torch.empty((rows, ncols), dtype=torch.float32, device=W.device)
        _lib.check(_lib.load().samq_gptq_block_fwd(
            _lib.ptr(W), ldw, rows, col0, ncols, _lib.ptr(U), U.shape[1], _lib.ptr(scale), _lib.ptr(zero),
            scale.shape[1], _lib.ptr(colmap), int(maxq), _lib.ptr(Q), Q.shape[1], _lib.ptr(E), _lib.ptr(loss),
            _lib.stream_ptr(W.device)))
    return E


def gather_cols(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """``x[..., perm]`` for fp16 ``x[..., K]`` and an int32 permutation of ``K`` entries
    (act-order layers on the fused GEMM path, see ``QuantLinear.sorted_pack``)."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    K = x.shape[-1]
    assert perm.dtype == torch.int32 and perm.numel() == K and perm.is_contiguous() and perm.device == x.device, \
        "perm must be a contiguous int32 vector of K entries on x's device"
    M = x.numel() // K
    with _dev_ctx(x):
        y = torch.empty_like(x)
        if M == 0:
            return y
        _lib.check(_lib.load().samq_gather_cols_fwd(_lib.ptr(x), _lib.ptr(perm), _lib.ptr(y), M, K,
                                                    _lib.stream_ptr(x.device)))
    return y


def qlinear(x: torch.Tensor, qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
            bits: int, groupsize: int, bias: Optional[torch.Tensor] = None,
            g_idx: Optional[torch.Tensor] = None, epilogue: int = _lib.EPI_NONE,
            residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            wt_ready: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``epi(x @ W + bias) + residual`` for a GPTQ-packed W; ``x[..., K]`` fp16 -> ``[..., N]`` fp16.
    ``wt_ready``: scratch that already holds this weight unpacked (``qlinear_prefetch``): only the GEMM runs."""
    if bits not in (2, 3, 4, 8):
        raise NotImplementedError("Only 2, 3, 4 and 8 bits are supported.")
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float16, f"A must be float16 (got {x.dtype})"
    assert x.is_contiguous(), "A must be contiguous"          # quant_linear.py:381
    K = x.shape[-1]
    f_words = qweight.shape[0] * 32 // bits
    assert K == f_words, "A's last dimension must match the packed weight's infeatures"  # :378-380
    N = qweight.shape[1]
    x2 = x.view(-1, K)
    M = x2.shape[0]
    _check_packed(x, qweight, qzeros, scales, g_idx, bias, bits, groupsize, K, N)
    with _dev_ctx(x):
        y = out if out is not None else torch.empty(x.shape[:-1] + (N,), dtype=torch.float16, device=x.device)
        assert y.is_contiguous() and y.dtype == torch.float16 and y.numel() == M * N
        if residual is not None:
            _check_half(residual, "residual")
            assert residual.numel() == M * N
        if M == 0:
            return y
        fused = g_idx is None and (K if groupsize == -1 else groupsize) % 64 == 0
        # scratch for the dequantised weight: always for the non-int4 formats; for int4 only when
        # M is long enough that the library prefers unpack-once + dense GEMM (see csrc/qlinear.cu)
        need_ws = (not fused) or M >= TWO_KERNEL_MIN_M or _lib.OPTIONS["gemm"] == "dense"
        if wt_ready is not None:
            _check_ready(wt_ready, x, K, N)
            _lib.check(_lib.load().samq_qlinear_fwd(
                _lib.ptr(x2), None, None, None, None, _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y),
                _lib.ptr(wt_ready), M, K, N, bits, groupsize, epilogue, _lib.stream_ptr(x.device)))
            return y
        ws = torch.empty(K * N, dtype=torch.float16, device=x.device) if need_ws else None
        _lib.check(_lib.load().samq_qlinear_fwd(
            _lib.ptr(x2), _lib.ptr(qweight), _lib.ptr(qzeros), _lib.ptr(scales), _lib.ptr(g_idx),
            _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), _lib.ptr(ws), M, K, N, bits, groupsize,
            epilogue, _lib.stream_ptr(x.device)))
    return y


def _check_ready(wt: torch.Tensor, x: torch.Tensor, K: int, N: int) -> None:
    assert wt.dtype == torch.float16 and wt.is_contiguous() and wt.numel() >= K * N and wt.device == x.device, \
        "prefetched weight scratch must be a contiguous fp16 buffer of >= K*N elements on x's device"


def qlinear_prefetch(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, bits: int,
                     groupsize: int, out: torch.Tensor) -> None:
    """Unpack an int4 weight (contiguous groups) into the scratch ``out`` NEXT TO the kernel enqueued
    before this call (samq_qlinear_prefetch); consume it with ``qlinear*(..., wt_ready=out)``.
    ``out`` must not be read by any kernel still in flight on the current stream."""
    if bits != 4:
        raise NotImplementedError("weight prefetch exists for int4 only")
    N = qweight.shape[1]
    K = qweight.shape[0] * 8
    _check_packed(qweight, qweight, qzeros, scales, None, None, bits, groupsize, K, N)
    _check_ready(out, qweight, K, N)
    with _dev_ctx(qweight):
        _lib.check(_lib.load().samq_qlinear_prefetch(
            _lib.ptr(qweight), _lib.ptr(qzeros), _lib.ptr(scales), _lib.ptr(out), K, N, bits, groupsize,
            _lib.stream_ptr(qweight.device)))


def qlinear_unpartition(x: torch.Tensor, qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                        bits: int, groupsize: int, bias: Optional[torch.Tensor], shortcut: torch.Tensor,
                        window_size: int, g_idx: Optional[torch.Tensor] = None,
                        wt_ready: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``shortcut + window_unpartition(x @ W + bias)``: ``x`` is ``[B*nWin, ws, ws, K]`` (windowed
    tokens), ``shortcut`` ``[B, H, W, N]``; returns ``[B, H, W, N]`` in image order."""
    if bits not in (2, 3, 4, 8):
        raise NotImplementedError("Only 2, 3, 4 and 8 bits are supported.")
    _lib.require_cuda(x, "x")
    _check_half(x, "x"); _check_half(shortcut, "shortcut")
    B, H, W, N = shortcut.shape
    K = x.shape[-1]
    ws = window_size
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    M = B * nH * nW * ws * ws
    assert x.numel() == M * K, "x must hold every windowed token of the batch"
    assert x.is_contiguous() and shortcut.is_contiguous()
    assert qweight.shape[1] == N and qweight.shape[0] * 32 // bits == K
    _check_packed(x, qweight, qzeros, scales, g_idx, bias, bits, groupsize, K, N)
    with _dev_ctx(x):
        y = torch.empty_like(shortcut)
        fused = g_idx is None and (K if groupsize == -1 else groupsize) % 64 == 0
        need_ws = (not fused) or M >= TWO_KERNEL_MIN_M or _lib.OPTIONS["gemm"] == "dense"
        wsp = torch.empty(K * N, dtype=torch.float16, device=x.device) if (need_ws and wt_ready is None) else wt_ready
        if wt_ready is not None:
            _check_ready(wt_ready, x, K, N)
            qweight = qzeros = scales = g_idx = None
        _lib.check(_lib.load().samq_qlinear_unpartition_fwd(
            _lib.ptr(x), _lib.ptr(qweight), _lib.ptr(qzeros), _lib.ptr(scales), _lib.ptr(g_idx), _lib.ptr(bias),
            _lib.ptr(shortcut), _lib.ptr(y), _lib.ptr(wsp), B, H, W, ws, K, N, bits, groupsize,
            _lib.stream_ptr(x.device)))
    return y


def qlinear_partition(x: torch.Tensor, qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor,
                      bits: int, groupsize: int, bias: Optional[torch.Tensor], window_size: int,
                      g_idx: Optional[torch.Tensor] = None, wt_ready: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``window_partition(x) @ W + bias`` without multiplying the zero-padding tokens: ``x`` is
    ``[B, H, W, K]`` (image order); returns ``[B*nWin, ws, ws, N]``, pad rows set to ``bias``."""
    if bits not in (2, 3, 4, 8):
        raise NotImplementedError("Only 2, 3, 4 and 8 bits are supported.")
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    assert x.dim() == 4 and x.is_contiguous()
    B, H, W, K = x.shape
    ws = window_size
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    N = qweight.shape[1]
    assert qweight.shape[0] * 32 // bits == K
    _check_packed(x, qweight, qzeros, scales, g_idx, bias, bits, groupsize, K, N)
    M = B * H * W
    with _dev_ctx(x):
        y = torch.empty((B * nH * nW, ws, ws, N), dtype=torch.float16, device=x.device)
        fused = g_idx is None and (K if groupsize == -1 else groupsize) % 64 == 0
        need_ws = (not fused) or M >= TWO_KERNEL_MIN_M or _lib.OPTIONS["gemm"] == "dense"
        wsp = torch.empty(K * N, dtype=torch.float16, device=x.device) if (need_ws and wt_ready is None) else wt_ready
        if wt_ready is not None:
            _check_ready(wt_ready, x, K, N)
            qweight = qzeros = scales = g_idx = None
        _lib.check(_lib.load().samq_qlinear_partition_fwd(
            _lib.ptr(x), _lib.ptr(qweight), _lib.ptr(qzeros), _lib.ptr(scales), _lib.ptr(g_idx), _lib.ptr(bias),
            _lib.ptr(y), _lib.ptr(wsp), B, H, W, ws, K, N, bits, groupsize, _lib.stream_ptr(x.device)))
    return y


def dense_linear(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 epilogue: int = _lib.EPI_NONE, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``epi(x @ wt.T + bias) + residual`` with ``wt[N, K]`` fp16 (already dequantised)."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    _check_half(wt, "wt")
    N, K = wt.shape
    assert x.shape[-1] == K
    x2 = x.view(-1, K)
    M = x2.shape[0]
    if bias is not None:
        _check_half(bias, "bias")
        assert bias.numel() == N and bias.device == x.device, "bias must be fp16[N] on x's device"
    if residual is not None:
        _check_half(residual, "residual")
        assert residual.numel() == M * N and residual.device == x.device, \
            f"residual must hold M*N = {M * N} fp16 values (got {residual.numel()})"
    assert wt.device == x.device
    with _dev_ctx(x):
        y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float16, device=x.device)
        if M == 0:
            return y
        _lib.check(_lib.load().samq_dense_linear_fwd(
            _lib.ptr(x2), _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), M, K, N,
            epilogue, _lib.stream_ptr(x.device)))
    return y


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    _lib.require_cuda(x, "x")
    _check_half(x, "x"); _check_half(weight, "weight"); _check_half(bias, "bias")
    C = x.shape[-1]
    rows = x.numel() // C
    with _dev_ctx(x):
        y = torch.empty_like(x)
        _lib.check(_lib.load().samq_layernorm_fwd(
            _lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y), rows, C, float(eps),
            _lib.stream_ptr(x.device)))
    return y


def layernorm_partition(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float,
                        window_size: int) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """LayerNorm + window_partition: ``[B,H,W,C]`` -> ``([B*nWin, ws, ws, C], (Hp, Wp))``."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x"); _check_half(weight, "weight"); _check_half(bias, "bias")
    B, H, W, C = x.shape
    ws = window_size
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    with _dev_ctx(x):
        y = torch.empty((B * nH * nW, ws, ws, C), dtype=torch.float16, device=x.device)
        _lib.check(_lib.load().samq_layernorm_partition_fwd(
            _lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y), B, H, W, C, ws, float(eps),
            _lib.stream_ptr(x.device)))
    return y, (nH * ws, nW * ws)


def unpartition_residual(windows: torch.Tensor, shortcut: torch.Tensor, window_size: int) -> torch.Tensor:
    """``shortcut + window_unpartition(windows)`` -> ``[B,H,W,C]``."""
    _lib.require_cuda(windows, "windows")
    _check_half(windows, "windows"); _check_half(shortcut, "shortcut")
    B, H, W, C = shortcut.shape
    with _dev_ctx(windows):
        out = torch.empty_like(shortcut)
        _lib.check(_lib.load().samq_unpartition_residual(
            _lib.ptr(windows), _lib.ptr(shortcut), _lib.ptr(out), B, H, W, C, window_size,
            _lib.stream_ptr(windows.device)))
    return out


def patchify(x: torch.Tensor, patch: int) -> torch.Tensor:
    """``[B, C, H, W]`` fp16 -> ``[B*(H/P)*(W/P), C*P*P]`` rows of non-overlapping patches."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    B, C, H, W = x.shape
    with _dev_ctx(x):
        out = torch.empty((B * (H // patch) * (W // patch), C * patch * patch), dtype=torch.float16, device=x.device)
        _lib.check(_lib.load().samq_patchify_fwd(_lib.ptr(x), _lib.ptr(out), B, C, H, W, patch,
                                                 _lib.stream_ptr(x.device)))
    return out


def im2col3x3(x: torch.Tensor) -> torch.Tensor:
    """NHWC ``[B, H, W, C]`` fp16 -> ``[B*H*W, 9*C]`` rows of zero-padded 3x3 neighbourhoods,
    column order (ky, kx, c)."""
    _lib.require_cuda(x, "x")
    _check_half(x, "x")
    if x.dim() != 4:
        raise ValueError("im2col3x3 expects a [B, H, W, C] tensor")
    B, H, W, C = x.shape
    with _dev_ctx(x):
        out = torch.empty((B * H * W, 9 * C), dtype=torch.float16, device=x.device)
        _lib.check(_lib.load().samq_im2col3x3_fwd(_lib.ptr(x), _lib.ptr(out), B, H, W, C,
                                                  _lib.stream_ptr(x.device)))
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(a, "a")
    _check_half(a, "a"); _check_half(b, "b")
    assert a.shape == b.shape
    with _dev_ctx(a):
        out = torch.empty_like(a)
        _lib.check(_lib.load().samq_add(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.numel(),
                                        _lib.stream_ptr(a.device)))
    return out


def attn_relpos(qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, B: int, H: int,
                W: int, num_heads: int, scale: float, relw_mode: int = _lib.RELW_REFERENCE) -> torch.Tensor:
    """Fused attention on the packed qkv GEMM output ``[B, H*W, 3*heads*hd]`` -> ``[B, H, W, heads*hd]``."""
    _lib.require_cuda(qkv, "qkv")
    _check_half(qkv, "qkv"); _check_half(rel_pos_h, "rel_pos_h"); _check_half(rel_pos_w, "rel_pos_w")
    D3 = qkv.shape[-1]
    assert D3 % (3 * num_heads) == 0
    hd = D3 // 3 // num_heads
    assert qkv.numel() == B * H * W * D3
    assert rel_pos_h.shape == (2 * H - 1, hd) and rel_pos_w.shape == (2 * W - 1, hd), \
        "rel_pos tables must be [2*size-1, head_dim] (no interpolation path)"
    with _dev_ctx(qkv):
        out = torch.empty((B, H, W, num_heads * hd), dtype=torch.float16, device=qkv.device)
        _lib.check(_lib.load().samq_attn_relpos_fwd(
            _lib.ptr(qkv), _lib.ptr(rel_pos_h), _lib.ptr(rel_pos_w), _lib.ptr(out), B, H, W,
            num_heads, hd, float(scale), relw_mode, _lib.stream_ptr(qkv.device)))
    return out


def attn_relpos_unpartition(qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, B: int, H: int,
                            W: int, window_size: int, num_heads: int, scale: float,
                            relw_mode: int = _lib.RELW_REFERENCE) -> torch.Tensor:
    """Windowed attention + window_unpartition: ``qkv[B*nWin, ws, ws, 3*heads*hd]`` (windowed tokens,
    zero-padding tokens included) -> ``[B, H, W, heads*hd]`` in image order."""
    _lib.require_cuda(qkv, "qkv")
    _check_half(qkv, "qkv"); _check_half(rel_pos_h, "rel_pos_h"); _check_half(rel_pos_w, "rel_pos_w")
    ws = window_size
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    D3 = qkv.shape[-1]
    assert D3 % (3 * num_heads) == 0
    hd = D3 // 3 // num_heads
    assert qkv.numel() == B * nH * nW * ws * ws * D3
    assert rel_pos_h.shape == (2 * ws - 1, hd) and rel_pos_w.shape == (2 * ws - 1, hd), \
        "rel_pos tables must be [2*window-1, head_dim] (no interpolation path)"
    with _dev_ctx(qkv):
        out = torch.empty((B, H, W, num_heads * hd), dtype=torch.float16, device=qkv.device)
        _lib.check(_lib.load().samq_attn_relpos_unpartition_fwd(
            _lib.ptr(qkv), _lib.ptr(rel_pos_h), _lib.ptr(rel_pos_w), _lib.ptr(out), B, H, W, ws,
            num_heads, hd, float(scale), relw_mode, _lib.stream_ptr(qkv.device)))
    return out
