"""CPU oracle for the GPTQ packed-weight format and dequant arithmetic.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``sam_quantization_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it, as the checker.

Plain numpy restatement of the reference algorithm (no torch, no CUDA):

* ``find_params`` / ``fake_quantize``  -- ``Quantizer.find_params`` and ``quantize``
  (/root/reference/gptq.py:218-258, 183-187), asymmetric / symmetric min-max, per
  output row, applied per ``groupsize`` columns as the RTN path does
  (/root/reference/gptq4sam.py:61-70).
* ``pack``      -- ``pack_linear`` (/root/reference/gptq4sam.py:434-497): integer grid
  ``round((W + zero*scale)/scale)``, LSB-first fields, ``qzeros`` stores ``zero-1``.
  3-bit (extension, not supported by the reference QuantLinear) follows
  ``Quant3Linear.pack`` (/root/reference/quant.py:160-180): 32 values form a 96-bit
  little-endian bit stream over 3 consecutive int32 words.
* ``unpack_qweight`` / ``unpack_qzeros`` -- the shift/mask of ``matmul4_kernel``
  (/root/reference/gptq_triton/quant_linear.py:291-301, 312, 338).
* ``dequant``   -- ``b * scales - (zeros + 1) * scales``
  (/root/reference/gptq_triton/quant_linear.py:313, 334-339) in three rounding forms;
  ``fma`` -- ``fp16(q*s - fp16((z+1)*s))``, one fused multiply-add after the separately
  rounded zero term -- is THE definition the CUDA kernels are held to bit-for-bit: it is
  what the reference's Triton kernel computes on the B200 (Triton 3.6 contracts the
  multiply-subtract into ``fma.rn.f16x2``; 0 mismatches in 18 M elements of the
  identity-matrix extraction ``triton_matmul4(gs, I_K, ...)`` run by ``oracle/ref_gpu.py``,
  against 33 % for ``stepwise`` -- ``profiles/r02_ref_gpu.json``).  ``stepwise`` is what
  PyTorch eager gives for the literal expression; ``single`` rounds ``(q-z-1)*s`` once.
* ``qlinear``   -- ``x @ W + bias`` with fp32 accumulation
  (/root/reference/gptq_triton/quant_linear.py:341, 431-435).

Parity pin: ``tests/golden/make_golden.py`` runs the reference's own
``pack_linear`` (AST-extracted) and ``Quantizer`` in this container and stores their
outputs; ``tests/golden/dequant_triton_b4.npz`` holds the output of the reference's own
Triton kernel on a B200 (written by ``oracle/ref_gpu.py --sections dequant``);
``tests/test_oracle_quant.py`` checks this module against those fixtures bit-for-bit.  bits 2/3/8 in QuantLinear and
``g_idx`` are extensions: "parity unpinned by the reference" (it has no such path).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "find_params",
    "fake_quantize",
    "rtn_quantize",
    "pack",
    "unpack_qweight",
    "unpack_qzeros",
    "dequant",
    "qlinear",
    "default_g_idx",
]


# ---------------------------------------------------------------------------
# Quantizer (gptq.py:183-258)
# ---------------------------------------------------------------------------
def find_params(x: np.ndarray, bits: int, sym: bool = False):
    """Per-row min/max quantisation parameters of ``x[N, cols]`` (fp32).

    gptq.py:238-258 with perchannel=True, weight=True, mse=False.
    Returns (scale[N], zero[N]) as float32.
    """
    x = np.asarray(x, dtype=np.float32)
    maxq = np.float32(2**bits - 1)
    zero_row = np.zeros(x.shape[0], dtype=np.float32)
    xmin = np.minimum(x.min(axis=1), zero_row)
    xmax = np.maximum(x.max(axis=1), zero_row)
    if sym:
        xmax = np.maximum(np.abs(xmin), xmax)
        neg = xmin < 0
        xmin = np.where(neg, -xmax, xmin)
    dead = (xmin == 0) & (xmax == 0)
    xmin = np.where(dead, np.float32(-1), xmin).astype(np.float32)
    xmax = np.where(dead, np.float32(+1), xmax).astype(np.float32)
    scale = ((xmax - xmin) / maxq).astype(np.float32)
    if sym:
        zero = np.full_like(scale, (maxq + 1) / 2)
    else:
        zero = np.round(-xmin / scale).astype(np.float32)
    return scale, zero


def fake_quantize(x: np.ndarray, scale: np.ndarray, zero: np.ndarray, bits: int) -> np.ndarray:
    """``quantize`` of gptq.py:183-187: scale * (clamp(round(x/scale)+zero, 0, maxq) - zero)."""
    maxq = np.float32(2**bits - 1)
    x = np.asarray(x, dtype=np.float32)
    q = np.clip(np.round(x / scale) + zero, 0, maxq)
    return (scale * (q - zero)).astype(np.float32)


def rtn_quantize(weight: np.ndarray, bits: int, groupsize: int, sym: bool = False):
    """Round-to-nearest per group (the ``--nearest`` path, gptq4sam.py:61-70 applied per
    ``groupsize`` input columns).  ``weight`` is ``[N, K]`` (nn.Linear layout).

    Returns (W_fake[N,K] float32, scale[N,G] float32, zero[N,G] float32).
    """
    w = np.asarray(weight, dtype=np.float32)
    n, k = w.shape
    gs = k if groupsize == -1 else groupsize
    g = (k + gs - 1) // gs
    scale = np.empty((n, g), dtype=np.float32)
    zero = np.empty((n, g), dtype=np.float32)
    out = np.empty_like(w)
    for gi in range(g):
        sl = slice(gi * gs, min((gi + 1) * gs, k))
        s, z = find_params(w[:, sl], bits, sym)
        scale[:, gi] = s
        zero[:, gi] = z
        out[:, sl] = fake_quantize(w[:, sl], s[:, None], z[:, None], bits)
    return out, scale, zero


# ---------------------------------------------------------------------------
# bit-field helpers
# ---------------------------------------------------------------------------
def _pack_fields(vals: np.ndarray, bits: int) -> np.ndarray:
    """Pack ``vals[L, C]`` (uint, one field per row) along axis 0 into int32 words.

    2/4/8 bit: word r holds fields r*f .. r*f+f-1 at bit offsets bits*j
    (gptq4sam.py:472-477).  3 bit: 96-bit little-endian stream per 32 fields
    (quant.py:160-180).  Like the reference, fields are OR-ed in WITHOUT masking, so a
    negative value (zero-1 == -1) sign-fills the higher fields of its word
    (SURVEY trap 8); int32 wrap-around semantics are reproduced with uint64 math.
    """
    vals = np.asarray(vals)
    length, cols = vals.shape
    if bits in (2, 4, 8):
        f = 32 // bits
        assert length % f == 0, f"{length} fields do not fill int32 words of {f} fields"
        v = vals.astype(np.int64).reshape(length // f, f, cols)
        words = np.zeros((length // f, cols), dtype=np.int64)
        for j in range(f):
            # python-int shift on int64 keeps the sign fill of negative fields
            words |= v[:, j, :] << (bits * j)
        return (words & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    if bits == 3:
        assert length % 32 == 0, "3-bit packing needs a multiple of 32 fields"
        v = vals.astype(np.uint64).reshape(length // 32, 32, cols) & np.uint64(7)
        lo = np.zeros((length // 32, cols), dtype=np.uint64)  # stream bits 0..63
        hi = np.zeros((length // 32, cols), dtype=np.uint64)  # stream bits 64..95
        for j in range(32):
            p = 3 * j
            if p + 3 <= 64:
                lo |= v[:, j, :] << np.uint64(p)
            elif p >= 64:
                hi |= v[:, j, :] << np.uint64(p - 64)
            else:  # straddles bit 64 (j == 21: bits 63,64,65)
                lo |= (v[:, j, :] << np.uint64(p)) & np.uint64(0xFFFFFFFFFFFFFFFF)
                hi |= v[:, j, :] >> np.uint64(64 - p)
        words = np.empty((length // 32, 3, cols), dtype=np.uint32)
        words[:, 0, :] = (lo & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        words[:, 1, :] = (lo >> np.uint64(32)).astype(np.uint32)
        words[:, 2, :] = (hi & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        return words.reshape(length // 32 * 3, cols).view(np.int32)
    raise NotImplementedError("Only 2,3,4,8 bits are supported.")


def _unpack_fields(words: np.ndarray, bits: int, length: int) -> np.ndarray:
    """Inverse of ``_pack_fields`` along axis 0: returns uint8 ``[length, C]``."""
    w = np.ascontiguousarray(words).view(np.uint32)
    cols = w.shape[1]
    mask = np.uint32(2**bits - 1)
    if bits in (2, 4, 8):
        f = 32 // bits
        out = np.empty((w.shape[0], f, cols), dtype=np.uint8)
        for j in range(f):
            out[:, j, :] = ((w >> np.uint32(bits * j)) & mask).astype(np.uint8)
        return out.reshape(w.shape[0] * f, cols)[:length]
    if bits == 3:
        w3 = w.reshape(-1, 3, cols).astype(np.uint64)
        lo = w3[:, 0, :] | (w3[:, 1, :] << np.uint64(32))
        hi = w3[:, 2, :]
        out = np.empty((w3.shape[0], 32, cols), dtype=np.uint8)
        for j in range(32):
            p = 3 * j
            if p + 3 <= 64:
                v = lo >> np.uint64(p)
            elif p >= 64:
                v = hi >> np.uint64(p - 64)
            else:
                v = (lo >> np.uint64(p)) | (hi << np.uint64(64 - p))
            out[:, j, :] = (v & np.uint64(7)).astype(np.uint8)
        return out.reshape(-1, cols)[:length]
    raise NotImplementedError("Only 2,3,4,8 bits are supported.")


def default_g_idx(k: int, groupsize: int) -> np.ndarray:
    """Contiguous groups, ``g_idx[k] = k // groupsize`` (gptq4sam.py:457)."""
    gs = k if groupsize == -1 else groupsize
    return (np.arange(k) // gs).astype(np.int32)


# ---------------------------------------------------------------------------
# pack (gptq4sam.py:434-497)
# ---------------------------------------------------------------------------
def pack(weight, scale, zero, bits: int, groupsize: int, g_idx=None):
    """``pack_linear``: fake-quantised ``weight[N,K]`` + ``scale,zero[N,G]`` -> packed buffers.

    Returns dict(qweight int32[K*bits/32, N], qzeros int32[G, N*bits/32],
    scales fp16[G, N]).  ``weight`` may be fp16 or fp32: like the reference
    (fp16 model weights + fp32 quantiser outputs) the grid arithmetic runs in fp32.
    """
    w = np.asarray(weight)
    n, k = w.shape
    gs = k if groupsize == -1 else groupsize
    gi = default_g_idx(k, gs) if g_idx is None else np.asarray(g_idx, dtype=np.int64)
    scales_t = np.ascontiguousarray(np.asarray(scale, dtype=np.float32).T)  # [G, N]
    zeros_t = np.ascontiguousarray(np.asarray(zero, dtype=np.float32).T)    # [G, N]
    scale_zeros = zeros_t * scales_t
    # gptq4sam.py:462: round((W[:, idx] + scale_zeros[g]) / scales[g]) in fp32
    wf = w.astype(np.float32).T  # [K, N]
    intweight = np.round((wf + scale_zeros[gi]) / scales_t[gi]).astype(np.int32)  # [K, N]
    qweight = _pack_fields(intweight, bits)
    zeros_m1 = (zeros_t - 1).astype(np.int32)  # gptq4sam.py:482-485
    qzeros = np.ascontiguousarray(_pack_fields(np.ascontiguousarray(zeros_m1.T), bits).T)
    return {
        "qweight": np.ascontiguousarray(qweight),
        "qzeros": qzeros,
        "scales": scales_t.astype(np.float16),
    }


def unpack_qweight(qweight: np.ndarray, bits: int, k: int) -> np.ndarray:
    """q[k, n] (uint8) -- quant_linear.py:291-301, 338."""
    return _unpack_fields(qweight, bits, k)


def unpack_qzeros(qzeros: np.ndarray, bits: int, n: int) -> np.ndarray:
    """z[g, n] (uint8, the stored zero-1) -- quant_linear.py:312."""
    return np.ascontiguousarray(_unpack_fields(np.ascontiguousarray(qzeros.T), bits, n).T)


# ---------------------------------------------------------------------------
# dequant (quant_linear.py:313, 334-339)
# ---------------------------------------------------------------------------
def dequant(qweight, qzeros, scales, bits: int, groupsize: int, g_idx=None, form: str = "fma"):
    """Dequantised weight ``W[K, N]`` as fp16.

    form = "fma"     : fp16(q*s - fp16((z+1)*s))         <- the pinned definition (= the Triton kernel)
           "stepwise": fp16(fp16(q*s) - fp16((z+1)*s))   (PyTorch eager on the literal expression)
           "single"  : fp16((q - (z+1)) * s)
    """
    qweight = np.asarray(qweight)
    n = qweight.shape[1]
    k = qweight.shape[0] * 32 // bits
    gs = k if groupsize == -1 else groupsize
    gi = default_g_idx(k, gs) if g_idx is None else np.asarray(g_idx, dtype=np.int64)
    q = unpack_qweight(qweight, bits, k).astype(np.float16)           # exact (<= 255)
    z1 = (unpack_qzeros(qzeros, bits, n).astype(np.int32) + 1).astype(np.float16)  # exact (<= 256)
    s = np.asarray(scales).astype(np.float16)
    if form == "stepwise":
        zs = (z1 * s).astype(np.float16)                  # fp16 rounding
        return ((q * s[gi]).astype(np.float16) - zs[gi]).astype(np.float16)
    if form == "fma":
        zs = (z1 * s).astype(np.float16)
        exact = q.astype(np.float64) * s[gi].astype(np.float64) - zs[gi].astype(np.float64)
        return exact.astype(np.float16)
    if form == "single":
        exact = (q.astype(np.float64) - z1[gi].astype(np.float64)) * s[gi].astype(np.float64)
        return exact.astype(np.float16)
    raise ValueError(f"unknown dequant form {form!r}")


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """Exact-erf GELU (nn.GELU default, segment_anything/modeling/common.py:19,26)."""
    from math import erf, sqrt

    xv = np.asarray(x, dtype=np.float64)
    e = np.vectorize(erf)(xv / sqrt(2.0))
    return 0.5 * xv * (1.0 + e)


def qlinear(x, qweight, qzeros, scales, bits, groupsize, bias=None, g_idx=None,
            epilogue: str = "none", residual=None) -> np.ndarray:
    """fp32-accumulated ``x @ W + bias`` on the dequantised fp16 weight (pinned "fma" form)
    (quant_linear.py:341, 431-435); returns float32 (un-rounded) so tests can state
    their tolerance against the exact value."""
    w = dequant(qweight, qzeros, scales, bits, groupsize, g_idx).astype(np.float32)
    y = np.asarray(x, dtype=np.float32) @ w
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float32)
    if epilogue == "gelu":
        y = gelu_erf(y).astype(np.float32)
    if residual is not None:
        y = y + np.asarray(residual, dtype=np.float32)
    return y
