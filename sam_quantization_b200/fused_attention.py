"""QuantAttention -- drop-in for ``gptq_triton.fused_attention`` on B200.

Constructor and ``forward(x[B, H, W, C])`` follow the reference
(/root/reference/gptq_triton/fused_attention.py:83-149).  The reference forward runs
qkv GEMM -> permuted copy to slice q (:118) -> ``add_decomposed_rel_pos`` (two
``get_rel_pos`` gathers + two batched matmuls writing ``[B*heads, H, W, H]`` tables to
HBM, :46-80) -> Triton flash kernel reading those tables (:159-358) -> proj GEMM.
Here the middle three steps are ONE CUDA kernel (``samq_attn_relpos_fwd``): the
rel-pos products are tcgen05 MMAs in the kernel prologue and the bias add happens
inside the softmax; nothing but qkv and the attention output touches HBM.

``relw_mode`` (additive): "reference" (default) reproduces the fork's rel_w semantics
(``torch.matmul(q, Rw.transpose(1, 2))`` broadcasts Rw's leading axis against the query
ROW index, fused_attention.py:76-78); "upstream" is Meta's
``einsum("bhwc,wkc->bhwk")``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops

__all__ = ["QuantAttention", "make_quant_attn"]

_RELW = {"reference": _lib.RELW_REFERENCE, "upstream": _lib.RELW_UPSTREAM}


class QuantAttention(nn.Module):
    """Multi-head attention with decomposed rel-pos over quantized qkv / proj."""

    def __init__(self, qkv_proj, o_proj, num_heads, scale, use_rel_pos, rel_pos_h=None,
                 rel_pos_w=None, relw_mode: str = "reference"):
        super().__init__()
        if relw_mode not in _RELW:
            raise ValueError(f"relw_mode must be one of {sorted(_RELW)}")
        self.qkv_proj = qkv_proj
        self.o_proj = o_proj
        self.num_heads = num_heads
        self.scale = scale
        self.rel_pos_h = rel_pos_h
        self.rel_pos_w = rel_pos_w
        self.use_rel_pos = use_rel_pos
        self.relw_mode = relw_mode

    def _tables(self, kh: int, kw: int):
        """fp16 rel-pos tables of exactly 2*kh-1 / 2*kw-1 rows for the kernels: the parameters
        themselves in SAM's own case, linearly interpolated (image_encoder.py:348-358) when the
        checkpoint was trained at another grid size; cached per parameter version."""
        if not self.use_rel_pos:
            raise NotImplementedError          # fused_attention.py:134-135
        key = (kh, kw, self.rel_pos_h.data_ptr(), self.rel_pos_h._version, self.rel_pos_w.data_ptr(),
               self.rel_pos_w._version, self.rel_pos_h.dtype)
        hit = getattr(self, "_table_cache", None)
        if hit is None or hit[0] != key:
            from .image_encoder import resize_rel_pos

            rph = resize_rel_pos(self.rel_pos_h.detach(), 2 * kh - 1).half().contiguous()
            rpw = resize_rel_pos(self.rel_pos_w.detach(), 2 * kw - 1).half().contiguous()
            hit = self._table_cache = (key, rph, rpw)
        return hit[1], hit[2]

    def attention(self, qkv: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        """softmax(scale q k^T + rel-pos bias) v on the packed qkv GEMM output."""
        rph, rpw = self._tables(H, W)
        return ops.attn_relpos(qkv, rph, rpw, B, H, W, self.num_heads, self.scale, _RELW[self.relw_mode])

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None,
                unpartition_window: int = 0, partition_window: int = 0) -> torch.Tensor:
        """Input ``[B, H, W, C]`` -> ``[B, H, W, C]`` (fused_attention.py:107-149).
        Additive: ``residual`` is fused into the proj GEMM epilogue; with
        ``unpartition_window = ws`` the input is windowed tokens ``[B*nWin, ws, ws, C]``, and the
        proj epilogue also performs window_unpartition + crop, returning
        ``residual + unpartition(attn)`` in the residual's ``[B, H, W, C]`` image order.
        With ``partition_window = ws`` the input is the UN-partitioned ``[B, H, W, C]``: the qkv GEMM
        does the window_partition in its store and the attention kernel the window_unpartition in
        its own, so the zero-padding tokens of the window layout are never multiplied by a weight
        (as keys they still take part, with q = k = v = bias, exactly like the reference)."""
        if partition_window:
            ws = partition_window
            B, H, W, _ = x.shape
            rph, rpw = self._tables(ws, ws)
            qkv = self.qkv_proj.forward_partition(x, ws)
            o = ops.attn_relpos_unpartition(qkv, rph, rpw, B, H, W, ws, self.num_heads,
                                            self.scale, _RELW[self.relw_mode])
            return self.o_proj(o, residual=residual) if residual is not None else self.o_proj(o)
        B, H, W, _ = x.shape
        qkv = self.qkv_proj(x)
        o = self.attention(qkv, B, H, W)
        if unpartition_window:
            return self.o_proj.forward_unpartition(o, residual, unpartition_window)
        if residual is not None:
            return self.o_proj(o, residual=residual)
        return self.o_proj(o)


def _is_attention(m: nn.Module) -> bool:
    # duck-typed match of segment_anything's Attention (image_encoder.py:210-265) so the
    # swap works on the reference's, upstream's and this package's encoder classes
    return (not isinstance(m, QuantAttention) and hasattr(m, "qkv") and hasattr(m, "proj")
            and hasattr(m, "num_heads") and hasattr(m, "scale") and hasattr(m, "use_rel_pos"))


def make_quant_attn(model: nn.Module, relw_mode: str = "reference") -> None:
    """Replace every SAM ``Attention`` by a ``QuantAttention`` sharing its (quantized)
    qkv / proj layers and rel-pos tables (fused_attention.py:12-43)."""
    for name, m in list(model.named_modules()):
        if not _is_attention(m):
            continue
        attn = QuantAttention(
            m.qkv, m.proj, m.num_heads, m.scale, m.use_rel_pos,
            m.rel_pos_h if m.use_rel_pos else None,
            m.rel_pos_w if m.use_rel_pos else None,
            relw_mode=relw_mode,
        )
        if "." in name:
            parent_name, child = name.rsplit(".", 1)
            parent = model.get_submodule(parent_name)
        else:
            parent, child = model, name
        setattr(parent, child, attn)
