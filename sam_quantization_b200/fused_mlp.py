"""Fused SAM MLP over quantized lin1 / lin2.

The reference's ``gptq_triton/fused_mlp.py`` is dead LLaMA SwiGLU code
(``silu(x W_gate) * (x W_up)``, fused_mlp.py:261,372-375; ``LlamaMLP`` undefined at :17,
import commented out at gptq_triton/__init__.py:11).  SAM's MLP is
``lin2(GELU_erf(lin1(x)))`` (segment_anything/modeling/common.py:21-26), so what is
built here is what that file's role is for SAM:

    lin1 dequant-GEMM with bias + exact-erf GELU in the epilogue (no [M, 4D] round trip
    for the activation)  ->  lin2 dequant-GEMM with bias (+ the block's residual add).

``make_fused_mlp`` keeps the reference's function name (fused_mlp.py:10-31).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .quant_linear import QuantLinear

__all__ = ["QuantMLP", "make_fused_mlp"]


class QuantMLP(nn.Module):
    def __init__(self, lin1: QuantLinear, lin2: QuantLinear):
        super().__init__()
        assert isinstance(lin1, QuantLinear) and isinstance(lin2, QuantLinear), \
            "QuantMLP needs quantized lin1/lin2 (run make_quant first)"
        self.lin1 = lin1
        self.lin2 = lin2

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        h = self.lin1(x, epilogue=_lib.EPI_GELU)
        return self.lin2(h, residual=residual)


def _is_mlp(m: nn.Module) -> bool:
    return (not isinstance(m, QuantMLP) and isinstance(getattr(m, "lin1", None), QuantLinear)
            and isinstance(getattr(m, "lin2", None), QuantLinear) and hasattr(m, "act"))


def make_fused_mlp(model: nn.Module) -> None:
    """Replace every SAM ``MLPBlock`` whose lin1/lin2 are QuantLinear by a ``QuantMLP``.
    Only exact-erf GELU is fused; other activations are left untouched."""
    for name, m in list(model.named_modules()):
        if not _is_mlp(m):
            continue
        act = m.act
        if not isinstance(act, nn.GELU) or getattr(act, "approximate", "none") != "none":
            continue
        fused = QuantMLP(m.lin1, m.lin2)
        if "." in name:
            parent_name, child = name.rsplit(".", 1)
            parent = model.get_submodule(parent_name)
        else:
            parent, child = model, name
        setattr(parent, child, fused)
