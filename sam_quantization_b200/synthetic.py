"""Random-init packed encoders for benchmarking (no dataset / checkpoint is available
offline).  Weights are drawn directly in the packed domain -- uniform random int4 fields
and zero points, scales ~ U(0.003, 0.006) so that the dequantised weights have the
std ~ 0.02 of a freshly initialised ViT -- through the product's own modules."""
from __future__ import annotations

import torch
import torch.nn as nn

from .fused_attention import make_quant_attn
from .fused_mlp import make_fused_mlp
from .image_encoder import ImageEncoderViT, build_image_encoder
from .quant_linear import QuantLinear, make_quant

__all__ = ["randomize_packed_", "random_quantized_encoder"]


@torch.no_grad()
def randomize_packed_(model: nn.Module, seed: int = 0, scale_lo: float = 0.003, scale_hi: float = 0.006,
                      act_order: bool = False) -> None:
    """``act_order``: also give every layer a permutation-derived ``g_idx`` (``invperm // groupsize``,
    what GPTQ's act-order produces: non-contiguous groups of exactly ``groupsize`` members)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, QuantLinear):
            if act_order:
                perm = torch.randperm(m.infeatures, generator=g)
                inv = torch.empty_like(perm)
                inv[perm] = torch.arange(m.infeatures)
                m.g_idx = (inv // m.groupsize).to(torch.int32)
            m.qweight.copy_(torch.randint(-2**31, 2**31 - 1, m.qweight.shape, generator=g, dtype=torch.int64).to(torch.int32))
            m.qzeros.copy_(torch.randint(-2**31, 2**31 - 1, m.qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32))
            m.scales.copy_((torch.rand(m.scales.shape, generator=g) * (scale_hi - scale_lo) + scale_lo).half())
            if m.bias is not None:
                m.bias.copy_((torch.randn(m.bias.shape, generator=g) * 0.02).half())
    for name, p in model.named_parameters():
        if "rel_pos" in name or name.endswith("pos_embed"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)   # zero-init in the reference (trap 5)


@torch.no_grad()
def _materialize_meta_(model: nn.Module, seed: int) -> None:
    """Give real storage + a ViT-style init to every parameter still on the meta device
    (LayerNorm -> 1/0, everything else ~ N(0, 0.02))."""
    g = torch.Generator().manual_seed(seed + 1)
    for mod in model.modules():
        for pname, p in list(mod.named_parameters(recurse=False)):
            if not p.is_meta:
                continue
            if isinstance(mod, nn.LayerNorm) or type(mod).__name__ == "LayerNorm2d":
                val = torch.ones(p.shape) if pname == "weight" else torch.zeros(p.shape)
            else:
                val = torch.randn(p.shape, generator=g) * 0.02
            setattr(mod, pname, nn.Parameter(val, requires_grad=False))


def random_quantized_encoder(name: str = "vit_h", bits: int = 4, groupsize: int = 128, seed: int = 0,
                             device: str = "cuda", relw_mode: str = "reference", act_order: bool = False,
                             **overrides) -> ImageEncoderViT:
    torch.manual_seed(seed)
    with torch.device("meta"):          # skip the 2.5 GB fp32 init of Linear weights that get replaced
        enc = build_image_encoder(name, **overrides)
    make_quant(enc, bits, groupsize)
    _materialize_meta_(enc, seed)
    randomize_packed_(enc, seed, act_order=act_order)
    enc = enc.half()
    make_quant_attn(enc, relw_mode=relw_mode)
    make_fused_mlp(enc)
    return enc.to(device).eval()
