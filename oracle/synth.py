"""Deterministic synthetic SAM-encoder weights for tests (TEST INFRASTRUCTURE ONLY).

numpy ``default_rng`` streams (stable across numpy versions and machines), reference
state_dict key names (/root/reference/segment_anything/modeling/image_encoder.py), and the
reference's own quantise-then-pack route: RTN per 128-column group with the
``Quantizer`` rules (gptq.py:218-258) -> ``pack_linear`` layout (gptq4sam.py:434-497),
both via oracle/quant.py.  ``pos_embed`` and ``rel_pos_{h,w}`` are randomised because the
reference zero-initialises them (image_encoder.py:68-70, 246-247), which would hide the
rel-pos path (SURVEY trap 5).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import quant as oq


def fp_state(embed_dim: int, depth: int, num_heads: int, global_attn_indexes, window_size: int = 14,
             tokens: int = 64, patch_size: int = 16, out_chans: int = 256, mlp_ratio: float = 4.0,
             seed: int = 0, with_stem: bool = True) -> Dict[str, np.ndarray]:
    """fp32 weights keyed like the reference's ``ImageEncoderViT.state_dict()``."""
    rng = np.random.default_rng(seed)
    D = embed_dim
    hd = D // num_heads
    mlp = int(D * mlp_ratio)

    def lin(n, k):
        return (rng.standard_normal((n, k)) * 0.02).astype(np.float32), (rng.standard_normal(n) * 0.02).astype(np.float32)

    p: Dict[str, np.ndarray] = {}
    if with_stem:
        p["patch_embed.proj.weight"] = (rng.standard_normal((D, 3, patch_size, patch_size)) * 0.02).astype(np.float32)
        p["patch_embed.proj.bias"] = (rng.standard_normal(D) * 0.02).astype(np.float32)
        p["pos_embed"] = (rng.standard_normal((1, tokens, tokens, D)) * 0.02).astype(np.float32)
    for i in range(depth):
        pre = f"blocks.{i}."
        size = tokens if i in global_attn_indexes else window_size
        p[pre + "norm1.weight"] = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        p[pre + "norm1.bias"] = (0.1 * rng.standard_normal(D)).astype(np.float32)
        p[pre + "attn.qkv.weight"], p[pre + "attn.qkv.bias"] = lin(3 * D, D)
        p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"] = lin(D, D)
        p[pre + "attn.rel_pos_h"] = (rng.standard_normal((2 * size - 1, hd)) * 0.02).astype(np.float32)
        p[pre + "attn.rel_pos_w"] = (rng.standard_normal((2 * size - 1, hd)) * 0.02).astype(np.float32)
        p[pre + "norm2.weight"] = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        p[pre + "norm2.bias"] = (0.1 * rng.standard_normal(D)).astype(np.float32)
        p[pre + "mlp.lin1.weight"], p[pre + "mlp.lin1.bias"] = lin(mlp, D)
        p[pre + "mlp.lin2.weight"], p[pre + "mlp.lin2.bias"] = lin(D, mlp)
    if with_stem:
        p["neck.0.weight"] = (rng.standard_normal((out_chans, D, 1, 1)) * 0.02).astype(np.float32)
        p["neck.1.weight"] = np.ones(out_chans, dtype=np.float32)
        p["neck.1.bias"] = np.zeros(out_chans, dtype=np.float32)
        p["neck.2.weight"] = (rng.standard_normal((out_chans, out_chans, 3, 3)) * 0.02).astype(np.float32)
        p["neck.3.weight"] = np.ones(out_chans, dtype=np.float32)
        p["neck.3.bias"] = np.zeros(out_chans, dtype=np.float32)
    return p


_LINEARS = ("attn.qkv", "attn.proj", "mlp.lin1", "mlp.lin2")


def quantize_state(p: Dict[str, np.ndarray], bits: int, groupsize: int, act_order_seed: Optional[int] = None
                   ) -> Dict[str, np.ndarray]:
    """Replace every block Linear's ``weight`` by packed ``qweight/qzeros/scales`` (fp16
    weights in, like the reference's model.half() before packing, gptq4sam.py:635) and
    cast everything else to fp16.  ``act_order_seed``: also emit a random-permutation
    ``g_idx`` (invperm // groupsize) -- the extension of BASELINE config 4."""
    out: Dict[str, np.ndarray] = {}
    rng = None if act_order_seed is None else np.random.default_rng(act_order_seed)
    for key, val in p.items():
        if key.endswith(".weight") and any(key.endswith(l + ".weight") for l in _LINEARS):
            base = key[: -len("weight")]
            w16 = val.astype(np.float16)
            n, k = w16.shape
            gs = k if groupsize == -1 else groupsize
            g_idx = None
            if rng is not None:
                perm = rng.permutation(k)
                inv = np.empty(k, dtype=np.int64)
                inv[perm] = np.arange(k)
                g_idx = (inv // gs).astype(np.int32)
                # quantise in permuted order so each group's params come from its members
                wp = w16.astype(np.float32)[:, perm]
                wfp, scale, zero = oq.rtn_quantize(wp, bits, gs)
                wf = np.empty_like(wfp)
                wf[:, perm] = wfp
            else:
                wf, scale, zero = oq.rtn_quantize(w16.astype(np.float32), bits, gs)
            packed = oq.pack(wf.astype(np.float16), scale, zero, bits, gs, g_idx)
            out[base + "qweight"] = packed["qweight"]
            out[base + "qzeros"] = packed["qzeros"]
            out[base + "scales"] = packed["scales"]
            if g_idx is not None:
                out[base + "g_idx"] = g_idx
        else:
            out[key] = val.astype(np.float16)
    return out


def to_torch(p: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}


def image(batch: int, size: int = 1024, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(1000 + seed).standard_normal((batch, 3, size, size)).astype(np.float32)


def tokens_input(batch: int, tokens: int, dim: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(2000 + seed).standard_normal((batch, tokens, tokens, dim)).astype(np.float32)
