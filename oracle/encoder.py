"""CPU oracle for the SAM image-encoder hot path (fp32, plain torch on the CPU).

TEST INFRASTRUCTURE ONLY -- see oracle/quant.py.  Functional restatement (no nn.Module,
no CUDA) of:

* ``layer_norm``          nn.LayerNorm(eps=1e-6) (/root/reference/segment_anything/build_sam.py:72)
* ``window_partition`` / ``window_unpartition``  generic formulas
  (/root/reference/fq_vit/models/sam/image_encoder.py:481-537; the fork's
  image_encoder.py:282-333 hard-codes ViT-H batch 1 and equals them there)
* ``get_rel_pos``         /root/reference/segment_anything/modeling/image_encoder.py:336-366
* ``rel_pos_bias``        image_encoder.py:369-408 with the fork's bug-compatible rel_w
  (``torch.matmul(r_q, Rw.transpose(1, 2))`` broadcasts Rw over the query ROW,
  image_encoder.py:401-402, gptq_triton/fused_attention.py:76-78); ``upstream`` mode is
  Meta's einsum.  ``round_tables=True`` rounds rel_h/rel_w to fp16 as the reference's
  fp16 matmul does (fused_attention.py:76-78 run under model.half()).
* ``attention``           image_encoder.py:249-265 / the eager formula of ``test_op``
  (gptq_triton/fused_attention.py:388-406): fp32 softmax
* ``block``               image_encoder.py:189-207
* ``encoder``             image_encoder.py:106-118 (+ PatchEmbed :434-442, neck :88-104,
  LayerNorm2d common.py:38-43)

Parity pin: tests/golden/make_golden.py imports the reference's own ``ImageEncoderViT``
from /root/reference in this container and stores its outputs (ViT-H batch 1 unpatched,
small configs with the generic partition patched in); tests/test_oracle_encoder.py
checks this module against those fixtures.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import quant as oq

__all__ = [
    "layer_norm", "window_partition", "window_unpartition", "get_rel_pos", "rel_pos_bias",
    "attention_core", "attention", "block", "encoder", "dequant_state", "CONFIGS",
]

# /root/reference/segment_anything/build_sam.py:14-44
CONFIGS = {
    "vit_h": dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
}


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def window_partition(x: torch.Tensor, ws: int) -> Tuple[torch.Tensor, Tuple[int, int]]:
    B, H, W, C = x.shape
    pad_h = (ws - H % ws) % ws
    pad_w = (ws - W % ws) % ws
    if pad_h > 0 or pad_w > 0:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(windows: torch.Tensor, ws: int, pad_hw: Tuple[int, int], hw: Tuple[int, int]) -> torch.Tensor:
    Hp, Wp = pad_hw
    H, W = hw
    B = windows.shape[0] // (Hp * Wp // ws // ws)
    x = windows.view(B, Hp // ws, Wp // ws, ws, ws, -1)
    x = x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    return x[:, :H, :W, :].contiguous()


def get_rel_pos(size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """R[i, j] = rel_pos[i - j + size - 1] (square case, table length 2*size-1)."""
    assert rel_pos.shape[0] == 2 * size - 1
    coords = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[coords.long()]


def rel_pos_bias(q: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, hw: Tuple[int, int],
                 relw_mode: str = "reference", round_tables: bool = False):
    """q: [B', H*W, hd] (unscaled).  Returns rel_h [B',H,W,kH], rel_w [B',H,W,kW]."""
    H, W = hw
    Rh = get_rel_pos(H, rel_pos_h)
    Rw = get_rel_pos(W, rel_pos_w)
    r_q = q.reshape(q.shape[0], H, W, q.shape[-1])
    rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh)
    if relw_mode == "reference":
        rel_w = torch.einsum("bhwc,hkc->bhwk", r_q, Rw)   # == matmul(r_q, Rw.transpose(1, 2))
    elif relw_mode == "upstream":
        rel_w = torch.einsum("bhwc,wkc->bhwk", r_q, Rw)
    else:
        raise ValueError(relw_mode)
    if round_tables:
        rel_h = rel_h.half().float()
        rel_w = rel_w.half().float()
    return rel_h, rel_w


def attention_core(qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, B: int, H: int, W: int,
                   heads: int, scale: float, relw_mode: str = "reference", round_tables: bool = True) -> torch.Tensor:
    """softmax(scale q k^T + bias) v on the packed qkv ``[B, H*W, 3*heads*hd]`` -> ``[B,H,W,heads*hd]``."""
    S = H * W
    x = qkv.float().reshape(B, S, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = x.reshape(3, B * heads, S, -1).unbind(0)
    attn = (q * scale) @ k.transpose(-2, -1)
    rel_h, rel_w = rel_pos_bias(q, rel_pos_h.float(), rel_pos_w.float(), (H, W), relw_mode, round_tables)
    attn = (attn.view(B * heads, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(
        B * heads, S, S)
    attn = attn.softmax(dim=-1)
    return (attn @ v).view(B, heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)


def attention(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, heads: int,
              relw_mode: str = "reference", round_tables: bool = False) -> torch.Tensor:
    B, H, W, C = x.shape
    qkv = F.linear(x, p[prefix + "qkv.weight"], p.get(prefix + "qkv.bias"))
    scale = (C // heads) ** -0.5
    o = attention_core(qkv.reshape(B, H * W, -1), p[prefix + "rel_pos_h"], p[prefix + "rel_pos_w"], B, H, W,
                       heads, scale, relw_mode, round_tables)
    return F.linear(o, p[prefix + "proj.weight"], p.get(prefix + "proj.bias"))


def block(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, heads: int, window_size: int,
          relw_mode: str = "reference", eps: float = 1e-6) -> torch.Tensor:
    shortcut = x
    x = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], eps)
    H, W = x.shape[1], x.shape[2]
    if window_size > 0:
        x, pad_hw = window_partition(x, window_size)
    x = attention(x, p, prefix + "attn.", heads, relw_mode)
    if window_size > 0:
        x = window_unpartition(x, window_size, pad_hw, (H, W))
    x = shortcut + x
    h = layer_norm(x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], eps)
    h = F.linear(h, p[prefix + "mlp.lin1.weight"], p.get(prefix + "mlp.lin1.bias"))
    h = F.gelu(h)  # exact erf
    h = F.linear(h, p[prefix + "mlp.lin2.weight"], p.get(prefix + "mlp.lin2.bias"))
    return x + h


def _layer_norm_2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return x * w[:, None, None] + b[:, None, None]


def tokens_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], depth: int, heads: int, window_size: int,
                   global_attn_indexes, relw_mode: str = "reference") -> torch.Tensor:
    for i in range(depth):
        ws = 0 if i in global_attn_indexes else window_size
        x = block(x, p, f"blocks.{i}.", heads, ws, relw_mode)
    return x


def encoder(img: torch.Tensor, p: Dict[str, torch.Tensor], depth: int, num_heads: int, global_attn_indexes,
            window_size: int = 14, patch_size: int = 16, relw_mode: str = "reference", **_) -> torch.Tensor:
    """img [B,3,S,S] -> [B,256,S/16,S/16]; ``p`` holds fp32 tensors keyed like the
    reference state_dict (``patch_embed.proj.weight``, ``pos_embed``, ``blocks.i...``, ``neck.j...``)."""
    x = F.conv2d(img, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=patch_size)
    x = x.permute(0, 2, 3, 1)
    if "pos_embed" in p:
        x = x + p["pos_embed"]
    x = tokens_forward(x, p, depth, num_heads, window_size, global_attn_indexes, relw_mode)
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, p["neck.0.weight"])
    x = _layer_norm_2d(x, p["neck.1.weight"], p["neck.1.bias"])
    x = F.conv2d(x, p["neck.2.weight"], padding=1)
    return _layer_norm_2d(x, p["neck.3.weight"], p["neck.3.bias"])


def dequant_state(state: Dict[str, torch.Tensor], bits: int, groupsize: int) -> Dict[str, torch.Tensor]:
    """Packed state_dict (``...qweight/qzeros/scales[/g_idx][/bias]``) -> fp32 state_dict with
    ``...weight`` ``[N, K]`` from the oracle's stepwise dequant; every other tensor -> fp32."""
    out: Dict[str, torch.Tensor] = {}
    for key, val in state.items():
        if key.endswith(".qweight"):
            base = key[: -len("qweight")]
            g_idx = state.get(base + "g_idx")
            w = oq.dequant(val.cpu().numpy(), state[base + "qzeros"].cpu().numpy(),
                           state[base + "scales"].cpu().numpy(), bits, groupsize,
                           None if g_idx is None else g_idx.cpu().numpy())
            out[base + "weight"] = torch.from_numpy(np.ascontiguousarray(w.T.astype(np.float32)))
        elif key.endswith((".qzeros", ".scales", ".g_idx")):
            continue
        else:
            out[key] = val.detach().cpu().float()
    return out
