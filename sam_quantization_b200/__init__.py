"""sam_quantization_b200 -- B200-native GPTQ-quantized SAM image-encoder hot path.

Drop-in for the reference's ``gptq_triton`` package: same public names and call
signatures (/root/reference/gptq_triton/__init__.py:8-12), same packed-checkpoint
directory layout (``quant_config.json`` + ``model.safetensors`` | ``model.pt``,
gptq4sam.py:651-663), backed by hand-written sm_100a CUDA kernels reached through the
C ABI of ``include/samq.h``.  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import torch

from . import _lib, ops  # noqa: F401
from .fused_attention import QuantAttention, make_quant_attn
from .fused_mlp import QuantMLP, make_fused_mlp
from .gptq import GPTQ, Quantizer, encoder_pack, encoder_sequential
from .quant_linear import QuantLinear, make_quant, matmul4, triton_matmul4

__all__ = [
    "QuantLinear", "make_quant", "matmul4", "triton_matmul4", "QuantAttention", "make_quant_attn",
    "QuantMLP", "make_fused_mlp", "load_quant", "save_quant", "autotune_warmup", "ops",
    "GPTQ", "Quantizer", "encoder_sequential", "encoder_pack",
]


def _register_g_idx(model: torch.nn.Module, state: dict) -> None:
    """Checkpoints written with act-order carry ``<layer>.g_idx`` (extension); create the
    buffers so ``load_state_dict`` accepts them."""
    mods = dict(model.named_modules())
    for key, val in state.items():
        if key.endswith(".g_idx"):
            m = mods.get(key[: -len(".g_idx")])
            if isinstance(m, QuantLinear) and m.g_idx is None:
                m.g_idx = torch.empty_like(val, dtype=torch.int32)


def load_quant(model, checkpoint: str, warmup_autotune: bool = True, device: Optional[str] = "cuda",
               fuse_mlp: Optional[bool] = None, sub_module: Optional[str] = None,
               relw_mode: str = "reference"):
    """Load a packed checkpoint into ``model`` (reference: gptq_triton/__init__.py:15-81).

    Same steps and argument meaning as the reference: read ``quant_config.json``
    {wbits, groupsize}; ``make_quant`` on ``model`` (or ``model.<sub_module>``); load
    ``model.safetensors`` (strict) or ``model.pt`` (strict=False); drop all-zero biases
    (:57-60); ``make_quant_attn``; optionally fuse the MLPs; move to ``device``.
    ``fuse_mlp=None`` fuses (the reference's SwiGLU fusion was slower with groups, :66 --
    the GELU-epilogue fusion here is never slower).  ``warmup_autotune`` is accepted for
    compatibility; there is nothing to autotune.
    """
    ckpt = Path(checkpoint)
    with open(ckpt / "quant_config.json") as f:
        quant_config = json.load(f)
    wbits = quant_config["wbits"]
    groupsize = quant_config["groupsize"]
    model_quant = getattr(model, sub_module) if sub_module else model

    make_quant(model_quant, wbits, groupsize)

    if (ckpt / "model.safetensors").exists():
        from safetensors.torch import load_file as safe_load

        state = safe_load(str(ckpt / "model.safetensors"))
        _register_g_idx(model, state)
        model.load_state_dict(state)
    elif (ckpt / "model.pt").exists():
        state = torch.load(ckpt / "model.pt", map_location="cpu")
        _register_g_idx(model, state)
        model.load_state_dict(state, strict=False)
    else:
        raise FileNotFoundError(
            f"Could not find model checkpoint at {checkpoint}; please ensure that the path is "
            "correct and contains a `model.pt` or `model.safetensors` file.")

    for _, m in model.named_modules():
        if isinstance(m, QuantLinear) and m.bias is not None and bool((m.bias == 0).all()):
            m.bias = None

    make_quant_attn(model_quant, relw_mode=relw_mode)
    if fuse_mlp is None or fuse_mlp:
        make_fused_mlp(model_quant)

    if device is not None:
        model = model.to(device)
    if warmup_autotune and device is None:
        raise ValueError("You must specify a device when warmup_autotune is True.")
    return model


def save_quant(model, checkpoint: str, wbits: int, groupsize: int, safetensors: bool = False) -> None:
    """Write the reference's checkpoint directory (gptq4sam.py:651-663).  Must be called
    BEFORE ``make_quant_attn`` / ``make_fused_mlp`` so the keys are the reference's
    (``...attn.qkv.qweight`` etc.)."""
    ckpt = Path(checkpoint)
    ckpt.mkdir(parents=True, exist_ok=True)
    state = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
    if safetensors:
        from safetensors.torch import save_file

        save_file(state, str(ckpt / "model.safetensors"))
    else:
        torch.save(state, ckpt / "model.pt")
    with open(ckpt / "quant_config.json", "w") as f:
        json.dump({"wbits": wbits, "groupsize": groupsize}, f, indent=4)


def autotune_warmup(model) -> None:
    """No-op kept for API compatibility (reference: gptq_triton/__init__.py:84-104)."""
    return None
