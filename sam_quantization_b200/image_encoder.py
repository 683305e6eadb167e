"""Host module for the SAM ViT image encoder that the quantized operators drop into.

Mirrors the module tree, parameter names and forward semantics of
/root/reference/segment_anything/modeling/image_encoder.py:17-442 and
common.py:13-43 (so reference / upstream checkpoints load by key), written for the
fused B200 path:

  * ``window_partition`` / ``window_unpartition`` are the GENERIC formulas (the fork
    hard-codes ViT-H, batch 1: image_encoder.py:297-305, 324-332; generic form:
    fq_vit/models/sam/image_encoder.py:481-537) -- and on the fused path they are not
    separate ops at all: partition is fused into LayerNorm, unpartition + crop into the
    residual add.
  * ``Block.forward`` (image_encoder.py:189-207) has a fused CUDA path used whenever the
    block is quantized and fed fp16 CUDA tensors:
        LN1(+partition) -> qkv dequant-GEMM -> attention(+rel-pos) -> proj dequant-GEMM
        (+residual | unpartition+residual) -> LN2 -> lin1 dequant-GEMM(+bias+GELU)
        -> lin2 dequant-GEMM(+bias+residual)
    i.e. 6 kernels per block (+ the unpack of each weight at long M), no eager elementwise op.
  * Before quantization (fp32/fp16 nn.Linear) the module runs plain PyTorch, which is
    how weights are calibrated / packed and how the CPU tests drive the host logic.
    That eager path is NOT a fallback of the quantized path: a quantized block on a
    non-CUDA tensor raises.

Patch embedding and neck (0.26 % of FLOPs) run on the same kernels: patch / 3x3
re-layouts + the dense tcgen05 GEMM, LayerNorm2d on the token LayerNorm kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple, Type

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .fused_attention import QuantAttention
from .fused_mlp import QuantMLP
from .quant_linear import QuantLinear, WeightPrefetchChain

__all__ = [
    "ImageEncoderViT", "Block", "Attention", "MLPBlock", "PatchEmbed", "LayerNorm2d",
    "window_partition", "window_unpartition", "get_rel_pos", "decomposed_rel_pos",
    "ENCODER_CONFIGS", "build_image_encoder",
]

# build_sam.py:14-44 (shapes), :67-80 (shared settings)
ENCODER_CONFIGS = {
    "vit_h": dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
}


def build_image_encoder(name: str = "vit_h", **overrides) -> "ImageEncoderViT":
    """SAM image encoder with the registry's settings (build_sam.py:55-80): img 1024,
    patch 16, mlp_ratio 4, LayerNorm eps 1e-6, qkv_bias, rel-pos, window 14, out 256."""
    cfg = dict(ENCODER_CONFIGS[name])
    cfg.update(img_size=1024, patch_size=16, mlp_ratio=4.0, out_chans=256, qkv_bias=True,
               use_rel_pos=True, window_size=14, norm_eps=1e-6)
    cfg.update(overrides)
    return ImageEncoderViT(**cfg)


class LayerNorm2d(nn.Module):
    """Channel-wise LayerNorm over NCHW (common.py:31-43)."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mean = x.mean(1, keepdim=True)
        var = (x - mean).pow(2).mean(1, keepdim=True)
        x = (x - mean) / torch.sqrt(var + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class MLPBlock(nn.Module):
    """lin2(act(lin1(x))) (common.py:13-26)."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.lin2(self.act(self.lin1(x)))


class PatchEmbed(nn.Module):
    """conv16x16/16 then NCHW -> NHWC (image_encoder.py:411-442)."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans: int = 3,
                 embed_dim: int = 768) -> None:
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride,
                              padding=padding)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(x).permute(0, 2, 3, 1)


def window_partition(x: torch.Tensor, window_size: int) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """[B,H,W,C] -> ([B*nWin, ws, ws, C], (Hp, Wp)), zero padding bottom/right.
    Generic form (fq_vit/models/sam/image_encoder.py:481-507); equals the fork's
    hard-coded version for ViT-H, batch 1 (image_encoder.py:282-306)."""
    B, H, W, C = x.shape
    pad_h = (window_size - H % window_size) % window_size
    pad_w = (window_size - W % window_size) % window_size
    if pad_h or pad_w:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    x = x.view(B, Hp // window_size, window_size, Wp // window_size, window_size, C)
    windows = x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)
    return windows, (Hp, Wp)


def window_unpartition(windows: torch.Tensor, window_size: int, pad_hw: Tuple[int, int],
                       hw: Tuple[int, int]) -> torch.Tensor:
    """Inverse of ``window_partition`` + crop (fq_vit/.../image_encoder.py:510-537)."""
    Hp, Wp = pad_hw
    H, W = hw
    B = windows.shape[0] // (Hp * Wp // window_size // window_size)
    x = windows.view(B, Hp // window_size, Wp // window_size, window_size, window_size, -1)
    x = x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    if Hp > H or Wp > W:
        x = x[:, :H, :W, :].contiguous()
    return x


def resize_rel_pos(rel_pos: torch.Tensor, length: int) -> torch.Tensor:
    """Rel-pos table ``[L, C]`` linearly interpolated to ``length`` rows when ``L != length``
    (image_encoder.py:348-358: a checkpoint trained at another grid size); identity otherwise."""
    if rel_pos.shape[0] == length:
        return rel_pos
    t = F.interpolate(rel_pos.float().reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=length, mode="linear")
    return t.reshape(-1, length).permute(1, 0).to(rel_pos.dtype)


def get_rel_pos(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """R[i, j, :] = table[i' - j' + (k_size - 1) * max(q_size / k_size, 1)] with the coordinates of
    the shorter side stretched to the longer one and the table interpolated to
    2 * max(q_size, k_size) - 1 rows when its length differs (image_encoder.py:336-366).  SAM's own
    case is square with a table of exactly 2 * size - 1 rows: R[i, j] = rel_pos[i - j + size - 1]."""
    table = resize_rel_pos(rel_pos, int(2 * max(q_size, k_size) - 1))
    dev = rel_pos.device
    q_coords = torch.arange(q_size, device=dev)[:, None] * max(k_size / q_size, 1.0)
    k_coords = torch.arange(k_size, device=dev)[None, :] * max(q_size / k_size, 1.0)
    idx = (q_coords - k_coords) + (k_size - 1) * max(q_size / k_size, 1.0)
    return table[idx.long()]


def decomposed_rel_pos(q: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor,
                       hw: Tuple[int, int], relw_mode: str = "reference"):
    """(rel_h, rel_w), each [B', H, W, k] (image_encoder.py:369-408).  In "reference" mode
    rel_w uses the fork's matmul broadcasting (row-indexed Rw, image_encoder.py:401-402)."""
    H, W = hw
    Rh = get_rel_pos(H, H, rel_pos_h)
    Rw = get_rel_pos(W, W, rel_pos_w)
    r_q = q.reshape(q.shape[0], H, W, q.shape[-1])
    rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh)
    rel_w = torch.einsum("bhwc,hkc->bhwk" if relw_mode == "reference" else "bhwc,wkc->bhwk", r_q, Rw)
    return rel_h, rel_w


class Attention(nn.Module):
    """Eager multi-head attention with decomposed rel-pos (image_encoder.py:210-265);
    the pre-quantization form that ``make_quant_attn`` replaces."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, input_size: Optional[Tuple[int, int]] = None,
                 relw_mode: str = "reference") -> None:
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        self.relw_mode = relw_mode
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, _ = x.shape
        qkv = self.qkv(x).reshape(B, H * W, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.reshape(3, B * self.num_heads, H * W, -1).unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        if self.use_rel_pos:
            rel_h, rel_w = decomposed_rel_pos(q, self.rel_pos_h, self.rel_pos_w, (H, W), self.relw_mode)
            attn = (attn.view(-1, H, W, H, W) + rel_h[..., :, None] + rel_w[..., None, :]).view(-1, H * W, H * W)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).view(B, self.num_heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
        return self.proj(x)


class Block(nn.Module):
    """Transformer block with optional 14x14 window attention (image_encoder.py:139-207)."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU,
                 use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size

    # -- fused CUDA path -------------------------------------------------------
    def _fused_ready(self) -> bool:
        return (isinstance(self.attn, QuantAttention) and isinstance(self.attn.qkv_proj, QuantLinear)
                and isinstance(self.attn.o_proj, QuantLinear) and isinstance(self.mlp, QuantMLP)
                and isinstance(self.norm1, nn.LayerNorm) and isinstance(self.norm2, nn.LayerNorm))

    def _forward_fused(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        n1, n2 = self.norm1, self.norm2
        if self.window_size > 0:
            ws = self.window_size
            # proj epilogue does window_unpartition + crop + residual: shortcut + unpartition(attn)
            if ws == 14 and _lib.OPTIONS["pad_skip"]:   # "0": partition first, multiply the pad rows
                # qkv GEMM does the window_partition in its store, the attention kernel the
                # window_unpartition in its own: the zero-padding tokens (16 % of the window layout
                # at 64x64 / 14) are neither normalised nor multiplied by qkv / proj weights
                xn = ops.layernorm(x, n1.weight, n1.bias, n1.eps)
                x = self.attn(xn, residual=x, partition_window=ws)
            else:
                xw, _ = ops.layernorm_partition(x, n1.weight, n1.bias, n1.eps, ws)
                x = self.attn(xw, residual=x, unpartition_window=ws)
        else:
            xn = ops.layernorm(x, n1.weight, n1.bias, n1.eps)
            x = self.attn(xn, residual=x)                       # residual fused into proj
        xn = ops.layernorm(x, n2.weight, n2.bias, n2.eps)
        return self.mlp(xn, residual=x)                         # residual fused into lin2

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._fused_ready():
            if not (x.is_cuda and x.dtype == torch.float16):
                raise RuntimeError(
                    "quantized Block needs a float16 CUDA tensor: there is no CPU / eager "
                    f"fallback for the quantized path (got {x.dtype} on {x.device})")
            return self._forward_fused(x.contiguous())
        shortcut = x
        x = self.norm1(x)
        H, W = x.shape[1], x.shape[2]
        if self.window_size > 0:
            x, pad_hw = window_partition(x, self.window_size)
        x = self.attn(x)
        if self.window_size > 0:
            x = window_unpartition(x, self.window_size, pad_hw, (H, W))
        x = shortcut + x
        return x + self.mlp(self.norm2(x))


class ImageEncoderViT(nn.Module):
    """ViTDet-style SAM image encoder (image_encoder.py:17-118):
    ``forward(x[B,3,1024,1024]) -> [B,256,64,64]``."""

    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0,
                 out_chans: int = 256, qkv_bias: bool = True, norm_layer: Optional[Type[nn.Module]] = None,
                 act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, window_size: int = 0,
                 global_attn_indexes: Tuple[int, ...] = (), norm_eps: float = 1e-6) -> None:
        super().__init__()
        self.img_size = img_size
        if norm_layer is None:
            def norm_layer(dim):  # partial(nn.LayerNorm, eps=1e-6), build_sam.py:72
                return nn.LayerNorm(dim, eps=norm_eps)
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(
                torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(
                dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                norm_layer=norm_layer, act_layer=act_layer, use_rel_pos=use_rel_pos,
                rel_pos_zero_init=rel_pos_zero_init,
                window_size=window_size if i not in global_attn_indexes else 0,
                input_size=(img_size // patch_size, img_size // patch_size)))
        self._pos_cache = {}
        self._conv3_cache = {}
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False),
            LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            LayerNorm2d(out_chans),
        )

    def _prefetch_chain(self, x: torch.Tensor):
        """The quantized linears of the fused blocks in execution order, linked so that each one's
        GEMM is followed by the unpack of the next one's weight (``WeightPrefetchChain``)."""
        layers = []
        for blk in self.blocks:
            if blk._fused_ready():
                layers += [blk.attn.qkv_proj, blk.attn.o_proj, blk.mlp.lin1, blk.mlp.lin2]
        chain = self.__dict__.get("_pf_chain")
        if chain is None or chain.key != tuple(id(m) for m in layers):
            chain = WeightPrefetchChain(layers)
            self.__dict__["_pf_chain"] = chain
        chain.begin(x.device)
        return chain

    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """The 32-block hot loop on tokens ``[B, 64, 64, D]`` (image_encoder.py:111-113)."""
        if x.is_cuda and x.dtype == torch.float16:
            self._prefetch_chain(x)
        for blk in self.blocks:
            x = blk(x)
        return x

    # -- fused stem / neck -------------------------------------------------
    def _fused_stem_ready(self, x: torch.Tensor) -> bool:
        conv = self.patch_embed.proj
        return (x.is_cuda and x.dtype == torch.float16 and len(self.blocks) > 0
                and all(b._fused_ready() for b in self.blocks) and self.pos_embed is not None
                and conv.weight.dtype == torch.float16 and conv.kernel_size == conv.stride
                and conv.padding == (0, 0) and conv.kernel_size[0] == conv.kernel_size[1]
                and conv.kernel_size[0] % 8 == 0 and conv.out_channels % 128 == 0
                and (conv.in_channels * conv.kernel_size[0] ** 2) % 64 == 0 and conv.bias is not None
                and conv.bias.dtype == torch.float16 and self.pos_embed.dtype == torch.float16)

    def _stem_fused(self, x: torch.Tensor) -> torch.Tensor:
        """PatchEmbed + pos_embed as one patch re-layout + one tcgen05 GEMM (bias and the
        positional embedding ride in the epilogue) instead of cuDNN conv + permute + add."""
        conv = self.patch_embed.proj
        B = x.shape[0]
        P = conv.kernel_size[0]
        Hp, Wp = x.shape[2] // P, x.shape[3] // P
        if (Hp, Wp) != tuple(self.pos_embed.shape[1:3]) or x.shape[2] % P or x.shape[3] % P:
            # the eager path's `x + self.pos_embed` fails to broadcast here (image_encoder.py:108-109)
            raise RuntimeError(f"input of {tuple(x.shape[2:])} pixels gives {Hp}x{Wp} patches, but pos_embed is "
                               f"{tuple(self.pos_embed.shape[1:3])}")
        rows = ops.patchify(x.contiguous(), P)
        key = (B, self.pos_embed.data_ptr(), self.pos_embed._version, str(x.device))
        pos = self._pos_cache.get(key)
        if pos is None:     # batch-expanded positional embedding = the GEMM's residual operand
            pos = self.pos_embed.detach().expand(B, -1, -1, -1).contiguous()
            self._pos_cache = {key: pos}
        tok = ops.dense_linear(rows, conv.weight.view(conv.out_channels, -1), conv.bias, residual=pos)
        return tok.view(B, Hp, Wp, conv.out_channels)

    def _neck_fused(self, x: torch.Tensor) -> torch.Tensor:
        """conv1x1 as a tcgen05 GEMM, LayerNorm2d as the token LayerNorm kernel (channels are
        the last dim in NHWC), conv3x3 as 3x3-neighbourhood rows (``samq_im2col3x3_fwd``) times
        the weight laid out ``[O, (ky, kx, c)]`` on the same GEMM -- everything stays NHWC."""
        c1, n1, c3, n2 = self.neck[0], self.neck[1], self.neck[2], self.neck[3]
        B, H, W, D = x.shape
        y = ops.dense_linear(x.view(-1, D), c1.weight.view(c1.out_channels, D))
        y = ops.layernorm(y, n1.weight, n1.bias, n1.eps)
        C = y.shape[-1]
        if (c3.kernel_size == (3, 3) and c3.stride == (1, 1) and c3.padding == (1, 1) and c3.bias is None
                and c3.groups == 1 and C % 8 == 0 and (9 * C) % 64 == 0 and c3.out_channels % 256 == 0
                and not _lib.OPTIONS["neck_cudnn"]):
            key = (c3.weight.data_ptr(), c3.weight._version, str(c3.weight.device))
            w3 = self._conv3_cache.get(key)
            if w3 is None:      # [O, C, 3, 3] -> [O, (ky, kx, c)], once per weight version
                w3 = c3.weight.detach().permute(0, 2, 3, 1).reshape(c3.out_channels, 9 * C).contiguous()
                self._conv3_cache = {key: w3}
            y = ops.dense_linear(ops.im2col3x3(y.view(B, H, W, C)), w3)           # [B*H*W, O], NHWC
        else:                   # other conv shapes / the ablation switch: cuDNN in channels_last
            y = y.view(B, H, W, -1).permute(0, 3, 1, 2)
            y = F.conv2d(y, c3.weight.contiguous(memory_format=torch.channels_last), None, padding=c3.padding)
            y = y.permute(0, 2, 3, 1).contiguous().view(B * H * W, -1)
        y = ops.layernorm(y, n2.weight, n2.bias, n2.eps)
        return y.view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()     # NCHW like the reference

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._fused_stem_ready(x):
            x = self._stem_fused(x)
            x = self.forward_tokens(x)
            return self._neck_fused(x)
        x = self.patch_embed(x)
        if self.pos_embed is not None:
            x = x + self.pos_embed
        x = self.forward_tokens(x)
        return self.neck(x.permute(0, 3, 1, 2))
