"""SAM mask decoder + two-way transformer, state-dict compatible with the
reference's ``segment_anything/modeling/{mask_decoder,transformer}.py`` (same module tree:
``transformer.layers.<i>.{self_attn,cross_attn_token_to_image,cross_attn_image_to_token}.{q,k,v,out}_proj``,
``norm1..4``, ``mlp.lin1/lin2``, ``final_attn_token_to_image``, ``norm_final_attn``, ``iou_token``,
``mask_tokens``, ``output_upscaling.<0|1|3>``, ``output_hypernetworks_mlps.<i>.layers.<j>``,
``iou_prediction_head.layers.<j>``).

On a B200 with fp16 weights every op of the decoder runs on libsamq kernels (``_fused``):
  * linears over the 4096 image tokens (k / v / q projections, ``out_proj`` of image->token,
    both transposed convolutions as GEMMs)                -> tcgen05 dense GEMM (``ops.dense_linear``)
  * linears over the few prompt tokens, hypernetwork / IoU heads, mask product
                                                           -> ``ops.small_linear``
  * token<->image attention without positional bias        -> ``ops.attn_small``
  * LayerNorm / LayerNorm2d (NHWC rows)                     -> ``ops.layernorm``
  * ``x + pe`` and residual adds                            -> ``ops.add`` or a GEMM epilogue
Everything else (``cat`` / ``expand`` / ``permute``) is tensor plumbing.  On the CPU, or in fp32,
``forward`` runs the same graph with torch ops (checked against the reference's modules by
``tests/test_decoder_cpu.py``); the CUDA path never silently falls back to it.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .image_encoder import LayerNorm2d, MLPBlock

__all__ = ["MaskDecoder", "TwoWayTransformer", "TwoWayAttentionBlock", "DecoderAttention", "MLP"]


def _on_kernels(*tensors: torch.Tensor) -> bool:
    return all(t.is_cuda and t.dtype == torch.float16 for t in tensors)


def _linear(x: torch.Tensor, lin: nn.Linear, act: int = ops.ACT_NONE, residual=None) -> torch.Tensor:
    """nn.Linear on kernels: the tensor-core GEMM for long inputs whose shape tiles it, else the
    skinny kernel.  ``act``: ops.ACT_*; ``residual`` is added after the fp16 rounding."""
    N, K = lin.weight.shape
    rows = x.numel() // K
    x = x.contiguous()
    if rows >= 1024 and N % 128 == 0 and K % 64 == 0 and act in (ops.ACT_NONE, ops.ACT_GELU):
        from . import _lib

        return ops.dense_linear(x, lin.weight, lin.bias, _lib.EPI_GELU if act == ops.ACT_GELU else _lib.EPI_NONE,
                                None if residual is None else residual.contiguous())
    return ops.small_linear(x, lin.weight, lin.bias, act, None if residual is None else residual.contiguous())


def _ln(x: torch.Tensor, norm: nn.Module) -> torch.Tensor:
    return ops.layernorm(x.contiguous(), norm.weight, norm.bias, norm.eps)


class DecoderAttention(nn.Module):
    """Multi-head attention with an optional channel down-projection (transformer.py:185-240)."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, residual=None) -> torch.Tensor:
        """``out_proj(softmax(q k^T / sqrt(d)) v)`` (+ ``residual``); q ``[B, Nq, C]``, k / v ``[B, Nk, C]``."""
        if _on_kernels(q, k, v, self.q_proj.weight):
            o = ops.attn_small(_linear(q, self.q_proj), _linear(k, self.k_proj), _linear(v, self.v_proj), self.num_heads)
            return _linear(o, self.out_proj, residual=residual)
        B, Nq, _ = q.shape
        h, d = self.num_heads, self.internal_dim // self.num_heads
        qh = self.q_proj(q).reshape(B, Nq, h, d).transpose(1, 2)
        kh = self.k_proj(k).reshape(B, -1, h, d).transpose(1, 2)
        vh = self.v_proj(v).reshape(B, -1, h, d).transpose(1, 2)
        attn = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d), dim=-1)
        out = self.out_proj((attn @ vh).transpose(1, 2).reshape(B, Nq, h * d))
        return out if residual is None else residual + out


class TwoWayAttentionBlock(nn.Module):
    """Self-attention of the prompt tokens, tokens -> image cross attention, token MLP, image -> tokens
    cross attention, each followed by a post-LayerNorm (transformer.py:109-182)."""

    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048, activation=nn.ReLU,
                 attention_downsample_rate: int = 2, skip_first_layer_pe: bool = False):
        super().__init__()
        self.self_attn = DecoderAttention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = DecoderAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = DecoderAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        fused = _on_kernels(queries, keys, query_pe, key_pe, self.norm1.weight)
        add = ops.add if fused else torch.add
        norm = _ln if fused else (lambda x, n: n(x))
        if self.skip_first_layer_pe:
            queries = self.self_attn(queries, queries, queries)
        else:
            q = add(queries, query_pe)
            queries = self.self_attn(q, q, queries, residual=queries)
        queries = norm(queries, self.norm1)

        q = add(queries, query_pe)
        k = add(keys, key_pe)
        queries = norm(self.cross_attn_token_to_image(q, k, keys, residual=queries), self.norm2)

        if fused:
            act = ops.ACT_RELU if isinstance(self.mlp.act, nn.ReLU) else ops.ACT_GELU
            if not isinstance(self.mlp.act, (nn.ReLU, nn.GELU)):
                raise NotImplementedError("the decoder MLP kernel serves ReLU and GELU")
            hidden = _linear(queries, self.mlp.lin1, act)
            queries = _linear(hidden, self.mlp.lin2, residual=queries)
        else:
            queries = queries + self.mlp(queries)
        queries = norm(queries, self.norm3)

        q = add(queries, query_pe)
        keys = norm(self.cross_attn_image_to_token(k, q, queries, residual=keys), self.norm4)
        return queries, keys


class TwoWayTransformer(nn.Module):
    def __init__(self, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int, activation=nn.ReLU,
                 attention_downsample_rate: int = 2):
        super().__init__()
        self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = depth, embedding_dim, num_heads, mlp_dim
        self.layers = nn.ModuleList(
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, activation, attention_downsample_rate,
                                 skip_first_layer_pe=(i == 0)) for i in range(depth))
        self.final_attn_token_to_image = DecoderAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def forward(self, image_embedding, image_pe, point_embedding):
        """image_embedding / image_pe ``[B, C, h, w]``, point_embedding ``[B, N, C]`` ->
        (tokens ``[B, N, C]``, image tokens ``[B, h*w, C]``) (transformer.py:62-106)."""
        keys = image_embedding.flatten(2).permute(0, 2, 1).contiguous()
        key_pe = image_pe.flatten(2).permute(0, 2, 1).contiguous()
        queries = point_embedding.contiguous()
        for layer in self.layers:
            queries, keys = layer(queries, keys, point_embedding, key_pe)
        fused = _on_kernels(queries, keys, self.norm_final_attn.weight)
        add = ops.add if fused else torch.add
        q = add(queries, point_embedding)
        k = add(keys, key_pe)
        queries = self.final_attn_token_to_image(q, k, keys, residual=queries)
        queries = _ln(queries, self.norm_final_attn) if fused else self.norm_final_attn(queries)
        return queries, keys


class MLP(nn.Module):
    """ReLU MLP head (mask_decoder.py:155-178)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, sigmoid_output: bool = False):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.sigmoid_output = sigmoid_output

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fused = _on_kernels(x, self.layers[0].weight)
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            if fused:
                x = _linear(x, layer, ops.ACT_NONE if last else ops.ACT_RELU)
            else:
                x = layer(x) if last else F.relu(layer(x))
        return torch.sigmoid(x) if self.sigmoid_output else x


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation=nn.GELU, iou_head_depth: int = 3, iou_head_hidden_dim: int = 256):
        super().__init__()
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            activation(),
        )
        self.output_hypernetworks_mlps = nn.ModuleList(
            MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens))
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self._convt_cache: dict = {}

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (masks ``[B, 3 | 1, 4h, 4w]`` logits, IoU predictions ``[B, 3 | 1]``) (mask_decoder.py:71-108)."""
        masks, iou = self.predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings)
        pick = slice(1, None) if multimask_output else slice(0, 1)
        return masks[:, pick, :, :], iou[:, pick]

    # -- the two transposed convolutions as GEMMs over NHWC rows --------------------------------
    def _convt_matrix(self, conv: nn.ConvTranspose2d):
        """ConvTranspose2d(k=2, s=2) weight ``[Cin, Cout, 2, 2]`` -> GEMM weight ``[(dy, dx, o), Cin]`` and
        the bias repeated per (dy, dx); cached per weight version."""
        key = (id(conv), conv.weight.data_ptr(), conv.weight._version)
        hit = self._convt_cache.get(key)
        if hit is None:
            w = conv.weight.detach().permute(2, 3, 1, 0).reshape(-1, conv.in_channels).contiguous()
            b = conv.bias.detach().repeat(4).contiguous() if conv.bias is not None else None
            hit = self._convt_cache[key] = (w, b)
        return hit

    @staticmethod
    def _pixel_shuffle_rows(y: torch.Tensor, b: int, h: int, w: int, cout: int) -> torch.Tensor:
        """GEMM output rows ``[b*h*w, (dy, dx, o)]`` -> NHWC ``[b, 2h, 2w, o]``."""
        return y.view(b, h, w, 2, 2, cout).permute(0, 1, 3, 2, 4, 5).reshape(b, 2 * h, 2 * w, cout)

    def _upscale_fused(self, src_tokens: torch.Tensor, b: int, h: int, w: int) -> torch.Tensor:
        """output_upscaling on kernels: tokens ``[b, h*w, C]`` -> NHWC rows ``[b, 4h*4w, C/8]``."""
        from . import _lib

        c1, n1, a1, c2, a2 = self.output_upscaling
        for a in (a1, a2):
            if not isinstance(a, nn.GELU):
                raise NotImplementedError("the fused upscaling serves GELU")
        w1, b1 = self._convt_matrix(c1)
        y = ops.dense_linear(src_tokens.reshape(b * h * w, -1), w1, b1)
        y = self._pixel_shuffle_rows(y, b, h, w, c1.out_channels).contiguous()            # [b, 2h, 2w, C/4]
        y = ops.gelu(ops.layernorm(y.view(-1, c1.out_channels), n1.weight, n1.bias, n1.eps))
        w2, b2 = self._convt_matrix(c2)
        y = ops.dense_linear(y, w2, b2, _lib.EPI_GELU)                                    # GELU in the epilogue
        y = self._pixel_shuffle_rows(y, b, 2 * h, 2 * w, c2.out_channels).contiguous()    # [b, 4h, 4w, C/8]
        return y.view(b, 16 * h * w, c2.out_channels)

    def predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings):
        dt = image_embeddings.dtype
        sparse = sparse_prompt_embeddings.to(dt)
        dense = dense_prompt_embeddings.to(dt)
        out_tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0).to(dt)
        tokens = torch.cat((out_tokens.unsqueeze(0).expand(sparse.size(0), -1, -1), sparse), dim=1).contiguous()
        nb = tokens.shape[0]
        src = torch.repeat_interleave(image_embeddings, nb, dim=0)
        fused = _on_kernels(src, tokens, self.iou_token.weight, self.mask_tokens.weight)
        src = ops.add(src.contiguous(), dense.contiguous()) if fused else src + dense
        pos_src = torch.repeat_interleave(image_pe.to(dt), nb, dim=0)
        b, c, h, w = src.shape
        hs, src_tokens = self.transformer(src, pos_src, tokens)
        iou_token_out = hs[:, 0, :]
        mask_tokens_out = hs[:, 1:1 + self.num_mask_tokens, :]
        hyper: List[torch.Tensor] = [self.output_hypernetworks_mlps[i](mask_tokens_out[:, i, :].contiguous())
                                     for i in range(self.num_mask_tokens)]
        hyper_in = torch.stack(hyper, dim=1)                                              # [b, T, C/8]
        if fused:
            up = self._upscale_fused(src_tokens, b, h, w)                                 # [b, 16hw, C/8]
            masks = torch.stack([ops.small_linear(up[i], hyper_in[i].contiguous()) for i in range(b)])
            masks = masks.permute(0, 2, 1).reshape(b, -1, 4 * h, 4 * w)
        else:
            up = self.output_upscaling(src_tokens.transpose(1, 2).reshape(b, c, h, w))
            ub, uc, uh, uw = up.shape
            masks = (hyper_in @ up.view(ub, uc, uh * uw)).view(ub, -1, uh, uw)
        iou_pred = self.iou_prediction_head(iou_token_out.contiguous())
        return masks, iou_pred
