"""SAM prompt encoder, state-dict compatible with the reference's
``segment_anything/modeling/prompt_encoder.py`` (module / buffer names and shapes are the
checkpoint contract: ``pe_layer.positional_encoding_gaussian_matrix``, ``point_embeddings.<i>``,
``not_a_point_embed``, ``mask_downscaling.<0|1|3|4|6>``, ``no_mask_embed``).

What it computes (prompt_encoder.py:62-215): click / box-corner coordinates -> random Fourier
features (``sin | cos`` of ``2 pi (2 x - 1) G``) plus a learned embedding per prompt type; an
optional 256x256 mask prompt -> two stride-2 2x2 convolutions with LayerNorm2d + GELU and a 1x1
convolution down to the 64x64 embedding grid; without a mask the learned ``no_mask_embed``.

This is host-side torch code on a handful of points (a few hundred FLOP per prompt; the mask
branch has 1 / 4 / 16 channels): there is no kernel to write here.  Its outputs feed
``mask_decoder.MaskDecoder``, whose token/image work runs on libsamq kernels.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .image_encoder import LayerNorm2d

__all__ = ["PromptEncoder", "PositionEmbeddingRandom"]


class PositionEmbeddingRandom(nn.Module):
    """Random-frequency positional encoding (prompt_encoder.py:171-215)."""

    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None):
        super().__init__()
        scale = 1.0 if scale is None or scale <= 0.0 else scale
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def encode(self, unit_coords: torch.Tensor) -> torch.Tensor:
        """``[..., 2]`` coordinates in [0, 1] (x, y) -> ``[..., 2 * num_pos_feats]``, fp32."""
        proj = (2.0 * unit_coords.float() - 1.0) @ self.positional_encoding_gaussian_matrix.float()
        proj = 2.0 * math.pi * proj
        return torch.cat([proj.sin(), proj.cos()], dim=-1)

    def forward(self, size: Tuple[int, int]) -> torch.Tensor:
        """Dense encoding of an ``h x w`` grid of pixel centres -> ``[C, h, w]``."""
        h, w = size
        dev = self.positional_encoding_gaussian_matrix.device
        ys = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) / w
        grid = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)
        pe = self.encode(grid).to(self.positional_encoding_gaussian_matrix.dtype)
        return pe.permute(2, 0, 1)

    def forward_with_coords(self, coords: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
        """Pixel coordinates ``[B, N, 2]`` (x, y) of an image of ``image_size`` (h, w)."""
        unit = coords.to(torch.float32) / coords.new_tensor([image_size[1], image_size[0]], dtype=torch.float32)
        return self.encode(unit).to(coords.dtype)


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation=nn.GELU):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4                      # negative / positive click, two box corners
        self.point_embeddings = nn.ModuleList(nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings))
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans // 4),
            activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans),
            activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1),
        )
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self) -> torch.Tensor:
        """``[1, embed_dim, h, w]`` positional encoding of the image-embedding grid."""
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def _embed_points(self, points: torch.Tensor, labels: torch.Tensor, pad: bool) -> torch.Tensor:
        points = points + 0.5                              # pixel centre
        if pad:                                            # no box: one "not a point" slot (label -1)
            points = torch.cat([points, points.new_zeros((points.shape[0], 1, 2))], dim=1)
            labels = torch.cat([labels, -labels.new_ones((labels.shape[0], 1))], dim=1)
        emb = self.pe_layer.forward_with_coords(points, self.input_image_size)
        lab = labels.unsqueeze(-1)
        w = emb.dtype
        emb = torch.where(lab == -1, self.not_a_point_embed.weight.to(w).expand_as(emb), emb)
        emb = emb + (lab == 0).to(w) * self.point_embeddings[0].weight.to(w)
        emb = emb + (lab == 1).to(w) * self.point_embeddings[1].weight.to(w)
        return emb

    def _embed_boxes(self, boxes: torch.Tensor) -> torch.Tensor:
        corners = (boxes + 0.5).reshape(-1, 2, 2)
        emb = self.pe_layer.forward_with_coords(corners, self.input_image_size)
        offs = torch.stack([self.point_embeddings[2].weight[0], self.point_embeddings[3].weight[0]]).to(emb.dtype)
        return emb + offs

    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (sparse ``[B, N, embed_dim]``, dense ``[B, embed_dim, h, w]``) (prompt_encoder.py:128-168)."""
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        dev = self.point_embeddings[0].weight.device
        parts = [torch.empty((bs, 0, self.embed_dim), device=dev)]
        if points is not None:
            parts.append(self._embed_points(points[0], points[1], pad=boxes is None))
        if boxes is not None:
            parts.append(self._embed_boxes(boxes))
        sparse = torch.cat([p.to(parts[-1].dtype) for p in parts], dim=1)
        if masks is not None:
            dense = self.mask_downscaling(masks)
        else:
            dense = self.no_mask_embed.weight.to(sparse.dtype).reshape(1, -1, 1, 1).expand(
                bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense
