"""Whole-model wrapper and the interactive-segmentation evaluation loop.

``Sam`` holds the three sub-modules under the reference's names (``image_encoder``,
``prompt_encoder``, ``mask_decoder``; segment_anything/modeling/sam.py:18-50, build_sam.py:47-105)
so that a reference checkpoint's keys load unchanged; ``build_sam`` builds the ViT-B / L / H
variants; ``interactive_eval`` is the reference's 5-click mIoU protocol
(script/evaluation2.py:225-334: one random click inside the current error region per round --
positive on a missed pixel, negative on a false positive --, all clicks so far plus the previous
low-resolution mask as prompts, bilinear upsampling, IoU against the ground truth).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .image_encoder import ImageEncoderViT, build_image_encoder
from .mask_decoder import MaskDecoder, TwoWayTransformer
from .prompt_encoder import PromptEncoder

__all__ = ["Sam", "build_sam", "get_iou", "next_clicks", "interactive_eval"]


class Sam(nn.Module):
    mask_threshold: float = 0.0

    def __init__(self, image_encoder: ImageEncoderViT, prompt_encoder: PromptEncoder, mask_decoder: MaskDecoder,
                 pixel_mean: Sequence[float] = (123.675, 116.28, 103.53),
                 pixel_std: Sequence[float] = (58.395, 57.12, 57.375)):
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), persistent=False)

    @property
    def device(self):
        return self.pixel_mean.device

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        """Normalise the pixel values and zero-pad bottom / right to the encoder's square input (sam.py:164-174)."""
        x = (x - self.pixel_mean) / self.pixel_std
        size = self.image_encoder.img_size
        h, w = x.shape[-2:]
        return F.pad(x, (0, size - w, 0, size - h))

    def postprocess_masks(self, masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...]
                          ) -> torch.Tensor:
        """Low-resolution logits -> the original image frame: upsample to the encoder input, cut the
        padding off, resize to ``original_size`` (sam.py:133-162)."""
        size = self.image_encoder.img_size
        masks = F.interpolate(masks, (size, size), mode="bilinear", align_corners=False)
        masks = masks[..., : input_size[0], : input_size[1]]
        return F.interpolate(masks, original_size, mode="bilinear", align_corners=False)

    @torch.no_grad()
    def forward(self, batched_input: List[Dict[str, object]], multimask_output: bool) -> List[Dict[str, torch.Tensor]]:
        """End to end for a list of images with their prompts (sam.py:53-131).  Every record holds
        ``image`` (3 x H x W, already resized to the model's frame), ``original_size`` and any of
        ``point_coords`` + ``point_labels``, ``boxes``, ``mask_inputs``; returns per record ``masks``
        (bool, original frame), ``iou_predictions`` and ``low_res_logits``.  All images go through the
        encoder as ONE batch (the fused CUDA path when the encoder is a loaded quantised one)."""
        dt = next(self.image_encoder.parameters()).dtype
        images = torch.stack([self.preprocess(rec["image"]) for rec in batched_input], dim=0).to(dt)
        embeddings = self.image_encoder(images)
        outputs = []
        for rec, emb in zip(batched_input, embeddings):
            points = (rec["point_coords"], rec["point_labels"]) if "point_coords" in rec else None
            low_res, iou = self.predict_masks(emb.unsqueeze(0), points=points, boxes=rec.get("boxes"),
                                              mask_input=rec.get("mask_inputs"), multimask_output=multimask_output)
            masks = self.postprocess_masks(low_res.float(), input_size=rec["image"].shape[-2:],
                                           original_size=rec["original_size"])
            outputs.append({"masks": masks > self.mask_threshold, "iou_predictions": iou, "low_res_logits": low_res})
        return outputs

    @torch.no_grad()
    def predict_masks(self, image_embedding: torch.Tensor, points: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Prompts -> (low-resolution mask logits ``[B, 1 | 3, 256, 256]``, IoU predictions), for the
        embedding(s) of ONE image or one embedding per prompt row (evaluation2.py:290-303)."""
        sparse, dense = self.prompt_encoder(points=points, boxes=boxes, masks=mask_input)
        dt = image_embedding.dtype
        if image_embedding.shape[0] == 1 or sparse.shape[0] == 1:
            return self.mask_decoder(image_embedding, self.prompt_encoder.get_dense_pe().to(dt), sparse, dense,
                                     multimask_output)
        # one image per prompt row: the decoder repeats every embedding for every prompt row, so
        # rows are decoded one by one here (the reference's eval runs with batch size 1)
        outs = [self.mask_decoder(image_embedding[i:i + 1], self.prompt_encoder.get_dense_pe().to(dt),
                                  sparse[i:i + 1], dense[i:i + 1], multimask_output) for i in range(sparse.shape[0])]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])


def build_sam(name: str = "vit_h", **encoder_overrides) -> Sam:
    """ViT-B / L / H SAM with the reference's decoder configuration (build_sam.py:47-105)."""
    prompt_embed_dim, image_size, patch = 256, 1024, 16
    grid = image_size // patch
    enc = build_image_encoder(name, **encoder_overrides)
    return Sam(
        image_encoder=enc,
        prompt_encoder=PromptEncoder(embed_dim=prompt_embed_dim, image_embedding_size=(grid, grid),
                                     input_image_size=(image_size, image_size), mask_in_chans=16),
        mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                 transformer=TwoWayTransformer(depth=2, embedding_dim=prompt_embed_dim,
                                                               mlp_dim=2048, num_heads=8),
                                 transformer_dim=prompt_embed_dim, iou_head_depth=3, iou_head_hidden_dim=256))


def get_iou(gt_mask: torch.Tensor, pred_mask: torch.Tensor, ignore_label: int = -1) -> torch.Tensor:
    """IoU of ``pred_mask`` (bool / 0-1) with the pixels ``gt_mask == 1``, ignoring ``ignore_label``
    pixels (evaluation2.py:156-167)."""
    keep = gt_mask != ignore_label
    obj = gt_mask == 1
    pred = pred_mask.bool()
    inter = (pred & obj & keep).sum()
    union = ((pred | obj) & keep).sum()
    return inter / union


def next_clicks(prev_logits: torch.Tensor, gt_masks: torch.Tensor, rng: np.random.Generator
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One simulated click per sample: a uniformly random pixel of the current error region,
    labelled 1 if it is a missed object pixel, 0 if a false positive (evaluation2.py:170-200).
    ``prev_logits`` / ``gt_masks`` ``[B, 1, H, W]`` -> (coords ``[B, 1, 2]`` as (x, y), labels ``[B, 1]``).
    A sample without error pixels gets a positive click on a random object pixel."""
    pred = prev_logits > 0
    true = gt_masks > 0
    fn = true & ~pred
    err = fn | (~true & pred)
    pts, labs = [], []
    for i in range(gt_masks.shape[0]):
        cand = torch.argwhere(err[i, 0])
        if cand.numel() == 0:
            cand = torch.argwhere(true[i, 0])
        y, x = (int(v) for v in cand[int(rng.integers(len(cand)))])
        pts.append([[x, y]])
        labs.append([int(bool(fn[i, 0, y, x]) or not bool(err[i, 0, y, x]))])
    dev = gt_masks.device
    return torch.tensor(pts, device=dev, dtype=torch.float32), torch.tensor(labs, device=dev, dtype=torch.float32)


@torch.no_grad()
def interactive_eval(sam: Sam, images: torch.Tensor, gt_masks: torch.Tensor, num_clicks: int = 5, seed: int = 0,
                     image_embeddings: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """The reference's click loop for a batch: ``images [B, 3, S, S]`` (already normalised and padded),
    ``gt_masks [B, 1, S, S]`` in {0, 1, -1 = ignore}.  Returns the IoU after every click
    (``iou_per_click [num_clicks, B]``), the final mean IoU and the final low-resolution logits."""
    dt = next(sam.image_encoder.parameters()).dtype
    emb = sam.image_encoder(images.to(dt)) if image_embeddings is None else image_embeddings
    rng = np.random.default_rng(seed)
    prev = torch.zeros_like(gt_masks, dtype=torch.float32)
    coords: List[torch.Tensor] = []
    labels: List[torch.Tensor] = []
    low_res = None
    ious = []
    for click in range(num_clicks):
        c, l = next_clicks(prev, gt_masks, rng)
        coords.append(c.to(dt))
        labels.append(l.to(dt))
        low_res, _ = sam.predict_masks(emb, points=(torch.cat(coords, dim=1), torch.cat(labels, dim=1)),
                                       mask_input=None if click == 0 else low_res, multimask_output=False)
        prev = F.interpolate(low_res.float(), size=gt_masks.shape[-2:], mode="bilinear", align_corners=False)
        pred = prev > sam.mask_threshold
        ious.append(torch.stack([get_iou(gt_masks[i], pred[i]) for i in range(gt_masks.shape[0])]))
    iou = torch.stack(ious)
    return {"iou_per_click": iou, "miou": float(iou[-1].mean()), "low_res_logits": low_res}
