"""Golden outputs of the REFERENCE's prompt encoder + mask decoder.

    python tests/golden/make_decoder_fixture.py      # build container only (needs /root/reference)

Weights: this package's modules are initialised under torch.manual_seed(SEED) and their state_dict
is loaded -- strict=True, i.e. the key / shape contract is checked here -- into the reference's
PromptEncoder / MaskDecoder / TwoWayTransformer
(/root/reference/segment_anything/modeling/{prompt_encoder,mask_decoder,transformer}.py), which then
run in fp32 on the CPU.  Only inputs' seeds and the reference OUTPUTS are stored
(tests/golden/decoder.npz); tests rebuild the same weights from the seed."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

SEED = 11


def build_ours():
    from sam_quantization_b200.mask_decoder import MaskDecoder, TwoWayTransformer
    from sam_quantization_b200.prompt_encoder import PromptEncoder

    torch.manual_seed(SEED)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    md = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                     transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    g = torch.Generator().manual_seed(SEED + 1)
    with torch.no_grad():                         # LayerNorms off their identity init, embeddings larger
        for m in list(pe.modules()) + list(md.modules()):
            if isinstance(m, torch.nn.LayerNorm) or type(m).__name__ == "LayerNorm2d":
                m.weight.add_(0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.add_(0.1 * torch.randn(m.bias.shape, generator=g))
    return pe.eval(), md.eval()


def inputs():
    rng = np.random.default_rng(SEED + 2)
    emb = (rng.standard_normal((1, 256, 64, 64)) * 0.5).astype(np.float16).astype(np.float32)
    pts = rng.uniform(0, 1024, size=(2, 3, 2)).astype(np.float32)
    labs = rng.integers(0, 2, size=(2, 3)).astype(np.float32)
    boxes = np.sort(rng.uniform(0, 1024, size=(2, 2, 2)), axis=1).reshape(2, 4).astype(np.float32)
    mask = (rng.standard_normal((2, 1, 256, 256)) * 3).astype(np.float16).astype(np.float32)
    return emb, pts, labs, boxes, mask


def main():
    from segment_anything.modeling.mask_decoder import MaskDecoder as RefDecoder
    from segment_anything.modeling.prompt_encoder import PromptEncoder as RefPrompt
    from segment_anything.modeling.transformer import TwoWayTransformer as RefTransformer

    pe, md = build_ours()
    rpe = RefPrompt(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    rmd = RefDecoder(num_multimask_outputs=3, transformer=RefTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                     transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    rpe.load_state_dict(pe.state_dict(), strict=True)
    rmd.load_state_dict(md.state_dict(), strict=True)
    rpe.eval(), rmd.eval()
    emb, pts, labs, boxes, mask = (torch.from_numpy(a) for a in inputs())
    out = {}
    with torch.no_grad():
        out["dense_pe_sub"] = rpe.get_dense_pe()[:, ::8, ::4, ::4].numpy()
        cases = {"points": dict(points=(pts, labs), boxes=None, masks=None, multi=True),
                 "points_mask": dict(points=(pts, labs), boxes=None, masks=mask, multi=False),
                 "boxes": dict(points=None, boxes=boxes, masks=None, multi=True),
                 "points_boxes": dict(points=(pts, labs), boxes=boxes, masks=None, multi=False)}
        for name, c in cases.items():
            sparse, dense = rpe(points=c["points"], boxes=c["boxes"], masks=c["masks"])
            masks, iou = rmd(image_embeddings=emb, image_pe=rpe.get_dense_pe(), sparse_prompt_embeddings=sparse,
                             dense_prompt_embeddings=dense, multimask_output=c["multi"])
            out[f"{name}_sparse"] = sparse.numpy()
            out[f"{name}_dense_sub"] = dense[:, ::8, ::4, ::4].contiguous().numpy()
            out[f"{name}_masks_sub"] = masks[:, :, ::4, ::4].contiguous().numpy()
            out[f"{name}_masks_absmax"] = np.float64(masks.abs().max())
            out[f"{name}_iou"] = iou.numpy()
            print(name, tuple(masks.shape), float(masks.abs().max()), iou.flatten()[:3].tolist())
    np.savez_compressed(os.path.join(HERE, "decoder.npz"), seed=SEED, **out)


if __name__ == "__main__":
    main()
