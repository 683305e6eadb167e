"""Mask decoder on the B200: the kernels behind it against torch fp32, and the
whole prompt-encoder + decoder chain in fp16 on libsamq kernels against the REFERENCE's modules'
fp32 outputs (tests/golden/decoder.npz).

Tolerances: kernels -- fp32 math on fp16 operands, one fp16 rounding: 2^-10 of the output's
magnitude; decoder chain -- fp16 storage between ~40 kernels: mask logits max-abs <= 3e-2 max|ref|
and cosine >= 0.999, IoU head abs 2e-2."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_decoder_fixture as mk  # noqa: E402

from sam_quantization_b200 import _lib, ops  # noqa: E402
from test_decoder_cpu import CASES, run_case  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,heads,Nq,Nk,hd", [(2, 8, 9, 4096, 16), (2, 8, 4096, 9, 16), (3, 8, 7, 7, 32),
                                              (1, 4, 33, 100, 64), (1, 8, 1, 1, 16)])
def test_attn_small(cuda_device, B, heads, Nq, Nk, hd):
    g = torch.Generator().manual_seed(Nq * 7 + Nk)
    C = heads * hd
    q, k, v = (torch.randn(B, n, C, generator=g).half() for n in (Nq, Nk, Nk))
    out = ops.attn_small(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), heads).float().cpu()
    qh, kh, vh = (t.float().reshape(B, -1, heads, hd).transpose(1, 2) for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / hd ** 0.5, dim=-1) @ vh).transpose(1, 2).reshape(B, Nq, C)
    assert (out - ref).abs().max().item() <= 2.0 ** -10 * ref.abs().max().item() + 1e-4


@pytest.mark.parametrize("M,N,K,act", [(9, 256, 256, 0), (18, 2048, 256, 2), (18, 256, 2048, 0), (4, 32, 256, 2),
                                       (65536, 4, 32, 0), (5, 4, 256, 1)])
def test_small_linear(cuda_device, M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b, r = (torch.randn(*s, generator=g).half() for s in ((M, K), (N, K), (N,), (M, N)))
    w = (w.float() / K ** 0.5).half()
    y = ops.small_linear(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), act, r.to(cuda_device)).float().cpu()
    pre = x.float() @ w.float().t() + b.float()
    pre = torch.nn.functional.gelu(pre) if act == 1 else torch.relu(pre) if act == 2 else pre
    ref = pre.half().float() + r.float()                      # fp16 add after the rounding
    assert (y - ref).abs().max().item() <= 2.0 ** -9 * ref.abs().max().item()
    assert torch.equal(ops.gelu(x.to(cuda_device)).cpu(), torch.nn.functional.gelu(x.float()).half()) or \
        (ops.gelu(x.to(cuda_device)).cpu().float() - torch.nn.functional.gelu(x.float())).abs().max() <= 2.0 ** -10 * 4


@pytest.mark.parametrize("name", list(CASES))
def test_decoder_chain_on_kernels_against_the_reference_modules(cuda_device, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "decoder.npz"))
    pe, md = mk.build_ours()
    pe, md = pe.half().to(cuda_device), md.half().to(cuda_device)
    launches = _lib.launch_count()
    with torch.no_grad():
        sparse, dense, masks, iou = run_case(pe, md, name, device=cuda_device, dtype=torch.float16)
    assert _lib.launch_count() - launches >= 60, "the decoder did not run on libsamq kernels"
    ref = torch.from_numpy(g[f"{name}_masks_sub"])
    sub = masks[:, :, ::4, ::4].float().cpu()
    err = (sub - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(sub.flatten().double(), ref.flatten().double(), dim=0).item()
    print(f"{name}: masks max-abs {err:.3e} (max|ref| {float(g[f'{name}_masks_absmax']):.3f}) cosine {cos:.6f}")
    assert err <= 3e-2 * float(g[f"{name}_masks_absmax"]) and cos >= 0.999
    assert np.abs(iou.float().cpu().numpy() - g[f"{name}_iou"]).max() <= 2e-2
    # click coordinates arrive in fp16 here (0.5-pixel grid above 512): sin / cos of 2 pi G x moves by ~1e-2
    assert np.abs(sparse.float().cpu().numpy() - g[f"{name}_sparse"]).max() <= 3e-2


def test_click_loop_on_the_device(cuda_device):
    """interactive_eval end to end on the GPU (decoder on kernels, given embeddings)."""
    from sam_quantization_b200 import sam as S

    pe, md = mk.build_ours()
    sam = S.Sam(torch.nn.Module(), pe, md).half().to(cuda_device)
    sam.image_encoder.register_parameter("dummy", torch.nn.Parameter(torch.zeros(1, dtype=torch.float16, device=cuda_device)))
    emb = torch.from_numpy(mk.inputs()[0]).half().to(cuda_device)
    gt = torch.zeros(1, 1, 1024, 1024, device=cuda_device)
    gt[0, 0, 300:700, 200:800] = 1
    r = S.interactive_eval(sam, torch.zeros(1, 3, 1024, 1024, device=cuda_device), gt, num_clicks=5, seed=3,
                           image_embeddings=emb)
    assert r["iou_per_click"].shape == (5, 1) and torch.isfinite(r["low_res_logits"]).all()


def test_sam_forward_end_to_end_on_the_device(cuda_device):
    """Sam.forward on the GPU: preprocess -> GPTQ-int4 encoder on the fused kernels -> prompt encoder ->
    mask decoder on kernels -> postprocess, against the host chain (oracle encoder on the dequantised
    weights in fp32 + this package's fp32 decoder path, itself pinned to the reference's modules)."""
    from oracle import encoder as oe
    from sam_quantization_b200 import sam as S
    from sam_quantization_b200.synthetic import random_quantized_encoder

    cfg = dict(embed_dim=256, depth=2, num_heads=4, global_attn_indexes=(1,))
    enc = random_quantized_encoder("vit_b", 4, 128, seed=1, device=cuda_device, **cfg)
    pe, md = mk.build_ours()
    sam_gpu = S.Sam(enc, pe.half(), md.half()).to(cuda_device).eval()
    g = torch.Generator().manual_seed(8)
    batch = [
        {"image": torch.rand(3, 768, 1024, generator=g) * 255, "original_size": (600, 800),
         "point_coords": torch.rand(2, 2, 2, generator=g) * 700, "point_labels": torch.ones(2, 2)},
        {"image": torch.rand(3, 1024, 1024, generator=g) * 255, "original_size": (512, 512),
         "boxes": torch.tensor([[100.0, 120.0, 400.0, 700.0]])},
    ]
    to_dev = lambda rec: {k: (v.to(cuda_device) if torch.is_tensor(v) else v) for k, v in rec.items()}
    launches = _lib.launch_count()
    out = sam_gpu([to_dev(r) for r in batch], multimask_output=False)
    assert _lib.launch_count() - launches > 150
    # host chain on identical weights
    state = {k.replace(".attn.qkv_proj.", ".attn.qkv.").replace(".attn.o_proj.", ".attn.proj."): v.detach().cpu()
             for k, v in enc.state_dict().items()}
    p = oe.dequant_state(state, 4, 128)
    pe32, md32 = mk.build_ours()
    host = S.Sam(torch.nn.Module(), pe32, md32)
    host.image_encoder.img_size = 1024
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for rec, got in zip(batch, out):
        img = host.preprocess(rec["image"]).half().float().unsqueeze(0)
        with torch.no_grad():
            emb = oe.encoder(img, p, cfg["depth"], cfg["num_heads"], cfg["global_attn_indexes"])
            pts = (rec["point_coords"], rec["point_labels"]) if "point_coords" in rec else None
            low, iou = host.predict_masks(emb, points=pts, boxes=rec.get("boxes"), multimask_output=False)
        ref = low.float()
        err = (got["low_res_logits"].float().cpu() - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got["low_res_logits"].float().cpu().flatten().double(),
                                                    ref.flatten().double(), dim=0).item()
        print(f"Sam.forward: low-res logits max-abs {err:.3e} (max|ref| {ref.abs().max().item():.3f}) cosine {cos:.6f}")
        assert err <= 5e-2 * max(1.0, ref.abs().max().item()) and cos >= 0.995
        assert got["masks"].shape[-2:] == tuple(rec["original_size"]) and got["masks"].dtype == torch.bool
