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
