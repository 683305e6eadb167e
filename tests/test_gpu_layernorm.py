"""LayerNorm / window partition / unpartition+residual kernels vs the oracle (fp32 torch, CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oe
from sam_quantization_b200 import ops
from gpu_util import report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [768, 1024, 1280, 64, 2048])
def test_layernorm(cuda_device, C):
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(3, 64, 64, C, generator=g) * 2 + 0.5).half()
    w = (1 + 0.1 * torch.randn(C, generator=g)).half()
    b = (0.1 * torch.randn(C, generator=g)).half()
    y = ops.layernorm(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), 1e-6)
    ref = oe.layer_norm(x.float(), w.float(), b.float(), 1e-6)
    err, mag, cos = report(y, ref)
    assert err <= 2.0 ** -10 * mag and cos > 0.999999     # one fp16 ulp at the output magnitude


def test_layernorm_at_the_benchmarked_batch(cuda_device):
    """131072 x 1280 rows (batch 32 of ViT-H; ragged tail: + 5 rows), with a large common offset in
    the input (mean >> std: the two-pass statistics must not cancel): every row against an fp32
    LayerNorm on the GPU, and slice independence (first / last rows recomputed alone, same bits)."""
    rows, C = 131072 + 5, 1280
    g = torch.Generator(cuda_device).manual_seed(4)
    x = (torch.randn(rows, C, device=cuda_device, generator=g) * 1.5 + 40.0).half()
    w = (1 + 0.1 * torch.randn(C, device=cuda_device, generator=g)).half()
    b = (0.1 * torch.randn(C, device=cuda_device, generator=g)).half()
    y = ops.layernorm(x, w, b, 1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-6)
    err = (y.float() - ref).abs().max().item()
    assert err <= 2.0 ** -10 * ref.abs().max().item() + 1e-6, err      # one fp16 ulp of the output magnitude
    for sl in (slice(0, 37), slice(rows - 37, rows)):
        assert torch.equal(y[sl], ops.layernorm(x[sl].contiguous(), w, b, 1e-6))


@pytest.mark.parametrize("B,H,W,C", [(1, 64, 64, 1280), (2, 64, 64, 768), (3, 20, 30, 64)])
def test_layernorm_partition(cuda_device, B, H, W, C):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, C, generator=g).half()
    w = (1 + 0.1 * torch.randn(C, generator=g)).half()
    b = (0.1 * torch.randn(C, generator=g)).half()
    y, pad_hw = ops.layernorm_partition(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), 1e-6, 14)
    ref, ref_hw = oe.window_partition(oe.layer_norm(x.float(), w.float(), b.float(), 1e-6), 14)
    assert tuple(pad_hw) == tuple(ref_hw) and y.shape == ref.shape
    err, mag, _ = report(y, ref)
    assert err <= 2.0 ** -10 * mag
    # padded tokens are exactly zero (the pad is applied after the norm, image_encoder.py:300)
    assert torch.equal((y == 0).all(-1).cpu(), (ref == 0).all(-1))


def test_unpartition_residual_on_the_forks_fixture(cuda_device, golden_dir):
    """window_unpartition of the reference's own windows (ViT-H, batch 1) is reproduced exactly."""
    g = np.load(os.path.join(golden_dir, "partition_vith.npz"))
    win = torch.from_numpy(g["windows"]).half()
    back = torch.from_numpy(g["back"]).half()
    zero = torch.zeros_like(back)
    out = ops.unpartition_residual(win.to(cuda_device), zero.to(cuda_device), 14)
    assert torch.equal(out.cpu(), back)


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 1024), (1, 20, 30, 64)])
def test_unpartition_residual(cuda_device, B, H, W, C):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H, W, C, generator=g).half()
    sc = torch.randn(B, H, W, C, generator=g).half()
    win, pad_hw = oe.window_partition(x.float(), 14)
    out = ops.unpartition_residual(win.half().to(cuda_device), sc.to(cuda_device), 14)
    ref = (sc.float() + oe.window_unpartition(win, 14, pad_hw, (H, W))).half()
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(ops.add(x.to(cuda_device), sc.to(cuda_device)).cpu(), (x.float() + sc.float()).half())


def test_patchify_and_fused_patch_embed(cuda_device):
    """samq_patchify_fwd + dense GEMM == conv16x16/stride16 + permute + pos_embed add
    (image_encoder.py:107-109, 434-442); oracle = fp32 conv on the CPU."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 96, generator=g).half()
    rows = ops.patchify(x.to(cuda_device), 16)
    ref = torch.nn.functional.unfold(x.float(), kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 3 * 256).half()
    assert torch.equal(rows.cpu(), ref)
    w = (torch.randn(128, 3, 16, 16, generator=g) * 0.05).half()
    b = torch.randn(128, generator=g).half()
    pos = torch.randn(2, 4, 6, 128, generator=g).half()
    y = ops.dense_linear(rows, w.view(128, -1).to(cuda_device), b.to(cuda_device), residual=pos.to(cuda_device))
    conv = torch.nn.functional.conv2d(x.float(), w.float(), b.float(), stride=16).permute(0, 2, 3, 1) + pos.float()
    err, mag, cos = report(y.view(2, 4, 6, 128), conv)
    assert err <= 2.0 ** -9 * mag and cos > 0.99999


@pytest.mark.parametrize("B,H,W,C,O", [(2, 64, 64, 256, 256), (1, 5, 7, 64, 256), (3, 1, 1, 8, 256)])
def test_im2col3x3_and_neck_conv(cuda_device, B, H, W, C, O):
    """samq_im2col3x3_fwd + dense GEMM == Conv2d(C, O, 3, padding=1, bias=False) in NHWC
    (the neck's second convolution, image_encoder.py:96-103); the re-layout is exact against
    torch's unfold, the product is checked against an fp32 convolution on the CPU."""
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, H, W, C, generator=g).half()
    rows = ops.im2col3x3(x.to(cuda_device))
    assert rows.shape == (B * H * W, 9 * C)
    unf = torch.nn.functional.unfold(x.float().permute(0, 3, 1, 2), kernel_size=3, padding=1)   # [B, C*9, H*W], (c, ky, kx)
    ref = unf.view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C).half()          # -> (ky, kx, c)
    assert torch.equal(rows.cpu(), ref)
    if (9 * C) % 64 == 0:
        w = (torch.randn(O, C, 3, 3, generator=g) * 0.05).half()
        w3 = w.permute(0, 2, 3, 1).reshape(O, 9 * C).contiguous()
        y = ops.dense_linear(rows, w3.to(cuda_device)).view(B, H, W, O)
        conv = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1).permute(0, 2, 3, 1)
        err, mag, cos = report(y, conv)
        assert err <= 2.0 ** -9 * max(mag, 1.0) and cos > 0.99999
    if H > 1 and W > 1:
        with pytest.raises(AssertionError):                     # non-contiguous input
            ops.im2col3x3(x.to(cuda_device).transpose(1, 2))
    with pytest.raises(ValueError):
        ops.im2col3x3(x.to(cuda_device).view(B * H, W, C))      # not [B, H, W, C]
