"""samq_attn_relpos_fwd vs the oracle's eager attention (fp32 softmax; the formula of the
reference's only self-check, gptq_triton/fused_attention.py:388-406) at its two shapes, seed
20, N(0, 0.5) inputs.  Tolerance: the reference's own (commented-out) atol = 1e-2, rtol = 0
(fused_attention.py:409); measured errors are ~5e-4 and are asserted at 2e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oe
from sam_quantization_b200 import _lib, ops
from gpu_util import report

pytestmark = pytest.mark.gpu

ATOL = 2e-3


def make_inputs(B, E, heads, hd, seed=20, std=0.5, rp_std=0.3):
    g = torch.Generator().manual_seed(seed)
    qkv = torch.empty(B, E * E, 3 * heads * hd).normal_(0.0, std, generator=g).half()
    rph = torch.empty(2 * E - 1, hd).normal_(0.0, rp_std, generator=g).half()
    rpw = torch.empty(2 * E - 1, hd).normal_(0.0, rp_std, generator=g).half()
    return qkv, rph, rpw


@pytest.mark.parametrize("B,E,heads,hd", [(25, 14, 16, 80), (1, 64, 16, 80), (25, 14, 16, 64), (1, 64, 12, 64)])
@pytest.mark.parametrize("relw", ["reference", "upstream"])
def test_attention_test_op_shapes(cuda_device, B, E, heads, hd, relw):
    qkv, rph, rpw = make_inputs(B, E, heads, hd)
    out = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), B, E, E, heads, 0.5,
                          _lib.RELW_UPSTREAM if relw == "upstream" else _lib.RELW_REFERENCE)
    ref = oe.attention_core(qkv, rph, rpw, B, E, E, heads, 0.5, relw, round_tables=True)
    err, mag, cos = report(out, ref)
    assert not torch.isnan(out).any()
    assert err <= ATOL and cos >= 0.9999, (err, mag, cos)


@pytest.mark.parametrize("B,E", [(800, 14), (32, 64)])
def test_full_size_properties_at_the_benchmarked_batch(cuda_device, B, E):
    """BASELINE's full sizes (batch 32 of ViT-H: 800 windows x 16 heads x 196 keys, 32 images x 16 heads
    x 4096 keys), where the oracle would take minutes: size-independent properties of softmax(.) V.
    (1) With V = 1 every output is a convex combination of ones: exactly 1 up to the fp16 roundings of P
    and of O / l.  (2) The output is linear in V: attn(V1 + V2) = attn(V1) + attn(V2) for the same
    scores.  (3) No op crosses the batch: the first and last items recomputed alone give the same bits.
    (4) A sample of items against the oracle."""
    heads, hd = 16, 80
    D = heads * hd
    g = torch.Generator(device=cuda_device).manual_seed(31)
    qkv = torch.empty(B, E * E, 3 * D, device=cuda_device, dtype=torch.float16).normal_(0.0, 0.5, generator=g)
    rph = torch.empty(2 * E - 1, hd, device=cuda_device, dtype=torch.float16).normal_(0.0, 0.3, generator=g)
    rpw = torch.empty(2 * E - 1, hd, device=cuda_device, dtype=torch.float16).normal_(0.0, 0.3, generator=g)
    scale = hd ** -0.5
    run = lambda t, b=B: ops.attn_relpos(t, rph, rpw, b, E, E, heads, scale)
    out = run(qkv)
    assert not torch.isnan(out).any()
    ones = qkv.clone()
    ones[..., 2 * D:] = 1.0
    o1 = run(ones)
    assert float((o1.float() - 1.0).abs().max()) <= 2e-3
    v2 = torch.empty(B, E * E, D, device=cuda_device, dtype=torch.float16).normal_(0.0, 0.5, generator=g)
    other = qkv.clone()
    other[..., 2 * D:] = v2
    both = qkv.clone()
    both[..., 2 * D:] = (qkv[..., 2 * D:].float() + v2.float()).half()
    lin = (run(both).float() - out.float() - run(other).float()).abs().max()
    assert float(lin) <= 4e-3, float(lin)     # three fp16-rounded outputs + the rounding of V1 + V2
    k = 3 if E == 14 else 1
    assert torch.equal(out[:k], run(qkv[:k].contiguous(), k)) and torch.equal(out[-k:], run(qkv[-k:].contiguous(), k))
    pick = [0, B // 2, B - 1][:k if E == 64 else 3]
    ref = oe.attention_core(qkv[pick].cpu(), rph.cpu(), rpw.cpu(), len(pick), E, E, heads, scale, "reference",
                            round_tables=True)
    err, mag, cos = report(out[pick], ref)
    assert err <= ATOL and cos >= 0.9999, (err, mag, cos)


@pytest.mark.parametrize("B,heads,hd", [(1, 1, 80), (1, 3, 64), (10, 16, 80), (37, 5, 64), (60, 16, 80)])
def test_windowed_item_scheduling(cuda_device, B, heads, hd):
    """The windowed kernel is persistent (one CTA per SM looping over (window, head) items):
    fewer items than SMs, a ragged last round (160 = 148 + 12, 185 items) and many rounds."""
    qkv, rph, rpw = make_inputs(B, 14, heads, hd, seed=7)
    args = (qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), B, 14, 14, heads, hd ** -0.5)
    out = ops.attn_relpos(*args)
    ref = oe.attention_core(qkv, rph, rpw, B, 14, 14, heads, hd ** -0.5, "reference", round_tables=True)
    err, mag, cos = report(out, ref)
    assert not torch.isnan(out).any()
    assert err <= ATOL and cos >= 0.9999, (err, mag, cos)
    assert torch.equal(out, ops.attn_relpos(*args)), "run-to-run determinism"


@pytest.mark.parametrize("env,val,B,E", [("SAMQ_ATTN_WIN", "v2", 25, 14), ("SAMQ_ATTN_WIN", "v1", 25, 14),
                                         ("SAMQ_ATTN_GLOB", "v1", 1, 64), ("SAMQ_ATTN_GLOB", "v2", 1, 64),
                                         ("SAMQ_ATTN_MAX", "exact", 25, 14), ("SAMQ_ATTN_MAX", "exact", 1, 64)])
def test_ablation_kernels_agree_with_default(cuda_device, samq_env, env, val, B, E):
    """The earlier kernel designs stay selectable (SAMQ_ATTN_WIN=v1|v2, SAMQ_ATTN_GLOB=v1|v2), and so
    does the exact row maximum (SAMQ_ATTN_MAX=exact), for A/B timing; they must compute the same
    function."""
    from sam_quantization_b200 import _lib
    if val in ("v1", "v2") and not _lib.has_ablations():
        pytest.skip("earlier kernel generations are only in `make ABLATIONS=1` builds (SAMQ_LIB=...)")
    qkv, rph, rpw = make_inputs(B, E, 4, 80, seed=11)
    args = (qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), B, E, E, 4, 80 ** -0.5)
    new = ops.attn_relpos(*args)
    samq_env.set(env, val)
    old = ops.attn_relpos(*args)
    samq_env.unset(env)
    assert (new.float() - old.float()).abs().max().item() <= 1e-3


@pytest.mark.parametrize("B,H,W,heads,hd", [(2, 64, 64, 16, 80), (1, 64, 64, 12, 64), (3, 20, 30, 2, 80),
                                            (1, 14, 14, 3, 64), (2, 9, 40, 2, 80)])
@pytest.mark.parametrize("relw", [0, 1])
def test_windowed_attention_with_fused_unpartition(cuda_device, samq_env, B, H, W, heads, hd, relw):
    """samq_attn_relpos_unpartition_fwd == window_unpartition(samq_attn_relpos_fwd(windows)) bit for
    bit (image_encoder.py:309-333): each window tile leaves as one TMA box placed at the window's
    position in the image, and the box elements outside the image are not written."""
    samq_env.unset("SAMQ_ATTN_WIN")   # both sides on the same (default) kernel
    ws = 14
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    Bw = B * nH * nW
    qkv, rph, rpw = make_inputs(Bw, ws, heads, hd, seed=13)
    qkv_d, rph_d, rpw_d = qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device)
    win = ops.attn_relpos(qkv_d, rph_d, rpw_d, Bw, ws, ws, heads, hd ** -0.5, relw)
    out = torch.full((B, H, W, heads * hd), float("nan"), dtype=torch.float16, device=cuda_device)
    img = ops.attn_relpos_unpartition(qkv_d.view(Bw, ws, ws, -1), rph_d, rpw_d, B, H, W, ws, heads, hd ** -0.5, relw)
    ref = oe.window_unpartition(win.float().cpu(), ws, (nH * ws, nW * ws), (H, W)).half()
    assert img.shape == (B, H, W, heads * hd)
    assert not torch.isnan(img).any()
    assert torch.equal(img.cpu(), ref)
    del out


def test_relw_modes_differ(cuda_device):
    qkv, rph, rpw = make_inputs(2, 14, 2, 64)
    a = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), 2, 14, 14, 2, 0.125, 0)
    b = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), 2, 14, 14, 2, 0.125, 1)
    assert (a.float() - b.float()).abs().max().item() > 1e-2


def test_large_logits_exercise_the_lazy_rescale(cuda_device):
    """std 2.0 inputs give row maxima that keep growing across key tiles (rescale path)."""
    qkv, rph, rpw = make_inputs(1, 64, 2, 80, seed=3, std=2.0, rp_std=0.5)
    out = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), 1, 64, 64, 2, 80 ** -0.5)
    ref = oe.attention_core(qkv, rph, rpw, 1, 64, 64, 2, 80 ** -0.5, "reference", round_tables=True)
    err, mag, cos = report(out, ref)
    assert err <= 2e-2 * max(1.0, mag) and cos >= 0.999, (err, mag, cos)


def test_windowed_softmax_when_later_keys_dominate(cuda_device):
    """Keys of the later window rows are made up to 6x larger than the first rows': whatever reference
    the kernel subtracts in the exponent has to be the (bound on the) maximum over ALL 196 keys.  (A
    single-pass variant that took its reference from the first two key rows and rescaled P in TMEM on
    a jump was measured in round 2 and dropped -- same time as two passes -- but this case stays.)"""
    B, E, heads, hd = 9, 14, 2, 80
    qkv, rph, rpw = make_inputs(B, E, heads, hd, seed=11, std=1.0, rp_std=0.3)
    k = qkv.view(B, E * E, 3, heads, hd)[:, :, 1]
    k[:, 56:] *= 3.0
    k[:, 140:] *= 2.0
    out = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), B, E, E, heads, hd ** -0.5)
    ref = oe.attention_core(qkv, rph, rpw, B, E, E, heads, hd ** -0.5, "reference", round_tables=True)
    err, mag, cos = report(out, ref)
    assert not torch.isnan(out).any()
    assert err <= 2e-2 * max(1.0, mag) and cos >= 0.999, (err, mag, cos)


@pytest.mark.parametrize("B,E", [(25, 14), (1, 64)])
@pytest.mark.parametrize("rp_std,tol", [(0.0, 1.0), (0.05, 1.0), (0.3, 1.0), (1.5, 5.0)])
def test_bound_and_exact_row_maximum_paths(cuda_device, B, E, rp_std, tol):
    """The softmax replaces the row maximum by a bound when the column biases of a warp's rows lie
    within 2^15 of each other and takes the exact maximum otherwise: zero tables (the reference's
    benchmark init, image_encoder.py:235-236), narrow, default and wide bias spreads -- the last
    one sends every warp down the exact path (q.Rw has sigma ~ 10 in log2 units)."""
    qkv, rph, rpw = make_inputs(B, E, 2, 80, seed=5, std=0.5, rp_std=max(rp_std, 1e-9))
    if rp_std == 0.0:
        rph.zero_(), rpw.zero_()
    out = ops.attn_relpos(qkv.to(cuda_device), rph.to(cuda_device), rpw.to(cuda_device), B, E, E, 2, 80 ** -0.5)
    ref = oe.attention_core(qkv, rph, rpw, B, E, E, 2, 80 ** -0.5, "reference", round_tables=True)
    err, mag, cos = report(out, ref)
    assert not torch.isnan(out).any()
    assert err <= tol * ATOL * max(1.0, mag) and cos >= 0.9999, (err, mag, cos)


@pytest.mark.parametrize("name,size", [("attn_win", 14), ("attn_glob", 64)])
def test_against_reference_attention_module_fixture(cuda_device, golden_dir, name, size):
    """Output of the reference's own ``Attention`` module (fp32, CPU) reproduced with fp16 qkv/proj
    around the CUDA kernel.  Tolerance covers the fp16 rounding of qkv and of the output."""
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    heads = int(g["heads"])
    x = torch.from_numpy(g["x16"].astype(np.float32)).to(cuda_device)
    B, dim = x.shape[0], x.shape[-1]
    W = {k: torch.from_numpy(g[k]).to(cuda_device) for k in ("qkv.weight", "qkv.bias", "proj.weight", "proj.bias")}
    qkv = torch.nn.functional.linear(x, W["qkv.weight"], W["qkv.bias"]).half().reshape(B, size * size, -1)
    o = ops.attn_relpos(qkv.contiguous(), torch.from_numpy(g["rel_pos_h"]).half().to(cuda_device),
                        torch.from_numpy(g["rel_pos_w"]).half().to(cuda_device), B, size, size, heads,
                        (dim // heads) ** -0.5)
    y = torch.nn.functional.linear(o.float(), W["proj.weight"], W["proj.bias"]).reshape(B, size * size, dim)
    err, mag, cos = report(y[:, :: int(g["stride"])], g["y_sub"])
    assert err <= 5e-3 * max(1.0, mag) and cos >= 0.9999, (err, mag, cos)


def test_unsupported_shapes_raise(cuda_device):
    qkv = torch.zeros(1, 32 * 32, 3 * 2 * 64, dtype=torch.float16, device=cuda_device)
    rp = torch.zeros(63, 64, dtype=torch.float16, device=cuda_device)
    with pytest.raises(AssertionError):
        ops.attn_relpos(qkv, rp, rp, 1, 32, 32, 2, 0.1)


def test_quant_attention_interpolates_rel_pos_tables_of_another_grid(cuda_device):
    """A checkpoint trained at another grid size carries rel-pos tables that are not 2*size-1 rows
    long; the reference resizes them linearly (image_encoder.py:348-358).  QuantAttention does the
    same once per parameter version and hands the kernel the resized tables: result = the oracle's
    attention with tables resized by the (reference-pinned) host get_rel_pos path."""
    from sam_quantization_b200 import image_encoder as ie
    from sam_quantization_b200.fused_attention import QuantAttention

    heads, hd, E, B = 2, 64, 14, 3
    g = torch.Generator().manual_seed(5)
    qkv = torch.empty(B, E * E, 3 * heads * hd).normal_(0.0, 0.5, generator=g).half()
    rph = torch.empty(2 * 32 - 1, hd).normal_(0.0, 0.3, generator=g)          # trained for a 32 x 32 grid
    rpw = torch.empty(2 * 32 - 1, hd).normal_(0.0, 0.3, generator=g)
    attn = QuantAttention(None, None, heads, 0.5, True, torch.nn.Parameter(rph.to(cuda_device)),
                          torch.nn.Parameter(rpw.to(cuda_device)))
    out = attn.attention(qkv.to(cuda_device), B, E, E)
    t1 = attn._tables(E, E)
    assert t1[0].shape == (2 * E - 1, hd) and attn._tables(E, E)[0] is t1[0]                 # resized once, cached
    ref = oe.attention_core(qkv, ie.resize_rel_pos(rph, 2 * E - 1).half(), ie.resize_rel_pos(rpw, 2 * E - 1).half(),
                            B, E, E, heads, 0.5, "reference", round_tables=True)
    err, mag, cos = report(out, ref)
    assert err <= ATOL and cos >= 0.9999, (err, mag, cos)
