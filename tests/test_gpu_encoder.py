"""Block / encoder parity on the GPU: the fused CUDA path (loaded through load_quant from a
reference-layout checkpoint directory) vs the CPU oracle on the dequantised weights.

Tolerances (fp16 storage between kernels, fp32 accumulation inside; stated per test):
token outputs  max-abs <= 1.5e-2 * max|ref|, cosine >= 0.9999;
encoder embeddings (after the neck's LayerNorm2d) max-abs <= 6e-2, cosine >= 0.999."""
import os

import numpy as np
import pytest
import torch

import sam_quantization_b200 as sq
from oracle import encoder as oe
from oracle import synth
from sam_quantization_b200 import image_encoder as ie
from gpu_util import report

pytestmark = pytest.mark.gpu


def build_from_checkpoint(tmp_path, cfg, bits, gs, seed, device, act_order=False, relw_mode="reference"):
    p = synth.fp_state(seed=seed, **cfg)
    rng = np.random.default_rng(seed + 50)
    for k in p:
        if "rel_pos" in k:   # make the rel-pos path matter (zero-init in the reference, trap 5)
            p[k] = (rng.standard_normal(p[k].shape) * 0.2).astype(np.float32)
    packed = synth.quantize_state(p, bits, gs, act_order_seed=(seed if act_order else None))
    state = synth.to_torch(packed)
    torch.save(state, tmp_path / "model.pt")
    import json
    json.dump({"wbits": bits, "groupsize": gs}, open(tmp_path / "quant_config.json", "w"))
    enc = ie.ImageEncoderViT(img_size=1024, patch_size=16, use_rel_pos=True, window_size=14,
                             embed_dim=cfg["embed_dim"], depth=cfg["depth"], num_heads=cfg["num_heads"],
                             global_attn_indexes=cfg["global_attn_indexes"]).half()
    enc = sq.load_quant(enc, str(tmp_path), warmup_autotune=False, device=device, relw_mode=relw_mode).eval()
    ref_state = oe.dequant_state(state, bits, gs)
    return enc, ref_state


@pytest.mark.parametrize("dim,heads,gs", [(128, 2, 64), (640, 8, 128)])
@pytest.mark.parametrize("relw", ["reference", "upstream"])
def test_blocks_small(cuda_device, tmp_path, dim, heads, gs, relw):
    cfg = dict(embed_dim=dim, depth=2, num_heads=heads, global_attn_indexes=(1,))
    enc, ref_state = build_from_checkpoint(tmp_path, cfg, 4, gs, seed=1, device=cuda_device, relw_mode=relw)
    assert all(b._fused_ready() for b in enc.blocks)
    x = torch.from_numpy(synth.tokens_input(2, 64, dim, seed=3)).half()
    with torch.no_grad():
        y = enc.forward_tokens(x.to(cuda_device))
        ref = oe.tokens_forward(x.float(), ref_state, 2, heads, 14, (1,), relw)
    err, mag, cos = report(y, ref)
    print(f"blocks dim={dim} relw={relw}: max-abs {err:.3e} (max|ref| {mag:.3f}) cosine {cos:.7f}")
    assert err <= 1.5e-2 * mag and cos >= 0.9999


@pytest.mark.parametrize("bits", [3, 8])
def test_blocks_other_bits_with_act_order(cuda_device, tmp_path, bits):
    """BASELINE config 4 at a small width: 3-/8-bit with a permutation-derived g_idx."""
    cfg = dict(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,))
    enc, ref_state = build_from_checkpoint(tmp_path, cfg, bits, 64, seed=2, device=cuda_device, act_order=True)
    assert enc.blocks[0].attn.qkv_proj.g_idx is not None
    x = torch.from_numpy(synth.tokens_input(1, 64, 128, seed=4)).half()
    with torch.no_grad():
        y = enc.forward_tokens(x.to(cuda_device))
        ref = oe.tokens_forward(x.float(), ref_state, 2, 2, 14, (1,), "reference")
    err, mag, cos = report(y, ref)
    assert err <= 1.5e-2 * mag and cos >= 0.9999


def test_vit_b_encoder_config1(cuda_device, tmp_path):
    """BASELINE config 1: ViT-B, int4 g128, one synthetic 1024x1024 image, whole encoder."""
    cfg = dict(oe.CONFIGS["vit_b"])
    enc, ref_state = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=0, device=cuda_device)
    img = torch.from_numpy(synth.image(1, 1024, seed=0))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        y = enc(img.half().to(cuda_device))
        ref = oe.encoder(img.half().float(), ref_state, **cfg)
    assert y.shape == (1, 256, 64, 64)
    err, mag, cos = report(y, ref)
    print(f"ViT-B encoder: max-abs {err:.3e} (max|ref| {mag:.3f}) cosine {cos:.7f}")
    assert err <= 6e-2 and cos >= 0.999


def test_batch_invariance_and_sharding_property(cuda_device, tmp_path):
    """Size-independent property for the batched configs: encoding a batch equals encoding
    each image alone (bit-exact: no op crosses the batch), so shards over GPUs are replicas."""
    cfg = dict(embed_dim=640, depth=2, num_heads=8, global_attn_indexes=(1,))
    enc, _ = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=5, device=cuda_device)
    x = torch.from_numpy(synth.tokens_input(3, 64, 640, seed=6)).half().to(cuda_device)
    with torch.no_grad():
        full = enc.forward_tokens(x)
        parts = torch.cat([enc.forward_tokens(x[i:i + 1]) for i in range(3)])
    assert torch.equal(full, parts)


@pytest.mark.parametrize("batch", [1, 4])
def test_pad_skipping_block_is_bit_identical(cuda_device, tmp_path, samq_env, batch):
    """The windowed block's default form (LayerNorm in image order -> qkv GEMM that partitions in its
    store -> attention that un-partitions in its store -> plain proj GEMM) never multiplies the
    zero-padding tokens of the 70x70 window layout; SAMQ_PAD_SKIP=0 is the reference's order of
    operations (partition, multiply everything, drop).  Same dot products -> identical bits.
    Both batches take the unpack-once + CTA-pair GEMM path by default (M >= 2048)."""
    samq_env.unset("SAMQ_ATTN_WIN")   # both forms on the default windowed kernel
    cfg = dict(embed_dim=640, depth=2, num_heads=8, global_attn_indexes=(1,))
    enc, _ = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=8, device=cuda_device)
    x = torch.from_numpy(synth.tokens_input(batch, 64, 640, seed=9)).half().to(cuda_device)
    with torch.no_grad():
        samq_env.set("SAMQ_PAD_SKIP", "0")
        old = enc.forward_tokens(x)
        samq_env.set("SAMQ_PAD_SKIP", "1")
        new = enc.forward_tokens(x)
    assert torch.equal(old, new)


@pytest.mark.parametrize("batch", [1, 3])
def test_weight_prefetch_is_bit_identical_and_used(cuda_device, tmp_path, samq_env, batch, monkeypatch):
    """Default: right behind each linear's GEMM the NEXT linear's weight is unpacked next to it
    (samq_qlinear_prefetch, rotating scratch buffers) and that linear then runs only its GEMM;
    SAMQ_PREFETCH=0: every linear unpacks its own weight right before its GEMM.  Same kernels'
    arithmetic -> identical bits; repeated passes (buffer rotation) stay identical; and the
    prefetch really happens: 4 linears x depth - 1 times per pass."""
    from sam_quantization_b200 import ops
    cfg = dict(embed_dim=640, depth=3, num_heads=8, global_attn_indexes=(1,))
    enc, _ = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=18, device=cuda_device)
    x = torch.from_numpy(synth.tokens_input(batch, 64, 640, seed=19)).half().to(cuda_device)
    monkeypatch.setattr(ops, "PREFETCH_MIN_M", 2048)     # the product threshold is batch >= 8
    calls = []
    real = ops.qlinear_prefetch
    monkeypatch.setattr(ops, "qlinear_prefetch", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    with torch.no_grad():
        samq_env.set("SAMQ_PREFETCH", "0")
        plain = enc.forward_tokens(x)
        assert not calls
        samq_env.set("SAMQ_PREFETCH", "1")
        outs = [enc.forward_tokens(x) for _ in range(3)]
    assert len(calls) == 3 * (4 * 3 - 1)
    for o in outs:
        assert torch.equal(o, plain)


def test_stale_prefetched_weight_is_discarded(cuda_device, tmp_path, monkeypatch):
    """A prefetched weight is only used if the layer's packed buffers are still the ones that were
    unpacked: a standalone call leaves the next layer's weight prefetched; re-packing that layer
    in between must win."""
    from sam_quantization_b200 import ops, quant_linear as ql
    monkeypatch.setattr(ops, "PREFETCH_MIN_M", 2048)
    cfg = dict(embed_dim=640, depth=2, num_heads=8, global_attn_indexes=(1,))
    enc, _ = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=28, device=cuda_device)
    x = torch.from_numpy(synth.tokens_input(1, 64, 640, seed=29)).half().to(cuda_device)
    with torch.no_grad():
        enc.forward_tokens(x)                               # links the chain
        lin1, lin2 = enc.blocks[0].mlp.lin1, enc.blocks[0].mlp.lin2
        h = lin1(x.view(-1, 640))                           # prefetches lin2's weight
        chain, idx = ql._PREFETCH[lin2]
        assert chain.ready is not None and chain.ready[0] == idx
        before = lin2(h)                                    # consumes it
        lin1(x.view(-1, 640))                               # prefetches lin2 again ...
        lin2.scales.mul_(2.0)                               # ... then the layer changes
        after = lin2(h)
        bias = lin2.bias.float()
    assert torch.allclose((after.float() - bias), 2.0 * (before.float() - bias), rtol=2e-3, atol=2e-3)


def test_vit_l_width_blocks_batched(cuda_device, tmp_path):
    """BASELINE config 2 shapes (ViT-L: dim 1024, 16 heads of 64, batch > 1 so that the GEMMs take
    the unpack-once + CTA-pair path and the windowed block its pad-skipping form), two blocks
    (windowed + global) against the oracle.  Parity unpinned by reference for batch > 1."""
    cfg = dict(embed_dim=1024, depth=2, num_heads=16, global_attn_indexes=(1,))
    enc, ref_state = build_from_checkpoint(tmp_path, cfg, 4, 128, seed=21, device=cuda_device)
    x = torch.from_numpy(synth.tokens_input(3, 64, 1024, seed=22)).half()
    with torch.no_grad():
        y = enc.forward_tokens(x.to(cuda_device))
        ref = oe.tokens_forward(x.float(), ref_state, 2, 16, 14, (1,), "reference")
    err, mag, cos = report(y, ref)
    print(f"ViT-L width, batch 3: max-abs {err:.3e} (max|ref| {mag:.3f}) cosine {cos:.7f}")
    assert err <= 1.5e-2 * mag and cos >= 0.9999


def test_graphed_encoder_with_weight_prefetch(cuda_device, monkeypatch, samq_env):
    """The prefetch kernels are captured into the CUDA graph (programmatic edges behind the GEMMs):
    replays -- several, the scratch buffers rotate -- equal the eager pass without prefetch bit for bit."""
    from sam_quantization_b200 import ops
    from sam_quantization_b200.launcher import GraphedEncoder
    from sam_quantization_b200.synthetic import random_quantized_encoder

    monkeypatch.setattr(ops, "PREFETCH_MIN_M", 2048)
    enc = random_quantized_encoder("vit_b", 4, 128, seed=5, device=cuda_device, embed_dim=256, depth=3,
                                   num_heads=4, global_attn_indexes=(1,))
    g = torch.Generator().manual_seed(10)
    xs = [torch.randn(2, 3, 1024, 1024, generator=g).half().to(cuda_device) for _ in range(3)]
    with torch.no_grad():
        samq_env.set("SAMQ_PREFETCH", "0")
        plain = [enc(x).clone() for x in xs]
        samq_env.set("SAMQ_PREFETCH", "1")
        launches = []
        real = ops.qlinear_prefetch
        monkeypatch.setattr(ops, "qlinear_prefetch", lambda *a, **k: (launches.append(1), real(*a, **k))[1])
        ge = GraphedEncoder(enc, xs[0])
    assert launches, "the captured pass did not prefetch"
    for _ in range(2):
        for x, ref in zip(xs, plain):
            assert torch.equal(ge(x), ref)


def test_graphed_encoder_replay_and_pipelined_host_calls(cuda_device):
    """GraphedEncoder: the captured forward equals the eager one bit for bit, and run_host -- H2D
    of call i+1 / D2H of call i-1 on copy streams under the encoder of call i, double-buffered
    staging -- returns every call's own result in order, also when one output buffer per call is
    reused after its event."""
    from sam_quantization_b200.launcher import GraphedEncoder
    from sam_quantization_b200.synthetic import random_quantized_encoder

    enc = random_quantized_encoder("vit_b", 4, 128, seed=3, device=cuda_device, embed_dim=256, depth=2,
                                   num_heads=4, global_attn_indexes=(1,))
    g = torch.Generator().manual_seed(9)
    host_in = [torch.randn(2, 3, 1024, 1024, generator=g).half().pin_memory() for _ in range(5)]
    with torch.no_grad():
        eager = [enc(h.to(cuda_device)).cpu() for h in host_in]
    assert not torch.equal(eager[0], eager[1])
    ge = GraphedEncoder(enc, host_in[0].to(cuda_device))
    assert ge.kernels_per_replay > 0
    assert torch.equal(ge(host_in[1].to(cuda_device)).cpu(), eager[1])
    host_out = [torch.empty(2, 256, 64, 64, dtype=torch.float16).pin_memory() for _ in range(5)]
    events = [ge.run_host(host_in[i], host_out[i]) for i in range(5)]     # issued back to back
    for i, ev in enumerate(events):
        ev.synchronize()
        assert torch.equal(host_out[i], eager[i]), f"call {i}"
    # a second round through the same pipeline object, reusing ONE output buffer
    for i in (3, 0, 4):
        ge.run_host(host_in[i], host_out[0]).synchronize()
        assert torch.equal(host_out[0], eager[i])
    with pytest.raises(ValueError):
        ge.run_host(host_in[0].clone(), host_out[0])          # not pinned


def test_gptq_calibrated_encoder_on_the_cuda_path(cuda_device, tmp_path):
    """The whole offline + online chain inside this package: fp32 encoder -> encoder_sequential
    (GPTQ calibration on one synthetic image, CPU) -> encoder_pack -> save_quant -> load_quant on
    the GPU -> fused CUDA forward, against the fp32 forward of the ROUNDED encoder on the CPU."""
    import copy

    from sam_quantization_b200 import gptq as G

    torch.manual_seed(0)
    kw = dict(img_size=1024, patch_size=16, use_rel_pos=True, window_size=14, embed_dim=128, depth=2,
              num_heads=2, global_attn_indexes=(1,))
    enc = ie.ImageEncoderViT(**kw).float().eval()
    img = torch.from_numpy(synth.image(1, 1024, seed=2))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    qs = G.encoder_sequential(enc, [img.float()], wbits=4, nsamples=1, groupsize=64)
    rounded = copy.deepcopy(enc)
    with torch.no_grad():
        ref = rounded(img.half().float())
    G.encoder_pack(enc, qs, 4, 64)
    sq.save_quant(enc, str(tmp_path), 4, 64)
    enc2 = sq.load_quant(ie.ImageEncoderViT(**kw).half(), str(tmp_path), warmup_autotune=False,
                         device=cuda_device).eval()
    assert all(b._fused_ready() for b in enc2.blocks)
    with torch.no_grad():
        y = enc2(img.half().to(cuda_device))
    assert y.shape == ref.shape == (1, 256, 64, 64)
    err, mag, cos = report(y, ref)
    assert err <= 6e-2 * max(1.0, mag) and cos >= 0.999, (err, mag, cos)


def _vith_d2_q4(golden_dir, tmp_path, device, fixture="encoder_vith_d2_q4.npz", embed_dim=1280):
    g = np.load(os.path.join(golden_dir, fixture))
    cfg = dict(embed_dim=embed_dim, depth=2, num_heads=16, global_attn_indexes=(1,))
    p = synth.fp_state(seed=int(g["seed"]), **cfg)
    rng = np.random.default_rng(int(g["relpos_seed"]))
    for k in p:
        if "rel_pos" in k:
            p[k] = (rng.standard_normal(p[k].shape) * 0.2).astype(np.float32)
    state = synth.to_torch(synth.quantize_state(p, 4, 128))
    enc = ie.ImageEncoderViT(img_size=1024, patch_size=16, use_rel_pos=True, window_size=14, **cfg)
    sq.make_quant(enc, 4, 128)
    enc = enc.half()
    enc.load_state_dict(state, strict=False)
    sq.make_quant_attn(enc)
    sq.make_fused_mlp(enc)
    return g, enc.to(device).eval(), state


def test_vith_width_encoder_against_the_reference_modules_output(cuda_device, golden_dir, tmp_path):
    """CUDA path vs the REFERENCE's own ImageEncoderViT output (tests/golden/encoder_vith_d2_q4.npz,
    made by make_golden.py::make_encoder_q4_fixture: ViT-H width 1280 / 16 heads, block 0 windowed,
    block 1 global, batch 1 -- the only batch the reference's partition accepts -- fp32, with the
    dequantised int4-g128 weights in its nn.Linears).  Here the same seeds give the same packed
    weights, which run through the fused kernels in fp16.  Tolerance: block outputs
    <= 1.5e-2 max|ref| (+ cosine >= 0.9999), embeddings max-abs <= 6e-2 and cosine >= 0.999."""
    g, enc, _ = _vith_d2_q4(golden_dir, tmp_path, cuda_device)
    img = torch.from_numpy(synth.image(1, 1024, seed=int(g["seed"]))).half().to(cuda_device)
    grabbed = {}
    hooks = [enc.blocks[i].register_forward_hook(lambda _m, _i, o, i=i: grabbed.__setitem__(i, o.detach().float().cpu()))
             for i in range(2)]
    with torch.no_grad():
        y = enc(img).float().cpu()
    for h in hooks:
        h.remove()
    for i in range(2):
        sub = grabbed[i][:, ::4, ::4, ::8]
        ref = torch.from_numpy(g[f"tok{i}_sub"])
        err = (sub - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(sub.flatten().double(), ref.flatten().double(), dim=0).item()
        print(f"block {i}: max-abs {err:.3e} (max|ref| {float(g[f'tok{i}_absmax']):.2f}) cosine {cos:.7f}")
        assert err <= 1.5e-2 * float(g[f"tok{i}_absmax"]) and cos >= 0.9999
    ref = torch.from_numpy(g["y_sub"])
    sub = y[:, :, ::4, ::4]
    err = (sub - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(sub.flatten().double(), ref.flatten().double(), dim=0).item()
    print(f"embedding: max-abs {err:.3e} (max|ref| {float(g['y_absmax']):.2f}) cosine {cos:.7f}")
    assert err <= 6e-2 and cos >= 0.999
    assert abs(float(y.mean()) - float(g["y_mean"])) < 2e-3


def test_vitl_width_batch2_encoder_against_the_reference_modules_output(cuda_device, golden_dir, tmp_path):
    """What the reference's hard-coded partition cannot run is pinned to the reference all the same: its
    ImageEncoderViT with its OWN generic window_partition / window_unpartition (fq_vit/models/sam/
    image_encoder.py:481-537) patched in, ViT-L width (1024, heads of 64), batch 2, dequantised int4-g128
    weights (tests/golden/encoder_vitl_d2_b2_q4.npz).  Same tolerances as the ViT-H fixture test."""
    g, enc, _ = _vith_d2_q4(golden_dir, tmp_path, cuda_device, "encoder_vitl_d2_b2_q4.npz", 1024)
    img = torch.from_numpy(synth.image(2, 1024, seed=int(g["seed"]))).half().to(cuda_device)
    grabbed = {}
    hooks = [enc.blocks[i].register_forward_hook(lambda _m, _i, o, i=i: grabbed.__setitem__(i, o.detach().float().cpu()))
             for i in range(2)]
    with torch.no_grad():
        y = enc(img).float().cpu()
    for h in hooks:
        h.remove()
    for i in range(2):
        sub, ref = grabbed[i][:, ::4, ::4, ::8], torch.from_numpy(g[f"tok{i}_sub"])
        err = (sub - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(sub.flatten().double(), ref.flatten().double(), dim=0).item()
        print(f"ViT-L batch 2, block {i}: max-abs {err:.3e} (max|ref| {float(g[f'tok{i}_absmax']):.2f}) cosine {cos:.7f}")
        assert err <= 1.5e-2 * float(g[f"tok{i}_absmax"]) and cos >= 0.9999
    sub, ref = y[:, :, ::4, ::4], torch.from_numpy(g["y_sub"])
    err = (sub - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(sub.flatten().double(), ref.flatten().double(), dim=0).item()
    assert err <= 6e-2 and cos >= 0.999, (err, cos)


def test_vith_width_blocks_at_the_benchmarked_batch(cuda_device, golden_dir, tmp_path):
    """The benchmark's shapes: width 1280, 16 heads of 80, batch 32 (M = 131072 GEMM rows, 800
    windows, 32 global images), two blocks.  The oracle checks three images of the batch (CPU time);
    the rest is covered by the size-independent property that no op crosses the batch: a slice of
    four images run alone (same kernel path) gives identical bits."""
    g, enc, state = _vith_d2_q4(golden_dir, tmp_path, cuda_device)
    x = torch.from_numpy(synth.tokens_input(32, 64, 1280, seed=11)).half()
    ref_state = oe.dequant_state(state, 4, 128)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        y = enc.forward_tokens(x.to(cuda_device))
        for lo in (0, 28):
            assert torch.equal(y[lo:lo + 4], enc.forward_tokens(x[lo:lo + 4].to(cuda_device)))
        pick = [0, 13, 31]
        ref = oe.tokens_forward(x[pick].float(), ref_state, 2, 16, 14, (1,), "reference")
    err, mag, cos = report(y[pick], ref)
    print(f"ViT-H width, batch 32: max-abs {err:.3e} (max|ref| {mag:.3f}) cosine {cos:.7f}")
    assert err <= 1.5e-2 * mag and cos >= 0.9999
