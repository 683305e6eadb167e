"""Host-side logic of the drop-in modules on the CPU: constructor contract, pack(),
module swapping, checkpoint layout, and the 'no CPU fallback' rule."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import sam_quantization_b200 as sq
from sam_quantization_b200 import image_encoder as ie
from sam_quantization_b200.quant_linear import pack_fields
from oracle import encoder as oe
from oracle import quant as oq
from oracle import synth


def test_constructor_contract_matches_reference():
    """quant_linear.py:66-110: buffer names, dtypes, shapes; groupsize -1 -> K."""
    m = sq.QuantLinear(4, 128, 1280, 3840, True)
    assert m.qweight.shape == (160, 3840) and m.qweight.dtype == torch.int32
    assert m.qzeros.shape == (10, 480) and m.qzeros.dtype == torch.int32
    assert m.scales.shape == (10, 3840) and m.scales.dtype == torch.float16
    assert m.bias.shape == (3840,) and m.bias.dtype == torch.float16
    assert set(m.state_dict()) == {"qweight", "qzeros", "scales", "bias"}
    m = sq.QuantLinear(4, -1, 256, 512, False)
    assert m.groupsize == 256 and m.qzeros.shape == (1, 64) and m.bias is None
    assert sq.QuantLinear(3, 128, 256, 512, False).qweight.shape == (24, 512)
    assert sq.QuantLinear(8, 128, 256, 512, False).qweight.shape == (64, 512)
    with pytest.raises(NotImplementedError):
        sq.QuantLinear(5, 128, 256, 512, False)
    with pytest.raises(AssertionError):
        sq.QuantLinear(4, 128, 256, 10, False)        # quant_linear.py:84-86


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_pack_bit_identical_to_reference_pack_linear(golden_dir, bits):
    g = np.load(os.path.join(golden_dir, f"pack_b{bits}.npz"))
    n, k = g["wfake16"].shape
    lin = nn.Linear(k, n)
    lin.weight.data = torch.from_numpy(g["wfake16"])
    lin.bias.data = torch.from_numpy(g["bias"])
    m = sq.QuantLinear(bits, int(g["groupsize"]), k, n, True)
    m.pack(lin, torch.from_numpy(g["scale"]), torch.from_numpy(g["zero"]))
    assert torch.equal(m.qweight, torch.from_numpy(g["qweight"]))
    assert torch.equal(m.qzeros, torch.from_numpy(g["qzeros"]))
    assert torch.equal(m.scales.view(torch.int16), torch.from_numpy(g["scales"]).view(torch.int16))
    assert torch.equal(m.bias.view(torch.int16), torch.from_numpy(g["qbias"]).view(torch.int16))


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_pack_fields_equals_oracle(bits):
    rng = np.random.default_rng(bits)
    vals = rng.integers(0, 2**bits, size=(64, 5))
    a = pack_fields(torch.from_numpy(vals), bits).numpy()
    b = oq._pack_fields(vals, bits)
    assert np.array_equal(a, b)


def test_pack_with_g_idx_round_trip():
    rng = np.random.default_rng(3)
    n, k, gs, bits = 64, 256, 64, 4
    p = synth.fp_state(embed_dim=64, depth=1, num_heads=1, global_attn_indexes=(), with_stem=False, seed=1)
    w = (rng.standard_normal((n, k)) * 0.02).astype(np.float32)
    perm = rng.permutation(k)
    inv = np.empty(k, dtype=np.int64); inv[perm] = np.arange(k)
    g_idx = (inv // gs).astype(np.int32)
    wp, scale, zero = oq.rtn_quantize(w[:, perm], bits, gs)
    wf = np.empty_like(wp); wf[:, perm] = wp
    m = sq.QuantLinear(bits, gs, k, n, False)
    m.pack(torch.from_numpy(wf), torch.from_numpy(scale), torch.from_numpy(zero), torch.from_numpy(g_idx))
    ref = oq.pack(wf, scale, zero, bits, gs, g_idx)
    assert np.array_equal(m.qweight.numpy(), ref["qweight"]) and np.array_equal(m.qzeros.numpy(), ref["qzeros"])
    assert "g_idx" in m.state_dict() and m.g_idx.dtype == torch.int32


def test_no_cpu_fallback():
    m = sq.QuantLinear(4, 128, 128, 128, True)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        m(torch.zeros(4, 128, dtype=torch.float16))


def tiny_encoder(depth=2, dim=128, heads=2, glob=(1,)):
    return ie.ImageEncoderViT(img_size=1024, patch_size=16, embed_dim=dim, depth=depth, num_heads=heads,
                              use_rel_pos=True, window_size=14, global_attn_indexes=glob)


def test_module_tree_keys_match_reference_state_dict():
    enc = tiny_encoder()
    keys = set(enc.state_dict())
    p = synth.fp_state(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,))
    assert keys == set(p)
    assert enc.blocks[0].attn.rel_pos_h.shape == (27, 64) and enc.blocks[1].attn.rel_pos_h.shape == (127, 64)


def test_make_quant_and_swaps():
    enc = tiny_encoder()
    enc.lm_head = nn.Linear(4, 4)
    sq.make_quant(enc, 4, 64)
    assert isinstance(enc.lm_head, nn.Linear)                       # quant_linear.py:24-25
    assert isinstance(enc.blocks[0].attn.qkv, sq.QuantLinear) and isinstance(enc.blocks[1].mlp.lin2, sq.QuantLinear)
    assert not any(isinstance(m, nn.Linear) for n, m in enc.named_modules() if n != "lm_head")
    sq.make_quant_attn(enc)
    assert isinstance(enc.blocks[0].attn, sq.QuantAttention)
    assert enc.blocks[0].attn.qkv_proj.outfeatures == 384 and enc.blocks[0].attn.rel_pos_h.shape == (27, 64)
    sq.make_fused_mlp(enc)
    assert isinstance(enc.blocks[1].mlp, sq.QuantMLP)
    assert enc.blocks[0]._fused_ready()
    with pytest.raises(RuntimeError, match="no CPU"):
        enc.blocks[0](torch.zeros(1, 64, 64, 128))


def test_checkpoint_round_trip_reference_layout(tmp_path):
    """quant_config.json + model.pt with the reference's key names; zero biases elided
    (gptq_triton/__init__.py:33-60)."""
    cfg = dict(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,))
    p = synth.fp_state(seed=2, **cfg)
    p["blocks.0.attn.proj.bias"][:] = 0                               # must become None on load
    packed = synth.to_torch(synth.quantize_state(p, 4, 64))
    enc = tiny_encoder()
    sq.make_quant(enc, 4, 64)
    enc.half()
    enc.load_state_dict(packed, strict=True)
    sq.save_quant(enc, str(tmp_path), 4, 64)
    assert json.load(open(tmp_path / "quant_config.json")) == {"wbits": 4, "groupsize": 64}
    saved = torch.load(tmp_path / "model.pt")
    assert "blocks.0.attn.qkv.qweight" in saved and "blocks.1.mlp.lin1.scales" in saved

    class Wrapper(nn.Module):          # like Sam: the encoder is a sub-module
        def __init__(self):
            super().__init__()
            self.image_encoder = tiny_encoder()
    # reference layout has the sub-module prefix in the keys
    torch.save({"image_encoder." + k: v for k, v in saved.items()}, tmp_path / "model.pt")
    model = Wrapper().half()
    out = sq.load_quant(model, str(tmp_path), warmup_autotune=False, device=None, sub_module="image_encoder")
    blk = out.image_encoder.blocks[0]
    assert isinstance(blk.attn, sq.QuantAttention) and isinstance(blk.mlp, sq.QuantMLP)
    assert blk.attn.o_proj.bias is None and blk.attn.qkv_proj.bias is not None
    assert torch.equal(blk.attn.qkv_proj.qweight, packed["blocks.0.attn.qkv.qweight"])
    with pytest.raises(FileNotFoundError):
        os.remove(tmp_path / "model.pt")
        sq.load_quant(Wrapper().half(), str(tmp_path), warmup_autotune=False, device=None, sub_module="image_encoder")
    with pytest.raises(ValueError):
        torch.save({}, tmp_path / "model.pt")
        sq.load_quant(Wrapper().half(), str(tmp_path), warmup_autotune=True, device=None, sub_module="image_encoder")


def test_checkpoint_safetensors_with_act_order_groups(tmp_path):
    """The safetensors form of the checkpoint directory (gptq_triton/__init__.py:33-44 loads it
    strictly) including the ``g_idx`` extension for act-order groups, and the reference's
    preference order: ``model.safetensors`` wins over a ``model.pt`` lying next to it."""
    cfg = dict(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,))
    p = synth.fp_state(seed=4, **cfg)
    p["blocks.1.mlp.lin2.bias"][:] = 0
    packed = synth.to_torch(synth.quantize_state(p, 4, 64, act_order_seed=9))
    assert any(k.endswith(".g_idx") for k in packed)
    enc = tiny_encoder()
    sq.make_quant(enc, 4, 64)
    enc.half()
    sq._register_g_idx(enc, packed)
    enc.load_state_dict(packed, strict=True)
    sq.save_quant(enc, str(tmp_path), 4, 64, safetensors=True)
    assert (tmp_path / "model.safetensors").exists() and not (tmp_path / "model.pt").exists()
    torch.save({"garbage": torch.zeros(1)}, tmp_path / "model.pt")          # must be ignored
    out = sq.load_quant(tiny_encoder().half(), str(tmp_path), warmup_autotune=False, device=None)
    qkv = out.blocks[0].attn.qkv_proj
    assert qkv.g_idx is not None and qkv.g_idx.dtype == torch.int32
    assert torch.equal(qkv.g_idx, packed["blocks.0.attn.qkv.g_idx"].to(torch.int32))
    assert torch.equal(qkv.qweight, packed["blocks.0.attn.qkv.qweight"])
    assert out.blocks[1].mlp.lin2.bias is None and out.blocks[1].mlp.lin1.bias is not None
    # strict: a missing tensor in the safetensors file is an error (the .pt form is lenient)
    from safetensors.torch import load_file, save_file
    st = load_file(str(tmp_path / "model.safetensors"))
    st.pop("blocks.0.attn.qkv.scales")
    save_file(st, str(tmp_path / "model.safetensors"))
    with pytest.raises(RuntimeError):
        sq.load_quant(tiny_encoder().half(), str(tmp_path), warmup_autotune=False, device=None)


def test_eager_encoder_equals_oracle_fp32():
    """The pre-quantization host module (generic partition, both rel_w modes) vs the oracle."""
    cfg = dict(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,))
    p = synth.fp_state(seed=4, **cfg)
    for k in p:
        if "rel_pos" in k:
            p[k] = p[k] * 10
    enc = tiny_encoder()
    enc.load_state_dict(synth.to_torch(p), strict=True)
    x = torch.from_numpy(synth.tokens_input(2, 64, 128, seed=1))
    with torch.no_grad():
        y = enc.forward_tokens(x)
        ref = oe.tokens_forward(x, synth.to_torch(p), 2, 2, 14, (1,), "reference")
        assert torch.max(torch.abs(y - ref)) < 1e-4
        for blk in enc.blocks:
            blk.attn.relw_mode = "upstream"
        y2 = enc.forward_tokens(x)
        ref2 = oe.tokens_forward(x, synth.to_torch(p), 2, 2, 14, (1,), "upstream")
        assert torch.max(torch.abs(y2 - ref2)) < 1e-4
        assert torch.max(torch.abs(y - y2)) > 1e-4


def test_matmul4_reference_assertions():
    a = torch.zeros(4, 100, dtype=torch.float16)
    with pytest.raises(AssertionError):
        sq.triton_matmul4(128, a, torch.zeros(16, 256, dtype=torch.int32), torch.zeros(1, 256, dtype=torch.float16),
                          torch.zeros(1, 32, dtype=torch.int32))


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_unpack_fields_inverts_pack_fields(bits):
    from sam_quantization_b200.quant_linear import unpack_fields

    g = torch.Generator().manual_seed(bits)
    vals = torch.randint(0, 2 ** bits, (96, 40), generator=g)
    assert torch.equal(unpack_fields(pack_fields(vals, bits), bits), vals)


@pytest.mark.parametrize("bits", [3, 4, 8])
def test_sorted_pack_of_an_act_order_layer_is_the_same_weight(bits):
    """QuantLinear.sorted_pack: rows of qweight re-ordered so that groups are contiguous.  Checked
    with the oracle: dequant(sorted, contiguous groups) == dequant(original, g_idx)[perm], bit for bit."""
    K, N, gs = 256, 64, 64
    rng = np.random.default_rng(bits)
    layer = sq.QuantLinear(bits, gs, K, N, bias=False)
    qw = rng.integers(-2**31, 2**31, size=tuple(layer.qweight.shape), dtype=np.int64).astype(np.int32)
    qz = rng.integers(-2**31, 2**31, size=tuple(layer.qzeros.shape), dtype=np.int64).astype(np.int32)
    sc = rng.uniform(0.002, 0.02, size=tuple(layer.scales.shape)).astype(np.float16)
    perm0 = rng.permutation(K)
    inv = np.empty(K, dtype=np.int64)
    inv[perm0] = np.arange(K)
    gi = (inv // gs).astype(np.int32)
    layer.qweight, layer.qzeros, layer.scales = torch.from_numpy(qw), torch.from_numpy(qz), torch.from_numpy(sc)
    assert layer.sorted_pack() is None                       # no g_idx: nothing to sort
    layer.g_idx = torch.from_numpy(gi)
    perm, qw_sorted = layer.sorted_pack()
    assert layer.sorted_pack()[1] is qw_sorted               # cached
    w = oq.dequant(qw, qz, sc, bits, gs, gi)
    w_sorted = oq.dequant(qw_sorted.numpy(), qz, sc, bits, gs)
    assert np.array_equal(w_sorted.view(np.uint16), w[perm.numpy()].view(np.uint16))
    assert np.array_equal(np.sort(gi[perm.numpy()]), gi[perm.numpy()])   # groups contiguous and ascending
    layer.g_idx = torch.from_numpy(np.minimum(gi, 1))        # unequal group sizes: not sortable into gs-blocks
    assert layer.sorted_pack() is None


def test_get_rel_pos_matches_reference_including_interpolation(golden_dir):
    """image_encoder.py:336-366 with its interpolation branch (:348-358), against the reference
    function's outputs (tests/golden/rel_pos_interp.npz)."""
    g = np.load(os.path.join(golden_dir, "rel_pos_interp.npz"))
    for n, (q, k, L) in enumerate(g["cases"]):
        R = ie.get_rel_pos(int(q), int(k), torch.from_numpy(g[f"c{n}_table"]))
        assert R.shape == g[f"c{n}_R"].shape
        assert np.allclose(R.numpy(), g[f"c{n}_R"], atol=1e-6), (q, k, L)
    # SAM's own case stays the plain gather
    t = torch.arange(27, dtype=torch.float32)[:, None].repeat(1, 2)
    assert torch.equal(ie.get_rel_pos(14, 14, t)[3, 5], t[3 - 5 + 13])


def test_prefetch_chain_state_does_not_travel_with_the_model():
    """The weight-prefetch chain (runtime state: layer order + scratch buffers) is dropped by
    deepcopy and pickle -- the copy links its own on first use -- and never enters state_dict."""
    import copy
    import pickle

    from sam_quantization_b200.quant_linear import QuantLinear, WeightPrefetchChain, _PREFETCH

    layers = [QuantLinear(4, 128, 128, 256, True), QuantLinear(4, 128, 256, 128, True)]
    holder = torch.nn.Sequential(*layers)
    holder.__dict__["_pf_chain"] = chain = WeightPrefetchChain(layers)
    assert _PREFETCH[layers[1]] == (chain, 1)
    assert copy.deepcopy(holder).__dict__["_pf_chain"] is None
    assert pickle.loads(pickle.dumps(holder)).__dict__["_pf_chain"] is None
    assert not any("pf" in k for k in holder.state_dict())
    assert chain.take(0) is None            # nothing prefetched: the layer unpacks its own weight
