"""Generate the golden fixtures in this directory by running the REFERENCE's own code.

Run once in the build container (needs /root/reference, CPU only):

    python tests/golden/make_golden.py

It imports / executes, unmodified and in place:
  * ``pack_linear``       AST-extracted from /root/reference/gptq4sam.py:434-497 (the module
                          itself imports albumentations, which is absent)
  * ``Quantizer``, ``quantize``   /root/reference/gptq.py:183-299
  * the literal dequant expression of ``matmul4_kernel``
                          (/root/reference/gptq_triton/quant_linear.py:296-301, 312-313, 334-339)
                          evaluated by torch on CPU tensors (gptq_triton itself cannot be
                          imported without a GPU: CUDA allocation at quant_linear.py:13)
  * ``Attention``, ``Block``, ``ImageEncoderViT``, ``window_partition``, ``window_unpartition``
                          /root/reference/segment_anything/modeling/image_encoder.py
and stores inputs + outputs as small .npz files.  Nothing here is copied into the
repository; the fixtures travel to the GPU box, /root/reference does not.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import synth  # noqa: E402  (input generators only)


def extract_pack_linear():
    src = open(os.path.join(REF, "gptq4sam.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "pack_linear")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "Optional": __import__("typing").Optional}
    exec(compile(mod, "gptq4sam.py::pack_linear", "exec"), ns)
    return ns["pack_linear"]


def make_pack_fixtures():
    from gptq import Quantizer, quantize  # reference

    pack_linear = extract_pack_linear()
    rng = np.random.default_rng(7)
    N, K, gs = 64, 256, 128
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    W[3, :128] = np.abs(W[3, :128])          # a whole group >= 0 -> zero == 0 -> the zeros-1 quirk (trap 8)
    for bits in (2, 4, 8):
        w16 = torch.from_numpy(W).half()
        G = K // gs
        scale = torch.zeros(N, G)
        zero = torch.zeros(N, G)
        wfake = torch.zeros(N, K)
        for g in range(G):
            q = Quantizer()
            q.configure(bits, perchannel=True, sym=False, mse=False)
            blk = w16[:, g * gs:(g + 1) * gs].float()
            q.find_params(blk, weight=True)
            scale[:, g] = q.scale.flatten()
            zero[:, g] = q.zero.flatten()
            wfake[:, g * gs:(g + 1) * gs] = quantize(blk, q.scale, q.zero, q.maxq)
        quant = types.SimpleNamespace(
            bits=bits, groupsize=gs, infeatures=K,
            qweight=torch.zeros((K * bits // 32, N), dtype=torch.int32),
            qzeros=torch.zeros((G, N * bits // 32), dtype=torch.int32),
            scales=None, bias=torch.zeros(N, dtype=torch.float16))
        bias = torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float32))
        pack_linear(quant, wfake.half(), scale, zero, bias)
        np.savez_compressed(
            os.path.join(HERE, f"pack_b{bits}.npz"), weight16=w16.numpy(), wfake16=wfake.half().numpy(),
            scale=scale.numpy(), zero=zero.numpy(), bias=bias.numpy(), qweight=quant.qweight.numpy(),
            qzeros=quant.qzeros.numpy(), scales=quant.scales.numpy(), qbias=quant.bias.numpy(),
            groupsize=gs)
        print("pack fixture bits", bits, "zero==0 groups:", int((zero == 0).sum()))


def make_dequant_fixture():
    """The kernel's expression evaluated literally by torch (CPU):  int32 * fp16 -> fp16."""
    rng = np.random.default_rng(11)
    K, N, gs = 256, 64, 128
    qweight = torch.from_numpy(rng.integers(-2**31, 2**31, size=(K // 8, N), dtype=np.int64).astype(np.int32))
    qzeros = torch.from_numpy(rng.integers(-2**31, 2**31, size=(K // gs, N // 8), dtype=np.int64).astype(np.int32))
    scales = torch.from_numpy(rng.uniform(1e-4, 2e-2, size=(K // gs, N)).astype(np.float16))
    offs_k = torch.arange(K)
    offs_n = torch.arange(N)
    shifter = (offs_k % 8) * 4                      # quant_linear.py:300
    zeros_shifter = (offs_n % 8) * 4                # quant_linear.py:301
    b = qweight[offs_k // 8, :]                     # :291-294 (each word repeated 8x along k)
    out = torch.empty(K, N, dtype=torch.float16)
    for g in range(K // gs):
        sc = scales[g]                              # :327
        zeros = qzeros[g][offs_n // 8]              # :329-331
        zeros = (zeros >> zeros_shifter) & 0xF      # :334
        zeros = (zeros + 1) * sc                    # :335   int32 * fp16 -> fp16
        rows = slice(g * gs, (g + 1) * gs)
        bb = (b[rows] >> shifter[rows, None]) & 0xF  # :338
        out[rows] = bb * sc[None, :] - zeros[None, :]  # :339
    assert out.dtype == torch.float16
    np.savez_compressed(os.path.join(HERE, "dequant_b4.npz"), qweight=qweight.numpy(), qzeros=qzeros.numpy(),
                        scales=scales.numpy(), w=out.numpy(), groupsize=gs)
    print("dequant fixture", out.shape)


def load_ref_state(module, p):
    sd = {k: torch.from_numpy(v) for k, v in p.items()}
    missing, unexpected = module.load_state_dict(sd, strict=True), None
    return missing, unexpected


def make_attention_fixture():
    from segment_anything.modeling.image_encoder import Attention  # reference

    for name, (dim, heads, size, B) in {"attn_win": (128, 2, 14, 3), "attn_glob": (160, 2, 64, 1)}.items():
        rng = np.random.default_rng(13)
        m = Attention(dim, num_heads=heads, qkv_bias=True, use_rel_pos=True, input_size=(size, size))
        hd = dim // heads
        p = {
            "qkv.weight": (rng.standard_normal((3 * dim, dim)) * 0.05).astype(np.float32),
            "qkv.bias": (rng.standard_normal(3 * dim) * 0.05).astype(np.float32),
            "proj.weight": (rng.standard_normal((dim, dim)) * 0.05).astype(np.float32),
            "proj.bias": (rng.standard_normal(dim) * 0.05).astype(np.float32),
            "rel_pos_h": (rng.standard_normal((2 * size - 1, hd)) * 0.3).astype(np.float32),
            "rel_pos_w": (rng.standard_normal((2 * size - 1, hd)) * 0.3).astype(np.float32),
        }
        load_ref_state(m, p)
        x16 = rng.standard_normal((B, size, size, dim)).astype(np.float16)   # fp16-representable input
        with torch.no_grad():
            y = m(torch.from_numpy(x16.astype(np.float32))).numpy()
        stride = 4 if size > 14 else 1                                       # keep the fixture small
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), x16=x16, y_sub=y.reshape(B, size * size, dim)[:, ::stride],
                            stride=stride, heads=heads, **p)
        print("attention fixture", name, y.shape)


def make_partition_fixture():
    """The fork's hard-coded partition/unpartition (ViT-H, batch 1) on a thin tensor is
    impossible (C is hard-coded to 1280), so use C=1280 but store only a channel subset."""
    from segment_anything.modeling.image_encoder import window_partition, window_unpartition

    rng = np.random.default_rng(17)
    x = rng.standard_normal((1, 64, 64, 1280)).astype(np.float32)
    w, pad_hw = window_partition(torch.from_numpy(x), 14)
    back = window_unpartition(w, 14, pad_hw, (64, 64))
    ch = np.arange(0, 1280, 160)
    np.savez_compressed(os.path.join(HERE, "partition_vith.npz"), x=x[..., ch], windows=w.numpy()[..., ch],
                        back=back.numpy()[..., ch], pad_hw=np.array(pad_hw))
    print("partition fixture", w.shape, pad_hw)


def make_encoder_fixture():
    """Reference ImageEncoderViT at ViT-H width (the only width its hard-coded partition
    accepts), depth 2 (block 0 windowed, block 1 global), batch 1, fp32, unpatched."""
    from segment_anything.modeling.image_encoder import ImageEncoderViT
    from functools import partial

    cfg = dict(embed_dim=1280, depth=2, num_heads=16, global_attn_indexes=(1,))
    enc = ImageEncoderViT(img_size=1024, patch_size=16, embed_dim=1280, depth=2, num_heads=16, mlp_ratio=4,
                          out_chans=256, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          use_rel_pos=True, window_size=14, global_attn_indexes=(1,))
    p = synth.fp_state(seed=5, **cfg)
    # make the rel-pos path matter
    rng = np.random.default_rng(19)
    for k in p:
        if "rel_pos" in k:
            p[k] = (rng.standard_normal(p[k].shape) * 0.2).astype(np.float32)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=True)
    img = synth.image(1, 1024, seed=5)
    with torch.no_grad():
        y = enc(torch.from_numpy(img)).numpy()
    np.savez_compressed(os.path.join(HERE, "encoder_vith_d2.npz"), y_sub=y[:, :, ::4, ::4].astype(np.float32),
                        y_mean=np.float64(y.mean()), y_absmax=np.float64(np.abs(y).max()), seed=5, relpos_seed=19)
    print("encoder fixture", y.shape, float(np.abs(y).max()))


def make_encoder_q4_fixture():
    """The benchmarked configuration's numerics at ViT-H width: reference ImageEncoderViT (depth 2,
    batch 1, fp32, unpatched) whose block Linears hold the DEQUANTISED int4-g128 weights -- the
    reference's PyTorch dequant path -- on the image the CUDA test feeds
    through the packed weights.  Everything is derived from seeds, so only outputs are stored."""
    from functools import partial

    from oracle import encoder as oe
    from segment_anything.modeling.image_encoder import ImageEncoderViT

    cfg = dict(embed_dim=1280, depth=2, num_heads=16, global_attn_indexes=(1,))
    enc = ImageEncoderViT(img_size=1024, patch_size=16, embed_dim=1280, depth=2, num_heads=16, mlp_ratio=4,
                          out_chans=256, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          use_rel_pos=True, window_size=14, global_attn_indexes=(1,))
    p = synth.fp_state(seed=7, **cfg)
    rng = np.random.default_rng(23)
    for k in p:
        if "rel_pos" in k:
            p[k] = (rng.standard_normal(p[k].shape) * 0.2).astype(np.float32)
    state = oe.dequant_state(synth.to_torch(synth.quantize_state(p, 4, 128)), 4, 128)
    enc.load_state_dict(state, strict=True)
    img = synth.image(1, 1024, seed=7).astype(np.float16).astype(np.float32)
    grabbed = {}
    hooks = [enc.blocks[i].register_forward_hook(lambda _m, _i, o, i=i: grabbed.__setitem__(i, o.detach().numpy()))
             for i in range(2)]
    with torch.no_grad():
        y = enc(torch.from_numpy(img)).numpy()
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(HERE, "encoder_vith_d2_q4.npz"), y_sub=y[:, :, ::4, ::4].astype(np.float32),
                        y_mean=np.float64(y.mean()), y_absmax=np.float64(np.abs(y).max()),
                        tok0_sub=grabbed[0][:, ::4, ::4, ::8].astype(np.float32), tok0_absmax=np.float64(np.abs(grabbed[0]).max()),
                        tok1_sub=grabbed[1][:, ::4, ::4, ::8].astype(np.float32), tok1_absmax=np.float64(np.abs(grabbed[1]).max()),
                        seed=7, relpos_seed=23, bits=4, groupsize=128)
    print("encoder q4 fixture", y.shape, float(np.abs(y).max()))


def extract_generic_partition():
    """The reference's own GENERIC window_partition / window_unpartition (fq_vit/models/sam/image_encoder.py:
    481-537; the copies in segment_anything are hard-coded to ViT-H batch 1), AST-extracted because the fq_vit
    package imports quantisation tooling that is not installed."""
    src = open(os.path.join(REF, "fq_vit/models/sam/image_encoder.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": torch.nn.functional, "Tuple": __import__("typing").Tuple}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("window_partition", "window_unpartition"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "fq_vit_image_encoder", "exec"), ns)
    return ns["window_partition"], ns["window_unpartition"]


def make_encoder_vitl_b2_fixture():
    """Pins what the hard-coded partition cannot: another width and batch > 1.  Reference ImageEncoderViT at
    ViT-L width (1024, 16 heads of 64), depth 2 (windowed + global), BATCH 2, fp32, with the reference's own
    generic partition functions patched in and dequantised int4-g128 weights in its nn.Linears."""
    from functools import partial

    import segment_anything.modeling.image_encoder as rie
    from oracle import encoder as oe

    part, unpart = extract_generic_partition()
    saved = rie.window_partition, rie.window_unpartition
    rie.window_partition, rie.window_unpartition = part, unpart
    try:
        cfg = dict(embed_dim=1024, depth=2, num_heads=16, global_attn_indexes=(1,))
        enc = rie.ImageEncoderViT(img_size=1024, patch_size=16, embed_dim=1024, depth=2, num_heads=16, mlp_ratio=4,
                                  out_chans=256, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                                  use_rel_pos=True, window_size=14, global_attn_indexes=(1,))
        p = synth.fp_state(seed=9, **cfg)
        rng = np.random.default_rng(29)
        for k in p:
            if "rel_pos" in k:
                p[k] = (rng.standard_normal(p[k].shape) * 0.2).astype(np.float32)
        state = oe.dequant_state(synth.to_torch(synth.quantize_state(p, 4, 128)), 4, 128)
        enc.load_state_dict(state, strict=True)
        img = synth.image(2, 1024, seed=9).astype(np.float16).astype(np.float32)
        grabbed = {}
        hooks = [enc.blocks[i].register_forward_hook(lambda _m, _i, o, i=i: grabbed.__setitem__(i, o.detach().numpy()))
                 for i in range(2)]
        with torch.no_grad():
            y = enc(torch.from_numpy(img)).numpy()
        for h in hooks:
            h.remove()
    finally:
        rie.window_partition, rie.window_unpartition = saved
    np.savez_compressed(os.path.join(HERE, "encoder_vitl_d2_b2_q4.npz"), y_sub=y[:, :, ::4, ::4].astype(np.float32),
                        y_mean=np.float64(y.mean()), y_absmax=np.float64(np.abs(y).max()),
                        tok0_sub=grabbed[0][:, ::4, ::4, ::8].astype(np.float32), tok0_absmax=np.float64(np.abs(grabbed[0]).max()),
                        tok1_sub=grabbed[1][:, ::4, ::4, ::8].astype(np.float32), tok1_absmax=np.float64(np.abs(grabbed[1]).max()),
                        seed=9, relpos_seed=29, bits=4, groupsize=128, batch=2)
    print("encoder ViT-L batch-2 fixture", y.shape, float(np.abs(y).max()))


def make_relpos_fixture():
    """Reference get_rel_pos (image_encoder.py:336-366) incl. its interpolation branch (:348-358): tables
    whose length differs from 2 * max(q, k) - 1, and non-square q / k."""
    from segment_anything.modeling.image_encoder import get_rel_pos

    out = {}
    cases = [(14, 14, 27), (14, 14, 127), (64, 64, 27), (7, 14, 27), (14, 7, 13), (10, 10, 33)]
    for n, (q, k, L) in enumerate(cases):
        t = torch.from_numpy(np.random.default_rng(40 + n).standard_normal((L, 8)).astype(np.float32))
        out[f"c{n}_table"] = t.numpy()
        out[f"c{n}_R"] = get_rel_pos(q, k, t).numpy()
    np.savez_compressed(os.path.join(HERE, "rel_pos_interp.npz"), cases=np.array(cases), **out)
    print("rel-pos fixture", len(cases))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "encoder_q4":
        make_encoder_q4_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vitl_b2":
        make_encoder_vitl_b2_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "relpos":
        make_relpos_fixture()
        sys.exit(0)
    torch.manual_seed(0)
    make_pack_fixtures()
    make_dequant_fixture()
    make_attention_fixture()
    make_partition_fixture()
    make_encoder_fixture()
    make_encoder_q4_fixture()
    make_relpos_fixture()
    make_encoder_vitl_b2_fixture()
    print("done")
