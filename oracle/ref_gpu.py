"""Harness: run the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) on the GPU
box and record what the product is judged against.  TEST / MEASUREMENT INFRASTRUCTURE ONLY --
nothing in sam_quantization_b200/ imports this file, Triton or oracle/_ref.

    python oracle/ref_gpu.py [--out gpurun_out/ref_gpu] [--sections dequant,microbench,attention,encoder]

Sections:
  dequant     identity-matrix extraction of the reference kernel's dequantised weights
              (triton_matmul4(gs, I_K, qweight, scales, qzeros), quant_linear.py:355-437) and the
              mismatch count against each candidate rounding form of oracle/quant.py; writes the
              golden fixture dequant_triton_b4.npz (inputs + the Triton kernel's own output)
  microbench  BASELINE config 5: M x (K,N) sweep, reference Triton int4 vs ours (bits 4/3/2/8) vs
              cuBLAS fp16, CUDA events, L2 flushed between launches
  attention   fused_attention.test_op numbers on this GPU; reference rel-pos + _fwd_kernel1 vs our
              attention kernel on identical qkv / rel-pos tables (parity + time)
  encoder     ViT-H batch 1 with the reference's protocol (gptq4sam_infer.py:59-79: 25 warm-up +
              100 iterations, wall clock): reference eager fp16, reference quantized (Triton), ours
              (eager launches and CUDA-graph replay) on identical packed weights + output parity
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import make_ref  # noqa: E402
from oracle import quant as oq  # noqa: E402

RESULT = {}
OUT = None


def save():
    with open(OUT + ".json", "w") as f:
        json.dump(RESULT, f, indent=1, default=str)


def section(name):
    def deco(fn):
        def run(*a, **kw):
            t0 = time.time()
            try:
                RESULT[name] = fn(*a, **kw)
            except Exception:  # keep the other sections' results
                RESULT[name] = {"error": traceback.format_exc()}
                print(f"[{name}] FAILED\n{RESULT[name]['error']}", flush=True)
            RESULT.setdefault("_seconds", {})[name] = round(time.time() - t0, 1)
            save()
        return run
    return deco


def triton_api_shim():
    """The reference pins no Triton version (setup.py: install_requires=[]).  Triton >= 3.x hands the
    launch kwargs (grid=, warmup=) to ``early_config_prune``; the reference's pruner
    (gptq_triton/utils.py:5) takes only (configs, nargs).  The shim drops the extra kwargs at the call
    boundary -- the reference files themselves stay byte-identical."""
    import gptq_triton.quant_linear as rql

    kern = rql.matmul4_kernel
    orig = kern.early_config_prune
    if getattr(orig, "_samq_shim", False):
        return

    def prune(configs, nargs, **_ignored):
        return list(orig(configs, nargs))

    prune._samq_shim = True
    kern.early_config_prune = prune


def rand_packed(K, N, gs, seed, lo, hi, dev):
    rng = np.random.default_rng(seed)
    G = K // gs
    qweight = rng.integers(-2**31, 2**31, size=(K // 8, N), dtype=np.int64).astype(np.int32)
    qzeros = rng.integers(-2**31, 2**31, size=(G, N // 8), dtype=np.int64).astype(np.int32)
    scales = rng.uniform(lo, hi, size=(G, N)).astype(np.float16)
    t = lambda a: torch.from_numpy(a).to(dev)
    return (qweight, qzeros, scales), (t(qweight), t(qzeros), t(scales))


def bits16(a):
    return np.ascontiguousarray(a).view(np.uint16)


# ----------------------------------------------------------------------------------------------
@section("dequant")
def run_dequant(dev):
    import gptq_triton.quant_linear as rql

    out = {"cases": []}
    golden = {}
    cases = [  # (K, N, gs, scale lo, hi, tag)
        (1280, 3840, 128, 0.001, 0.004, "qkv_g128_small_scales"),
        (1280, 3840, 128, 0.002, 0.02, "qkv_g128_config5_scales"),
        (5120, 1280, 128, 0.002, 0.02, "lin2_g128"),
        (1280, 1280, 1280, 0.002, 0.02, "proj_nogroups"),
        (256, 512, 128, 0.002, 0.02, "golden_g128"),
        (256, 256, 256, 0.001, 0.05, "golden_nogroups"),
    ]
    for i, (K, N, gs, lo, hi, tag) in enumerate(cases):
        (qw, qz, sc), (dqw, dqz, dsc) = rand_packed(K, N, gs, 100 + i, lo, hi, dev)
        eye = torch.eye(K, dtype=torch.float16, device=dev)
        w_ref = rql.triton_matmul4(gs, eye, dqw, dsc, dqz).clone().cpu().numpy()      # [K, N] fp16
        # a second extraction with 2*I: every product doubles exactly -> same mantissas
        w_ref2 = rql.triton_matmul4(gs, 2 * eye, dqw, dsc, dqz).clone().cpu().numpy()
        rec = {"tag": tag, "K": K, "N": N, "groupsize": gs, "elements": int(K * N),
               "x2_consistent": bool(np.array_equal(bits16((w_ref.astype(np.float32) * 2).astype(np.float16)),
                                                    bits16(w_ref2)))}
        for form in ("stepwise", "fma", "single"):
            w = oq.dequant(qw, qz, sc, 4, gs, form=form)
            rec[f"mismatch_{form}"] = int((bits16(w) != bits16(w_ref)).sum())
        out["cases"].append(rec)
        print("[dequant]", rec, flush=True)
        if tag.startswith("golden"):
            golden[f"{tag}_qweight"], golden[f"{tag}_qzeros"], golden[f"{tag}_scales"] = qw, qz, sc
            golden[f"{tag}_w_triton"] = w_ref
            golden[f"{tag}_groupsize"] = np.int32(gs)
    np.savez_compressed(OUT + "_dequant_triton_b4.npz", **golden)
    # what the compiled kernel does: count the fp16 arithmetic opcodes in its PTX
    try:
        ptx_stats = {}
        fn = rql.matmul4_kernel
        jit = getattr(fn, "fn", fn)
        caches = getattr(jit, "device_caches", None) or {}
        for dkey, cache_tuple in caches.items():
            kcache = cache_tuple[0]
            for key, kern in kcache.items():
                ptx = kern.asm.get("ptx", "")
                stat = {op: ptx.count(op) for op in ("fma.rn.f16", "mul.f16", "mul.rn.f16", "sub.f16", "sub.rn.f16",
                                                     "fma.rn.f32", "mul.f32", "sub.f32", "cvt.rn.f16.s32",
                                                     "cvt.rn.f16.s16", "cvt.rn.f32.s32", "tcgen05.mma", "mma.sync",
                                                     "wgmma")}
                ptx_stats[str(hash(str(key)))[:8]] = {k: v for k, v in stat.items() if v}
                if "ptx_sample" not in out:
                    lines = [l for l in ptx.splitlines() if ("f16" in l and ("fma" in l or "mul" in l or "sub" in l))]
                    out["ptx_sample"] = lines[:12]
        out["ptx_opcode_counts_per_compiled_variant"] = ptx_stats
        with contextlib.suppress(Exception):
            out["best_configs"] = {str(k): str(v) for k, v in fn.cache.items()}
    except Exception:
        out["ptx_error"] = traceback.format_exc()
    return out


# ----------------------------------------------------------------------------------------------
def time_cold(fn, flush, iters=20, warm=5):
    """Per-launch CUDA events with the L2 flushed (a > L2 buffer rewritten) before each launch."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def time_loop(fn, iters=20, warm=5):
    """Back-to-back launches between one event pair (us per launch, operands L2-warm when small)."""
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) * 1e3 / iters


@section("microbench")
def run_microbench(dev, peaks):
    import gptq_triton.quant_linear as rql
    from sam_quantization_b200 import _lib, ops

    rql.workspace = torch.empty(32768 * 5120, dtype=torch.float16, device=dev)   # SURVEY trap 6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for (K, N) in ((1280, 3840), (1280, 5120), (5120, 1280)):
        for M in (196, 4096, 32768):
            flops = 2.0 * M * K * N
            x = torch.randn(M, K, device=dev, dtype=torch.float16)
            bias = (torch.randn(N, device=dev) * 0.02).half()
            _, (qw, qz, sc) = rand_packed(K, N, 128, 7, 0.002, 0.02, dev)
            wd = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
            rec = {"M": M, "K": K, "N": N, "gflop": flops / 1e9}

            def tf(us):
                return round(flops / us / 1e6, 1)

            with torch.no_grad():
                t = time_cold(lambda: rql.triton_matmul4(128, x, qw, sc, qz), flush)
                rec["ref_triton_int4_us"], rec["ref_triton_int4_tflops"] = round(t, 2), tf(t)
                t = time_cold(lambda: rql.triton_matmul4(128, x, qw, sc, qz, bias), flush)
                rec["ref_triton_int4_bias_us"], rec["ref_triton_int4_bias_tflops"] = round(t, 2), tf(t)
                t = time_loop(lambda: rql.triton_matmul4(128, x, qw, sc, qz, bias))
                rec["ref_triton_int4_bias_loop_us"] = round(t, 2)
                t = time_cold(lambda: torch.nn.functional.linear(x, wd, bias), flush)
                rec["cublas_fp16_bias_us"], rec["cublas_fp16_bias_tflops"] = round(t, 2), tf(t)
                for bits in (4, 3, 2, 8):
                    rng = np.random.default_rng(bits)
                    qwb = torch.from_numpy(rng.integers(-2**31, 2**31, size=(K * bits // 32, N), dtype=np.int64)
                                           .astype(np.int32)).to(dev)
                    qzb = torch.from_numpy(rng.integers(-2**31, 2**31, size=(K // 128, N * bits // 32), dtype=np.int64)
                                           .astype(np.int32)).to(dev)
                    t = time_cold(lambda: ops.qlinear(x, qwb, qzb, sc, bits, 128, bias), flush)
                    rec[f"ours_int{bits}_bias_us"], rec[f"ours_int{bits}_bias_tflops"] = round(t, 2), tf(t)
                    rec[f"ours_int{bits}_frac_burst"] = round(flops / t / 1e6 / peaks["bf16_tflops"], 3)
                    if bits == 4:
                        t = time_loop(lambda: ops.qlinear(x, qwb, qzb, sc, bits, 128, bias))
                        rec["ours_int4_bias_loop_us"] = round(t, 2)
                    # both product paths forced, to document the dispatch rule (default: fused below
                    # M = 2048, unpack-once + pair GEMM from there on)
                    for variant in ("fused", "dense"):
                        os.environ["SAMQ_GEMM"] = variant
                        _lib.reload_config()
                        t = time_cold(lambda: ops.qlinear(x, qwb, qzb, sc, bits, 128, bias), flush)
                        rec[f"ours_int{bits}_{variant}_us"] = round(t, 2)
                        rec[f"ours_int{bits}_{variant}_tflops"] = tf(t)
                    os.environ.pop("SAMQ_GEMM", None)
                    _lib.reload_config()
                # same result? (ours vs the reference kernel, identical packed int4 buffers)
                y_ref = rql.triton_matmul4(128, x, qw, sc, qz, bias).float()
                y = ops.qlinear(x, qw, qz, sc, 4, 128, bias).float()
                rec["ours_vs_triton_maxabs"] = float((y - y_ref).abs().max())
                rec["ours_vs_triton_max_ref"] = float(y_ref.abs().max())
            rec["speedup_vs_ref_triton"] = round(rec["ref_triton_int4_bias_us"] / rec["ours_int4_bias_us"], 2)
            rows.append(rec)
            print("[microbench]", rec, flush=True)
    return {"peaks": peaks, "timing": "median of 20 launches, CUDA events per launch, 256 MB buffer rewritten before "
            "each launch (L2 flushed); *_loop_us = 20 back-to-back launches between one event pair",
            "rows": rows}


# ----------------------------------------------------------------------------------------------
@section("attention")
def run_attention(dev):
    import gptq_triton.fused_attention as rfa
    from sam_quantization_b200 import _lib, ops

    out = {"test_op": [], "compare": []}
    for dtype in (torch.bfloat16, torch.float16):
        for shape in ((25, 14, 14, 3840), (1, 64, 64, 3840)):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                rfa.test_op(*shape, head_num=16, dtype=dtype)
            out["test_op"].append({"shape": shape, "dtype": str(dtype), "printed": buf.getvalue().strip()})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cfgs = [("vit_h win B=1", 25, 14, 16, 80), ("vit_h glob B=1", 1, 64, 16, 80),
            ("vit_h win B=8", 200, 14, 16, 80), ("vit_h glob B=8", 8, 64, 16, 80),
            ("vit_l win B=8", 200, 14, 16, 64), ("vit_l glob B=8", 8, 64, 16, 64),
            ("vit_h win B=32", 800, 14, 16, 80), ("vit_h glob B=32", 32, 64, 16, 80)]
    g = torch.Generator(device=dev).manual_seed(5)
    for tag, B, S, heads, hd in cfgs:
        qkv = (torch.randn(B, S, S, 3 * heads * hd, device=dev, generator=g) * 0.5).half()
        rph = (torch.randn(2 * S - 1, hd, device=dev, generator=g) * 0.2).half()
        rpw = (torch.randn(2 * S - 1, hd, device=dev, generator=g) * 0.2).half()
        scale = hd ** -0.5

        def ref_path():   # fused_attention.py:118-133 (everything between the two GEMMs)
            q = qkv.reshape(B, S * S, 3, heads, -1).permute(2, 0, 3, 1, 4).reshape(3, B * heads, S, S, -1)[0]
            rel_h, rel_w = rfa.add_decomposed_rel_pos(q, rph, rpw, (S, S), (S, S))
            return rfa.forward(qkv, rel_h, rel_w, heads, hd, sm_scale=scale)

        def ours():
            return ops.attn_relpos(qkv, rph, rpw, B, S, S, heads, scale, _lib.RELW_REFERENCE)

        rec = {"case": tag, "B": B, "S": S * S, "heads": heads, "hd": hd}
        try:
            with torch.no_grad():
                y_ref = ref_path().float()
                y = ours().float()
                rec["maxabs_ours_vs_ref_kernel"] = float((y - y_ref).abs().max())
                rec["cos_ours_vs_ref_kernel"] = float(torch.nn.functional.cosine_similarity(
                    y.flatten().double(), y_ref.flatten().double(), dim=0))
                rec["max_ref"] = float(y_ref.abs().max())
                rec["ref_us"] = round(time_cold(ref_path, flush, iters=10, warm=3), 1)
                rec["ours_us"] = round(time_cold(ours, flush, iters=10, warm=3), 1)
                rec["speedup"] = round(rec["ref_us"] / rec["ours_us"], 2)
                rec["ours_tflops"] = round(4.0 * (S * S) ** 2 * hd * heads * B / rec["ours_us"] / 1e6, 1)
        except Exception:
            rec["error"] = traceback.format_exc()
        out["compare"].append(rec)
        print("[attention]", rec, flush=True)
        del qkv
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
def bench_protocol(model, inp, num_iters=100, warmup_iters=25):
    """gptq4sam_infer.py:59-79: warm-up, then wall clock between two synchronizes."""
    with torch.no_grad():
        for _ in range(warmup_iters):
            model(inp)
        torch.cuda.synchronize()
        tik = time.time()
        for _ in range(num_iters):
            model(inp)
        torch.cuda.synchronize()
        tok = time.time()
    return (tok - tik) / num_iters


@section("encoder")
def run_encoder(dev, model_name="vit_h"):
    import gptq_triton
    import gptq_triton.quant_linear as rql
    from segment_anything.modeling.image_encoder import ImageEncoderViT as RefEncoder
    from functools import partial

    from sam_quantization_b200.image_encoder import ENCODER_CONFIGS
    from sam_quantization_b200.launcher import GraphedEncoder
    from sam_quantization_b200.synthetic import random_quantized_encoder

    rql.workspace = torch.empty(20971520, dtype=torch.float16, device=dev)   # the reference's own size
    cfg = ENCODER_CONFIGS[model_name]
    out = {"model": model_name, "protocol": "batch 1, (1,3,1024,1024) fp16, 25 warm-up + 100 iterations, wall clock"}

    ours = random_quantized_encoder(model_name, 4, 128, seed=0, device=dev)
    # random-init weights let the residual stream grow to a few hundred after 32 blocks; the
    # reference's eager fp16 LayerNorm2d squares (x - mean) in fp16 (common.py:38-43) and would
    # overflow to inf -> an all-zero embedding.  Shrinking the (linear, bias-free) neck conv1x1 keeps
    # its input inside fp16's square range without changing what LayerNorm2d then normalises to.
    with torch.no_grad():
        ours.neck[0].weight.mul_(1.0 / 32)
    state = {}
    for k, v in ours.state_dict().items():
        state[k.replace(".attn.qkv_proj.", ".attn.qkv.").replace(".attn.o_proj.", ".attn.proj.")] = v.detach().clone()

    def build_ref():     # build_sam.py:55-80
        torch.manual_seed(0)
        return RefEncoder(depth=cfg["depth"], embed_dim=cfg["embed_dim"], img_size=1024, mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=cfg["num_heads"],
                          patch_size=16, qkv_bias=True, use_rel_pos=True,
                          global_attn_indexes=list(cfg["global_attn_indexes"]), window_size=14, out_chans=256)

    inp = torch.randn(1, 3, 1024, 1024, generator=torch.Generator().manual_seed(3)).half().to(dev)

    probe_blocks = sorted({0, cfg["global_attn_indexes"][0], cfg["depth"] - 1})

    def run_with_hooks(model):
        """(final output, {block index: its output tokens}) of one forward pass."""
        grabbed, handles = {}, []
        for i in probe_blocks:
            handles.append(model.blocks[i].register_forward_hook(
                lambda _m, _inp, o, i=i: grabbed.__setitem__(i, o.detach().float().clone())))
        with torch.no_grad():
            y = model(inp).float().clone()
        for h in handles:
            h.remove()
        return y, grabbed

    def compare(a, b):
        return {"maxabs": float((a - b).abs().max()), "max_ref": float(b.abs().max()),
                "cosine": float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0)),
                "finite": bool(torch.isfinite(a).all() and torch.isfinite(b).all())}

    y_ref = ref_tokens = None
    if model_name == "vit_h":   # the reference's partition is hard-coded to ViT-H, batch 1 (SURVEY trap 3)
        ref = build_ref().half().to(dev).eval()
        t = bench_protocol(ref, inp)
        out["ref_eager_fp16_s_per_iter"] = t
        del ref
        torch.cuda.empty_cache()

        refq = build_ref().half()
        gptq_triton.make_quant(refq, 4, 128)
        missing = refq.load_state_dict(state, strict=False)
        out["ref_load_missing"] = list(missing.missing_keys)[:5]
        out["ref_load_unexpected"] = list(missing.unexpected_keys)[:5]
        gptq_triton.make_quant_attn(refq)
        refq = refq.to(dev).eval()
        y_ref, ref_tokens = run_with_hooks(refq)
        t = bench_protocol(refq, inp)
        out["ref_quant_triton_s_per_iter"] = t
        out["ref_quant_triton_images_per_s"] = 1.0 / t
        del refq
        torch.cuda.empty_cache()

    y, our_tokens = run_with_hooks(ours)
    if y_ref is not None:
        # identical packed weights + identical input through the reference's own GPU path
        # (Triton QuantLinear + Triton attention + eager fp16 LayerNorm / GELU / residual) and ours
        out["parity_ours_vs_reference_gpu"] = {
            "block_outputs": {str(i): compare(our_tokens[i], ref_tokens[i]) for i in probe_blocks},
            "embedding": compare(y, y_ref),
            "note": "identical packed weights and input through the reference's own GPU path (Triton QuantLinear + "
                    "Triton attention + eager fp16 LayerNorm / GELU / residual) and ours; neck conv1x1 weight "
                    "scaled by 1/32 in both so the reference's fp16 LayerNorm2d does not overflow"}
    t = bench_protocol(ours, inp)
    out["ours_eager_launch_s_per_iter"] = t
    genc = GraphedEncoder(ours, inp)
    t = bench_protocol(genc, inp)
    out["ours_graph_s_per_iter"] = t
    out["ours_graph_images_per_s"] = 1.0 / t
    if "ref_quant_triton_s_per_iter" in out:
        out["speedup_vs_reference_gpu"] = out["ref_quant_triton_s_per_iter"] / t
    return out


@section("decoder")
def run_decoder(dev):
    """Row f-3 on the GPU: the reference's PromptEncoder + MaskDecoder (fp16, eager torch / cuBLAS) against
    this package's modules (decoder on libsamq kernels) on identical weights and prompts: parity and time
    per call for the click loop's shapes (one image, 1..5 clicks; and a 64-prompt batch)."""
    from segment_anything.modeling.mask_decoder import MaskDecoder as RefDecoder
    from segment_anything.modeling.prompt_encoder import PromptEncoder as RefPrompt
    from segment_anything.modeling.transformer import TwoWayTransformer as RefTransformer

    from sam_quantization_b200.mask_decoder import MaskDecoder, TwoWayTransformer
    from sam_quantization_b200.prompt_encoder import PromptEncoder

    torch.manual_seed(11)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    md = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                     transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    rpe = RefPrompt(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    rmd = RefDecoder(num_multimask_outputs=3, transformer=RefTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                     transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    rpe.load_state_dict(pe.state_dict(), strict=True)
    rmd.load_state_dict(md.state_dict(), strict=True)
    pe, md, rpe, rmd = (m.half().to(dev).eval() for m in (pe, md, rpe, rmd))
    g = torch.Generator().manual_seed(5)
    emb = (torch.randn(1, 256, 64, 64, generator=g) * 0.5).half().to(dev)
    out = {"cases": []}
    for nprompt, nclick in ((1, 1), (1, 5), (64, 1)):
        pts = (torch.rand(nprompt, nclick, 2, generator=g) * 1024).half().to(dev)
        labs = torch.randint(0, 2, (nprompt, nclick), generator=g).half().to(dev)

        def run(p, m):
            sparse, dense = p(points=(pts, labs), boxes=None, masks=None)
            return m(image_embeddings=emb, image_pe=p.get_dense_pe().half(), sparse_prompt_embeddings=sparse,
                     dense_prompt_embeddings=dense, multimask_output=False)

        with torch.no_grad():
            ref_masks, ref_iou = run(rpe, rmd)
            our_masks, our_iou = run(pe, md)
            t_ref = time_loop(lambda: run(rpe, rmd), iters=20, warm=3)
            t_our = time_loop(lambda: run(pe, md), iters=20, warm=3)
        out["cases"].append({
            "prompts": nprompt, "clicks": nclick,
            "masks_maxabs": float((our_masks.float() - ref_masks.float()).abs().max()),
            "masks_max_ref": float(ref_masks.float().abs().max()),
            "masks_cosine": float(torch.nn.functional.cosine_similarity(our_masks.flatten().double(), ref_masks.flatten().double(), dim=0)),
            "iou_maxabs": float((our_iou.float() - ref_iou.float()).abs().max()),
            "reference_fp16_eager_us": round(t_ref, 1), "ours_us": round(t_our, 1)})
        print("[decoder]", out["cases"][-1], flush=True)
    out["note"] = ("both in fp16 on the same GPU, eager launches (no CUDA graph); the decoder is < 1 % of the encoder's "
                   "FLOPs -- the figure of merit is parity, the time is for the record")
    return out


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu"))
    ap.add_argument("--sections", default="dequant,microbench,attention,encoder,decoder")
    args = ap.parse_args()
    OUT = args.out
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if os.path.exists(OUT + ".json"):          # sections accumulate over calls
        with open(OUT + ".json") as f:
            RESULT.update(json.load(f))
    make_ref.add_to_path()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    import triton

    peaks = {"bf16_tflops": 1626.1, "bf16_tflops_sustained": 1367.0}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peaks = json.load(f)
    RESULT["env"] = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "triton": triton.__version__,
                     "host_cores": os.cpu_count()}
    t0 = time.time()
    import gptq_triton  # noqa: F401  (module-level CUDA workspace: quant_linear.py:13)

    RESULT["env"]["import_gptq_triton_s"] = round(time.time() - t0, 1)
    triton_api_shim()
    RESULT["env"]["shim"] = "matmul4_kernel.early_config_prune wrapped to drop Triton>=3 launch kwargs (grid, warmup)"
    save()
    secs = args.sections.split(",")
    if "dequant" in secs:
        run_dequant(dev)
    if "microbench" in secs:
        run_microbench(dev, peaks)
    if "attention" in secs:
        run_attention(dev)
    if "encoder" in secs:
        run_encoder(dev)
    if "decoder" in secs:
        run_decoder(dev)
    print(json.dumps(RESULT.get("_seconds")))


if __name__ == "__main__":
    main()
