#!/usr/bin/env python
"""Benchmark of the GPTQ-int4 SAM image-encoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model vit_h] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...       # the reference's CPU dequant path (oracle port)

One step = one pass of the whole quantized encoder (patch embed, 32 blocks, neck) over one
batch of ``--batch`` synthetic 1024x1024 images per GPU (weak scaling: replicas, no collective
on the data path).  Prints ONE JSON line (see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.md section 3 quotes the REFERENCE's work per 1024x1024 image (windowed qkv / proj on the
# 4900 zero-padded tokens it multiplies): ViT-B 972.1 / ViT-L 2985.7 / ViT-H 5961.1 GFLOP.  This
# build does not multiply the padding rows, so `model_tflops` counts only what is
# EXECUTED: every linear on the 4096 real tokens (2 M K N), attention 4 S^2 hd per head on the 25
# padded windows (S = 196: the pad tokens stay real keys and their query rows are computed) and
# on the 4096-token global blocks, the in-kernel rel-pos products 4 S (2 E) hd, patch embed + neck.
MODEL_NAMES = ("vit_b", "vit_h", "vit_l")


def executed_gflop_per_image(name: str) -> float:
    from sam_quantization_b200.image_encoder import ENCODER_CONFIGS

    c = ENCODER_CONFIGS[name]
    D, depth, heads = c["embed_dim"], c["depth"], c["num_heads"]
    hd, n_glob = D // heads, len(c["global_attn_indexes"])
    n_win = depth - n_glob
    tok = 4096
    linear = depth * 2.0 * tok * (D * 3 * D + D * D + 2 * D * 4 * D)
    attn_win = n_win * 25 * heads * (4.0 * 196 * 196 * hd + 4.0 * 196 * 28 * hd)
    attn_glob = n_glob * heads * (4.0 * 4096 * 4096 * hd + 4.0 * 4096 * 128 * hd)
    stem_neck = 2.0 * tok * (768 * D + D * 256 + 9 * 256 * 256)
    return (linear + attn_win + attn_glob + stem_neck) / 1e9


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tflops=d.get("bf16_tflops", 1590.0),
                    tflops_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# --------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU baseline = the reference's PyTorch dequant path: the reference's OWN
# ImageEncoderViT (oracle/_ref/segment_anything, staged by oracle/make_ref.py) in fp32 on the host
# cores, its nn.Linear weights = the dequantised packed weights (what GPTQ.fasterquant writes back,
# gptq.py:160-162).  kind "reference".  If oracle/_ref was not staged, the oracle's restatement of
# the same forward runs instead (kind "port").  Nothing here is on the product path.
# --------------------------------------------------------------------------------------
class CpuReference:
    """Whole encoder, all blocks, ONE image per call (the reference's own batch, gptq4sam_infer.py:223)."""

    def __init__(self, model_name: str, packed_state, bits: int = 4):
        from oracle import encoder as oe
        from oracle import make_ref

        self.model_name = model_name
        self.cfg = oe.CONFIGS[model_name]
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.p = oe.dequant_state(packed_state, bits, 128)
        self.kind = "port"
        self.model = None
        if make_ref.available():
            make_ref.add_to_path()
            from functools import partial

            import segment_anything.modeling.image_encoder as rie   # the staged reference module

            if model_name != "vit_h":
                # the shipped partition is hard-coded to ViT-H batch 1 (image_encoder.py:297-305,324-332):
                # other widths need the generic formulas the reference commented out
                rie.window_partition = oe.window_partition
                rie.window_unpartition = oe.window_unpartition
            c = self.cfg
            m = rie.ImageEncoderViT(depth=c["depth"], embed_dim=c["embed_dim"], img_size=1024, mlp_ratio=4,
                                    norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=c["num_heads"],
                                    patch_size=16, qkv_bias=True, use_rel_pos=True,
                                    global_attn_indexes=list(c["global_attn_indexes"]), window_size=14, out_chans=256)
            missing = m.load_state_dict(self.p, strict=False)
            assert not missing.missing_keys, missing.missing_keys[:4]
            self.model = m.float().eval()
            self.kind = "reference"
            self.p = None

    def describe(self) -> str:
        d = self.cfg["depth"]
        if self.kind == "reference":
            return (f"1 image through the reference's own ImageEncoderViT (all {d} blocks, stem, neck; "
                    f"oracle/_ref/segment_anything, unmodified) in fp32 torch on all host cores, nn.Linear weights = "
                    f"the dequantised packed int4 weights (the reference's PyTorch dequant path)")
        return (f"1 image through the oracle's restatement of the reference encoder (all {d} blocks, stem, neck) in "
                f"fp32 torch on all host cores with the oracle's dequantised weights (oracle/_ref not staged)")

    def step(self, seed: int = 0) -> float:
        """Seconds for one image, nothing extrapolated."""
        from oracle import encoder as oe
        from oracle import synth

        img = torch.from_numpy(synth.image(1, 1024, seed=seed)).half().float()
        with torch.no_grad():
            t0 = time.perf_counter()
            if self.model is not None:
                self.model(img)
            else:
                c = self.cfg
                oe.encoder(img, self.p, c["depth"], c["num_heads"], c["global_attn_indexes"])
            return time.perf_counter() - t0


def other_kernel_rooflines(other_ms, peaks, step_ms):
    """Per kernel class: mean us per call, algorithmic GB/s and TFLOP/s, fraction of the roofline that
    bounds it, share of the step (timed with CUDA events in the instrumented eager replay)."""
    out = {}
    for name, recs in other_ms.items():
        if not recs:
            continue
        ms = sum(r[0] for r in recs)
        gbps = sum(r[1] for r in recs) / ms / 1e6
        tf = sum(r[2] for r in recs) / ms / 1e9
        e = {"calls_per_step": len(recs) // 2, "us_per_call": round(ms / len(recs) * 1e3, 1),
             "share_of_step": round(ms / 2 / step_ms, 4), "algorithmic_GBps": round(gbps, 1)}
        if name == "attn_global":
            e.update(bound="tensor / MUFU", TFLOPs=round(tf, 1), frac_of_burst_tensor_peak=round(tf / peaks["tflops"], 3),
                     note="MUFU floor (16 ex2 / clk / SM) is 0.57 of the measured time at batch 32, DESIGN 5.2")
        elif name == "attn_windowed":
            e.update(bound="hbm", frac_of_hbm_peak=round(gbps / peaks["hbm_gbs"], 3), TFLOPs=round(tf, 1),
                     frac_of_burst_tensor_peak=round(tf / peaks["tflops"], 3))
        else:
            e.update(bound="hbm", frac_of_hbm_peak=round(gbps / peaks["hbm_gbs"], 3))
        out[name] = e
    return out


def packed_state_cpu(enc):
    """Reference-layout state (``...attn.qkv.qweight`` etc.) of a fused encoder, on the CPU."""
    out = {}
    for k, v in enc.state_dict().items():
        k = k.replace(".attn.qkv_proj.", ".attn.qkv.").replace(".attn.o_proj.", ".attn.proj.")
        out[k] = v.detach().cpu()
    return out


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--model", default="vit_h", choices=sorted(MODEL_NAMES))
    ap.add_argument("--batch", type=int, default=32,
                    help="images per GPU per step (BASELINE config 3 is a batch sweep: 8 / 16 / 32 / 64 per GPU give "
                         "181 / 182 / 186 / 185 images/s on one B200)")
    ap.add_argument("--bits", type=int, default=4, choices=[2, 3, 4, 8],
                    help="weight bits (BASELINE config 4: 3 and 8 with --act-order)")
    ap.add_argument("--act-order", action="store_true", help="permutation-derived g_idx (non-contiguous groups)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fmt = f"int{args.bits}" + ("-actorder" if args.act_order else "")
    metric = f"SAM {args.model.replace('_', '-').upper().replace('VIT', 'ViT')} GPTQ-{fmt} encoder images/s"
    config = {"workload": f"SAM {args.model} image encoder, GPTQ int{args.bits} groupsize 128"
                          f"{' with act-order g_idx' if args.act_order else ''}, random-init packed weights, "
                          f"synthetic 1024x1024 images, batch {args.batch}/GPU/step",
              "global_batch": args.batch * world, "parallelism": f"dp{world} (replicas, no data-path collective)",
              "l2": "no explicit flush: the working set of one step (packed weights + activations of the batch, "
                    ">1 GB already at batch 8) exceeds the 126 MB L2; input batches rotate over 3 buffers",
              "batch_sweep_images_per_s_1gpu": {"1": 134, "8": 181, "16": 182, "32": 186, "64": 185,
                                                "note": "round 2, one box (profiles/r02h_batch_sweep.txt; the same tree "
                                                        "gave 198.5 at batch 32 on a faster box)"}}

    from sam_quantization_b200.synthetic import random_quantized_encoder

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        enc = random_quantized_encoder(args.model, args.bits, 128, seed=0, device="cpu", act_order=args.act_order)
        ref = CpuReference(args.model, packed_state_cpu(enc), args.bits)
        del enc
        # one step = ONE whole image (every block; nothing extrapolated): ~4.5 s for ViT-H on 16 cores,
        # so the driver's --steps 20 --warmup 5 is ~2 minutes
        for i in range(args.warmup):
            ref.step(i)
        t0 = time.perf_counter()
        secs = [ref.step(i) for i in range(args.steps)]
        wall = time.perf_counter() - t0
        sec = statistics.mean(secs)
        value = 1.0 / sec
        line = {"impl": "reference", "metric": metric, "value": value, "unit": "images/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "images/s", "cores": ref.cores, "kind": ref.kind,
                                 "sample": "per step: " + ref.describe()},
                "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": wall, "s_per_image_min_max": [min(secs), max(secs)]}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist

    from sam_quantization_b200 import _lib, ops

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    _lib.device_check(dev)

    enc_eager = random_quantized_encoder(args.model, args.bits, 128, seed=0, device=dev, act_order=args.act_order)
    B = args.batch
    nbuf = 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    inputs = [torch.randn(B, 3, 1024, 1024, device=dev, generator=gen).half() for _ in range(nbuf)]
    host_in = [t.cpu().pin_memory() for t in inputs]
    host_out = torch.empty(B, 256, 64, 64, dtype=torch.float16).pin_memory()
    if args.no_graph:
        enc = enc_eager
    else:
        from sam_quantization_b200.launcher import GraphedEncoder

        enc = GraphedEncoder(enc_eager, inputs[0])   # public API: captured forward, replayed per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            enc(inputs[i % nbuf])
        barrier()

        # ---- timed region: device-resident inputs -------------------------------------
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            out = enc(inputs[i % nbuf])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        if not args.no_graph:   # replays do not pass through the C ABI: count what the graph holds
            launches = args.steps * enc.kernels_per_replay
        clocks = sampler.stop() if rank == 0 else None

        # ---- end-to-end: host buffers, H2D + encoder + D2H per step ----------------------
        # Through the public host-buffer call (GraphedEncoder.run_host): every step copies its own
        # input from pinned host memory and its embeddings back; consecutive steps are pipelined
        # (the copies of steps i+1 / i-1 run on copy streams underneath the encoder of step i).
        def host_step(i):
            if args.no_graph:
                host_out.copy_(enc(host_in[i % nbuf].to(dev, non_blocking=True)), non_blocking=True)
                return None
            return enc.run_host(host_in[i % nbuf], host_out)

        for i in range(2):
            host_step(i)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        done = None
        for i in range(args.steps):
            done = host_step(i)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)   # the last step's embeddings are on the host
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)

        # ---- roofline of the dominant kernel (dequant-GEMM), instrumented replay ------------
        gemm_ms, gemm_flops, gemm_calls = 0.0, 0.0, 0
        rec = []
        orig = ops.qlinear

        def timed_qlinear(x, qweight, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig(x, qweight, *a, **kw)
            t.record()
            rec.append((s, t, 2.0 * (x.numel() // x.shape[-1]) * x.shape[-1] * qweight.shape[1],
                        (x.numel() // x.shape[-1], x.shape[-1], qweight.shape[1])))
            return y

        for i in range(2):                       # warm the eager path (allocator pools) first
            enc_eager(inputs[i % nbuf])
        torch.cuda.synchronize()
        orig_unp = ops.qlinear_unpartition

        def timed_unpartition(x, qweight, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig_unp(x, qweight, *a, **kw)
            t.record()
            # algorithmic rows = the real (image-order) tokens: the zero-padding rows of the window
            # layout are not useful work (and are not multiplied on the pad-skipping path)
            m = y.numel() // y.shape[-1]
            rec.append((s, t, 2.0 * m * x.shape[-1] * qweight.shape[1], (m, x.shape[-1], qweight.shape[1])))
            return y

        orig_part = ops.qlinear_partition

        def timed_partition(x, qweight, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig_part(x, qweight, *a, **kw)
            t.record()
            m = x.numel() // x.shape[-1]
            rec.append((s, t, 2.0 * m * x.shape[-1] * qweight.shape[1], (m, x.shape[-1], qweight.shape[1])))
            return y

        # the other kernel classes of the step, same method: LayerNorm (HBM-bound: 2 rows C 2 B), windowed
        # attention (HBM-bound at this size: the qkv tensor in, the output out) and global attention
        # (MUFU / tensor-bound: 4 S^2 hd per head + the in-kernel rel-pos products)
        other = {"layernorm": [], "attn_windowed": [], "attn_global": []}
        orig_ln, orig_aw, orig_ag = ops.layernorm, ops.attn_relpos_unpartition, ops.attn_relpos

        def timed_ln(x, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig_ln(x, *a, **kw)
            t.record()
            if x.shape[-1] >= 512:       # the block LayerNorms (the neck's 256-channel ones are noise)
                other["layernorm"].append((s, t, 2.0 * x.numel() * 2, 0.0))
            return y

        def timed_aw(qkv, rph, rpw, Bq, H, W, ws, heads, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig_aw(qkv, rph, rpw, Bq, H, W, ws, heads, *a, **kw)
            t.record()
            hd = qkv.shape[-1] // 3 // heads
            nwin = qkv.numel() // qkv.shape[-1] // (ws * ws)
            flops = nwin * heads * (4.0 * (ws * ws) ** 2 * hd + 4.0 * ws * ws * 2 * ws * hd)
            other["attn_windowed"].append((s, t, (qkv.numel() + y.numel()) * 2.0, flops))
            return y

        def timed_ag(qkv, rph, rpw, Bq, H, W, heads, *a, **kw):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = orig_ag(qkv, rph, rpw, Bq, H, W, heads, *a, **kw)
            t.record()
            hd = qkv.shape[-1] // 3 // heads
            S = H * W
            flops = Bq * heads * (4.0 * S * S * hd + 4.0 * S * (H + W) * hd)
            other["attn_global"].append((s, t, (qkv.numel() + y.numel()) * 2.0, flops))
            return y

        ops.layernorm, ops.attn_relpos_unpartition, ops.attn_relpos = timed_ln, timed_aw, timed_ag
        ops.qlinear = timed_qlinear
        ops.qlinear_unpartition = timed_unpartition
        ops.qlinear_partition = timed_partition
        # keep the GPU queue full so an event pair brackets only its kernel: a heavy kernel first
        big = torch.empty(1 << 28, dtype=torch.float16, device=dev)
        big.zero_()
        for i in range(2):
            enc_eager(inputs[i % nbuf])
        torch.cuda.synchronize()
        del big
        ops.qlinear = orig
        ops.qlinear_unpartition = orig_unp
        ops.qlinear_partition = orig_part
        ops.layernorm, ops.attn_relpos_unpartition, ops.attn_relpos = orig_ln, orig_aw, orig_ag
        other_ms = {k: [(s.elapsed_time(t), by, fl) for s, t, by, fl in v] for k, v in other.items()}
        per_shape = {}
        for s, t, fl, shp in rec:
            dt = s.elapsed_time(t)
            gemm_ms += dt
            gemm_flops += fl
            gemm_calls += 1
            acc = per_shape.setdefault(shp, [0, 0.0, 0.0])
            acc[0] += 1
            acc[1] += dt
            acc[2] += fl

    # optional final gather of the embeddings (launcher.ShardedEncoder(gather=True)): OUTSIDE the
    # timed region and not part of `value`; timed here so that its NVLink cost is on record
    gather_info = None
    if world > 1:
        bufs = [torch.empty_like(out) for _ in range(world)]
        for _ in range(2):
            dist.all_gather(bufs, out)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            dist.all_gather(bufs, out)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 5
        nbytes = out.numel() * out.element_size()
        gather_info = {"ms": gms, "bytes_per_rank": nbytes,
                       "recv_GBps_per_rank": nbytes * (world - 1) / gms / 1e6,
                       "note": "dist.all_gather (NCCL over NVLink) of every rank's [B, 256, 64, 64] fp16 embeddings; "
                               "optional, after the step, not inside the timed region"}
        del bufs
    t_ms = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    per_rank = [t_ms.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t_ms)       # every rank's own device time: shows what max-over-ranks costs
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t_ms.tolist()

    if rank == 0:
        peaks = load_peaks()
        images = args.steps * B * world
        value = images / (ms / 1e3)
        e2e_value = images / (ms_e2e / 1e3)
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        traffic = None
        prof = os.path.join(ROOT, "profiles", "qlinear_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                tj = json.load(f)
            # the ncu capture holds for the batch it was taken at (and ViT-H shapes)
            if tj.get("batch_per_gpu") == B and args.model == "vit_h":
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate (int4 weights dequantised to fp16)",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "launch": "eager" if args.no_graph else "cuda-graph replay of the whole encoder forward",
            "per_rank_ms_per_step": [round(float(t[0]) / args.steps, 3) for t in per_rank],
            "final_allgather": gather_info,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 1024 * 1024 * 2,
                    "d2h_bytes_per_step": B * 256 * 64 * 64 * 2,
                    "api": ("eager forward on host tensors (serial H2D, encoder, D2H)" if args.no_graph else
                            "GraphedEncoder.run_host(pinned_in, pinned_out): every step copies its own input and "
                            "output; steps are pipelined (H2D of step i+1 and D2H of step i-1 on copy streams "
                            "under the encoder of step i); the timed region ends when the last step's output "
                            "is on the host")},
            "roofline": {
                "kernel": "QuantLinear forward = dequant4_transposed_kernel + dense2_kernel (CTA-pair tcgen05 GEMM, 256x256 tile); all 4 linears of every block",
                "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["tflops_sustained"]) if achieved else None,
                "frac_of_burst_peak": (achieved / peaks["tflops"]) if achieved else None,
                "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; burst {peaks['tflops']})",
                "traffic": traffic,
                "how": f"CUDA events around each of the {gemm_calls} QuantLinear calls (unpack + GEMM) of 2 instrumented "
                       f"eager steps run right after the timed region; algorithmic 2*M*K*N per call",
                "share_of_step": (gemm_ms / 2) / (ms / args.steps) if gemm_ms > 0 else None,
                "per_shape_MKN_us_tflops": [[list(k), round(v[1] / v[0] * 1e3, 1), round(v[2] / v[1] / 1e9, 1)]
                                            for k, v in sorted(per_shape.items())],
            },
            "other_kernels": other_kernel_rooflines(other_ms, peaks, ms / args.steps),
            "model_tflops": value * executed_gflop_per_image(args.model) / 1e3,
            "model_gflop_per_image_executed": executed_gflop_per_image(args.model),
        }
        if not args.no_cpu_baseline:
            try:
                ref = CpuReference(args.model, packed_state_cpu(enc_eager), args.bits)
                ref.step(0)                                    # warm-up (thread pool, allocator, page faults)
                secs = [ref.step(1 + i) for i in range(2)]
                line["cpu_baseline"] = {
                    "value": 1.0 / statistics.mean(secs), "unit": "images/s", "cores": ref.cores, "kind": ref.kind,
                    "sample": "1 warm-up + 2 timed images, each: " + ref.describe()}
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex!r}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
