"""Stress check of the weight prefetch / co-running pad fill (programmatic launches that write scratch
and output rows next to a running GEMM): many eager passes and many CUDA-graph replays of a ViT-H-width
encoder at batch 32 must reproduce the no-prefetch result bit for bit.
usage: python tests/runs/stress_prefetch.py [depth] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from sam_quantization_b200 import _lib
from sam_quantization_b200.launcher import GraphedEncoder
from sam_quantization_b200.synthetic import random_quantized_encoder

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
enc = random_quantized_encoder("vit_h", 4, 128, seed=1, device=dev, depth=depth, global_attn_indexes=(depth - 1,))
g = torch.Generator(device=dev).manual_seed(2)
xs = [torch.randn(32, 3, 1024, 1024, device=dev, generator=g).half() for _ in range(2)]
with torch.no_grad():
    os.environ["SAMQ_PREFETCH"] = "0"
    _lib.reload_config()
    ref = [enc(x).clone() for x in xs]
    os.environ["SAMQ_PREFETCH"] = "1"
    _lib.reload_config()
    bad = 0
    for i in range(iters):
        y = enc(xs[i & 1])
        bad += int(not torch.equal(y, ref[i & 1]))
    print(f"eager: {iters} passes, {bad} mismatches")
    ge = GraphedEncoder(enc, xs[0])
    badg = 0
    for i in range(iters):
        y = ge(xs[i & 1])
        badg += int(not torch.equal(y, ref[i & 1]))
    print(f"graph: {iters} replays, {badg} mismatches; kernels per replay {ge.kernels_per_replay}")
sys.exit(1 if bad or badg else 0)
