"""Tiny driver for profiling one dequant-GEMM shape: python tests/gemm_bench.py M K N [gelu] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_quantization_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import rand_packed, dev

M, K, N = (int(v) for v in sys.argv[1:4])
epi = _lib.EPI_GELU if len(sys.argv) > 4 and sys.argv[4] == "gelu" else _lib.EPI_NONE
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
d = torch.device("cuda:0")
qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=1)
tq, tz, ts = dev(qw, d), dev(qz, d), dev(sc, d)
x = torch.randn(M, K, device=d).half()
bias = torch.randn(N, device=d).half()
y = torch.empty(M, N, device=d, dtype=torch.float16)
for _ in range(3):
    ops.qlinear(x, tq, tz, ts, 4, 128, bias, epilogue=epi, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.qlinear(x, tq, tz, ts, 4, 128, bias, epilogue=epi, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"M={M} K={K} N={N} epi={epi}: {ms*1e3:.1f} us {2*M*K*N/ms/1e9:.1f} TFLOP/s")
