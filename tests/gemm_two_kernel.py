"""fused int4 GEMM vs (unpack_dequant -> dense tcgen05 GEMM) per call, sustained."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from sam_quantization_b200 import ops, _lib
from gpu_util import rand_packed, dev
d = torch.device("cuda:0")
def bench(fn, secs=1.0):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    iters = max(10, int(secs * 1e3 / e0.elapsed_time(e1)))
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (M, K, N, epi) in [(39200, 1280, 3840, 0), (39200, 1280, 1280, 0), (32768, 1280, 5120, 1), (32768, 5120, 1280, 0), (4096, 1280, 5120, 1), (4900, 1280, 3840, 0), (4096, 5120, 1280, 0)]:
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=1)
    tq, tz, ts = dev(qw, d), dev(qz, d), dev(sc, d)
    x = torch.randn(M, K, device=d).half(); b = torch.randn(N, device=d).half()
    y = torch.empty(M, N, device=d, dtype=torch.float16)
    t_f = bench(lambda: ops.qlinear(x, tq, tz, ts, 4, 128, b, epilogue=epi, out=y))
    t_d = bench(lambda: ops.unpack_dequant(tq, tz, ts, 4, 128, transposed=True))
    def two():
        wt = ops.unpack_dequant(tq, tz, ts, 4, 128, transposed=True)
        return ops.dense_linear(x, wt, b, epilogue=epi)
    t_2 = bench(two)
    fl = 2 * M * K * N
    print(f"M={M} K={K} N={N} epi={epi}: fused {t_f:7.1f} us {fl/t_f/1e6:7.1f} TF | dequant {t_d:6.1f} us | dequant+dense {t_2:7.1f} us {fl/t_2/1e6:7.1f} TF", flush=True)
