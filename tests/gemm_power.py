"""Sustained (power-capped) throughput of the GEMM variants with clocks/power sampling.
usage: python tests/gemm_power.py M K N seconds"""
import os, subprocess, sys, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sam_quantization_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import rand_packed, dev

M, K, N = (int(v) for v in sys.argv[1:4])
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
d = torch.device("cuda:0")
qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=1)
tq, tz, ts = dev(qw, d), dev(qz, d), dev(sc, d)
x = torch.randn(M, K, device=d).half()
y = torch.empty(M, N, device=d, dtype=torch.float16)
wt = ops.unpack_dequant(tq, tz, ts, 4, 128, transposed=True)
w = wt.t().contiguous()

def sample(tag, fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    iters = max(10, int(secs * 1e3 / e0.elapsed_time(e1)))
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
    lines = []
    th = threading.Thread(target=lambda: [lines.append(l) for l in p.stdout], daemon=True); th.start()
    time.sleep(0.15)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    p.terminate()
    ms = e0.elapsed_time(e1) / iters
    clk = [float(l.split(",")[0]) for l in lines[2:] if l.strip()]
    pw = [float(l.split(",")[1]) for l in lines[2:] if l.strip()]
    cap = sum("Active" in l and "Not" not in l for l in lines[2:])
    print(f"{tag:14s} {ms*1e3:8.1f} us {2*M*K*N/ms/1e9:7.1f} TFLOP/s  iters {iters} sm_mhz med {statistics.median(clk) if clk else 0:.0f} power med {statistics.median(pw) if pw else 0:.0f} W cap_samples {cap}/{len(lines)-2}", flush=True)

sample("fused", lambda: ops.qlinear(x, tq, tz, ts, 4, 128, out=y))
sample("dense(ss)", lambda: ops.dense_linear(x, wt))
sample("cublas fp16", lambda: torch.matmul(x, w, out=y))
