"""Ad-hoc GPU probe (not a pytest): prints detailed diagnostics for one kernel family.
usage: python tests/probe_gpu.py {dequant|ln|dense|fused|attn} """
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sam_quantization_b200 import ops, _lib
from oracle import quant as oq

dev = torch.device("cuda:0")

def rand_packed(K, N, bits, gs, seed=0, g_idx=False):
    rng = np.random.default_rng(seed)
    G = K // gs
    qweight = rng.integers(-2**31, 2**31, size=(K * bits // 32, N), dtype=np.int64).astype(np.int32)
    qzeros = rng.integers(-2**31, 2**31, size=(G, N * bits // 32), dtype=np.int64).astype(np.int32)
    scales = rng.uniform(0.002, 0.02, size=(G, N)).astype(np.float16)
    gi = None
    if g_idx:
        perm = rng.permutation(K)
        inv = np.empty(K, dtype=np.int64); inv[perm] = np.arange(K)
        gi = (inv // gs).astype(np.int32)
    return qweight, qzeros, scales, gi

def t(a):
    return None if a is None else torch.from_numpy(a).to(dev)

def probe_dequant():
    for bits in (4, 8, 2, 3):
        for gidx in (False, True):
            for tr in (False, True):
                K, N, gs = 256, 384, 128
                qw, qz, sc, gi = rand_packed(K, N, bits, gs, seed=bits, g_idx=gidx)
                ref = oq.dequant(qw, qz, sc, bits, gs, gi)
                out = ops.unpack_dequant(t(qw), t(qz), t(sc), bits, gs, t(gi), transposed=tr)
                torch.cuda.synchronize()
                o = out.cpu().numpy()
                if tr: o = o.T
                bad = (o.view(np.uint16) != ref.view(np.uint16)).sum()
                print(f"dequant bits={bits} g_idx={gidx} transposed={tr}: mismatches {bad}/{ref.size}")

def probe_ln():
    torch.manual_seed(0)
    for C in (768, 1024, 1280):
        x = torch.randn(2, 64, 64, C, device=dev).half()
        g = (1 + 0.1 * torch.randn(C, device=dev)).half(); b = (0.1 * torch.randn(C, device=dev)).half()
        y = ops.layernorm(x, g, b, 1e-6)
        ref = torch.nn.functional.layer_norm(x.float(), (C,), g.float(), b.float(), 1e-6)
        print(f"ln C={C}: maxabs {(y.float()-ref).abs().max().item():.3e}")
        yp, hw = ops.layernorm_partition(x, g, b, 1e-6, 14)
        refp = torch.nn.functional.pad(ref, (0, 0, 0, 6, 0, 6)).view(2, 5, 14, 5, 14, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 14, 14, C)
        print(f"ln+partition C={C}: maxabs {(yp.float()-refp).abs().max().item():.3e} hw={hw}")
        sc = torch.randn(2, 64, 64, C, device=dev).half()
        o = ops.unpartition_residual(yp, sc, 14)
        refo = sc.float() + refp.view(2, 5, 5, 14, 14, C).permute(0, 1, 3, 2, 4, 5).reshape(2, 70, 70, C)[:, :64, :64].half().float()
        print(f"unpartition+res C={C}: maxabs {(o.float()-refo).abs().max().item():.3e}")
        a = ops.add(x, sc)
        print(f"add: maxabs {(a.float()-(x.float()+sc.float())).abs().max().item():.3e}")

def gemm_report(name, y, ref):
    yf = y.float(); d = (yf - ref).abs()
    cos = torch.nn.functional.cosine_similarity(yf.flatten(), ref.flatten(), dim=0).item()
    print(f"{name}: maxabs {d.max().item():.4e} ref_absmax {ref.abs().max().item():.3f} cos {cos:.7f} nan {torch.isnan(yf).sum().item()}")
    if not (cos > 0.9999):
        # localise: per 32x32 block error map summary
        M, N = ref.shape
        bad = (d > 0.05 * ref.abs().max()).float()
        print("   bad frac", bad.mean().item(), "bad rows", bad.any(1).sum().item(), "/", M, "bad cols", bad.any(0).sum().item(), "/", N)
        rows = bad.any(1).nonzero().flatten()[:10].tolist(); cols = bad.any(0).nonzero().flatten()[:10].tolist()
        print("   first bad rows", rows, "cols", cols)
        print("   y[0,:8]", yf[0, :8].tolist()); print("   r[0,:8]", ref[0, :8].tolist())

def probe_dense():
    torch.manual_seed(0)
    for (M, K, N) in [(192, 64, 128), (192, 128, 128), (200, 256, 256), (4096, 1280, 1280), (4900, 1280, 3840)]:
        x = torch.randn(M, K, device=dev).half()
        wt = (torch.randn(N, K, device=dev) * 0.05).half()
        bias = torch.randn(N, device=dev).half()
        y = ops.dense_linear(x, wt, bias)
        torch.cuda.synchronize()
        ref = x.float() @ wt.float().t() + bias.float()
        gemm_report(f"dense M={M} K={K} N={N}", y, ref)
    x = torch.randn(4096, 1280, device=dev).half(); wt = (torch.randn(5120, 1280, device=dev) * 0.05).half(); bias = torch.randn(5120, device=dev).half()
    res = torch.randn(4096, 5120, device=dev).half()
    y = ops.dense_linear(x, wt, bias, epilogue=_lib.EPI_GELU, residual=res)
    ref = torch.nn.functional.gelu(x.float() @ wt.float().t() + bias.float()) + res.float()
    gemm_report("dense gelu+res 4096x1280x5120", y, ref)
    for _ in range(3): ops.dense_linear(x, wt, bias)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.dense_linear(x, wt, bias)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"dense 4096x1280x5120: {ms*1e3:.1f} us  {2*4096*1280*5120/ms/1e9:.1f} TFLOP/s")

def probe_fused():
    torch.manual_seed(0)
    for (M, K, N, gs) in [(192, 128, 128, 128), (192, 256, 128, 128), (200, 256, 256, 64), (4096, 1280, 1280, 128), (4900, 1280, 3840, 128), (4096, 5120, 1280, 128), (4096, 1280, 5120, -1)]:
        g = K if gs == -1 else gs
        qw, qz, sc, _ = rand_packed(K, N, 4, g, seed=1)
        x = torch.randn(M, K, device=dev).half()
        bias = torch.randn(N, device=dev).half()
        y = ops.qlinear(x, t(qw), t(qz), t(sc), 4, gs, bias)
        torch.cuda.synchronize()
        w = torch.from_numpy(oq.dequant(qw, qz, sc, 4, g)).to(dev).float()
        ref = x.float() @ w + bias.float()
        gemm_report(f"fused M={M} K={K} N={N} gs={gs}", y, ref)
    for bits in (8, 3, 2):
        K, N, gs, M = 256, 256, 128, 300
        qw, qz, sc, gi = rand_packed(K, N, bits, gs, seed=2, g_idx=True)
        x = torch.randn(M, K, device=dev).half()
        y = ops.qlinear(x, t(qw), t(qz), t(sc), bits, gs, None, g_idx=t(gi))
        w = torch.from_numpy(oq.dequant(qw, qz, sc, bits, gs, gi)).to(dev).float()
        gemm_report(f"qlinear bits={bits} g_idx", y, x.float() @ w)
    for (M, K, N) in [(4096, 1280, 5120), (4096, 5120, 1280), (32768, 1280, 5120), (32768, 1280, 3840), (32768, 5120, 1280), (4096, 1280, 3840)]:
        qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=1)
        x = torch.randn(M, K, device=dev).half(); bias = torch.randn(N, device=dev).half()
        tq, tz, ts = t(qw), t(qz), t(sc)
        for epi in (_lib.EPI_NONE, _lib.EPI_GELU):
            for _ in range(3): ops.qlinear(x, tq, tz, ts, 4, 128, bias, epilogue=epi)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.qlinear(x, tq, tz, ts, 4, 128, bias, epilogue=epi)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"fused M={M} K={K} N={N} epi={epi}: {ms*1e3:.1f} us  {2*M*K*N/ms/1e9:.1f} TFLOP/s")

if __name__ == "__main__":
    which = sys.argv[1]
    print("=== probe", which, torch.cuda.get_device_name(0), flush=True)
    {"dequant": probe_dequant, "ln": probe_ln, "dense": probe_dense, "fused": probe_fused}[which]()
    torch.cuda.synchronize()
    print("=== done", which)
