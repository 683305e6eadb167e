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

def attn_ref(qkv, rph, rpw, B, E, heads, scale, upstream=False):
    S = E * E
    hd = qkv.shape[-1] // 3 // heads
    x = qkv.float().view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, B * heads, S, hd)
    q, k, v = x[0], x[1], x[2]
    idx = (torch.arange(E)[:, None] - torch.arange(E)[None, :] + (E - 1)).to(qkv.device)
    Rh = rph.float()[idx]; Rw = rpw.float()[idx]
    r_q = q.view(B * heads, E, E, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh).half().float()
    rel_w = torch.einsum("bhwc,wkc->bhwk" if upstream else "bhwc,hkc->bhwk", r_q, Rw).half().float()
    attn = (q * scale) @ k.transpose(-2, -1)
    attn = (attn.view(B * heads, E, E, E, E) + rel_h[..., :, None] + rel_w[..., None, :]).view(B * heads, S, S)
    attn = attn.softmax(-1)
    return (attn @ v).view(B, heads, E, E, hd).permute(0, 2, 3, 1, 4).reshape(B, E, E, heads * hd)

def probe_attn():
    torch.manual_seed(20)
    for (B, E, heads, hd) in [(3, 14, 2, 64), (3, 14, 2, 80), (1, 64, 2, 64), (1, 64, 2, 80), (25, 14, 16, 80), (1, 64, 16, 80), (2, 64, 12, 64)]:
        for upstream in (False, True):
            qkv = (torch.randn(B, E * E, 3 * heads * hd, device=dev) * 0.5).half()
            rph = (torch.randn(2 * E - 1, hd, device=dev) * 0.3).half()
            rpw = (torch.randn(2 * E - 1, hd, device=dev) * 0.3).half()
            scale = hd ** -0.5
            out = ops.attn_relpos(qkv, rph, rpw, B, E, E, heads, scale, 1 if upstream else 0)
            torch.cuda.synchronize()
            ref = attn_ref(qkv, rph, rpw, B, E, heads, scale, upstream)
            d = (out.float() - ref).abs()
            cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
            print(f"attn B={B} E={E} heads={heads} hd={hd} upstream={upstream}: maxabs {d.max().item():.4e} refmax {ref.abs().max().item():.3f} cos {cos:.7f} nan {torch.isnan(out).sum().item()}", flush=True)
            if not cos > 0.999:
                bad = (d > 0.05).float().view(B, E * E, heads, hd)
                print("   bad frac", bad.mean().item(), "by head", bad.mean((0, 1, 3)).tolist()[:4], "by d (first 8, last 8)", bad.mean((0, 1, 2))[:8].tolist(), bad.mean((0, 1, 2))[-8:].tolist())
                print("   bad by row-block of 32:", bad.mean((0, 2, 3)).view(-1, 32).mean(1)[:8].tolist())
    for (B, E, heads, hd) in [(8, 64, 16, 80), (200, 14, 16, 80), (8, 64, 16, 64), (200, 14, 16, 64)]:
        qkv = (torch.randn(B, E * E, 3 * heads * hd, device=dev) * 0.5).half()
        rph = (torch.randn(2 * E - 1, hd, device=dev) * 0.3).half(); rpw = (torch.randn(2 * E - 1, hd, device=dev) * 0.3).half()
        for _ in range(3): ops.attn_relpos(qkv, rph, rpw, B, E, E, heads, hd ** -0.5)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.attn_relpos(qkv, rph, rpw, B, E, E, heads, hd ** -0.5)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4 * (E * E) ** 2 * hd * heads * B
        print(f"attn B={B} E={E} heads={heads} hd={hd}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")

if __name__ == "__main__":
    which = sys.argv[1]
    print("=== probe", which, torch.cuda.get_device_name(0), flush=True)
    {"dequant": probe_dequant, "ln": probe_ln, "dense": probe_dense, "fused": probe_fused, "attn": probe_attn}[which]()
    torch.cuda.synchronize()
    print("=== done", which)
