"""Helpers shared by the -m gpu tests."""
import numpy as np
import torch


def rand_packed(K, N, bits, gs, seed=0, g_idx=False, scale_lo=0.002, scale_hi=0.02):
    """BASELINE config-5 style random packed buffers (uniform random words)."""
    rng = np.random.default_rng(seed)
    G = K // gs
    qweight = rng.integers(-2**31, 2**31, size=(K * bits // 32, N), dtype=np.int64).astype(np.int32)
    qzeros = rng.integers(-2**31, 2**31, size=(G, N * bits // 32), dtype=np.int64).astype(np.int32)
    scales = rng.uniform(scale_lo, scale_hi, size=(G, N)).astype(np.float16)
    gi = None
    if g_idx:
        perm = rng.permutation(K)
        inv = np.empty(K, dtype=np.int64)
        inv[perm] = np.arange(K)
        gi = (inv // gs).astype(np.int32)
    return qweight, qzeros, scales, gi


def dev(a, device):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)


def report(y, ref):
    """(max-abs error, max |ref|, cosine) of a CUDA/CPU tensor against a float reference."""
    yf = y.detach().float().cpu().flatten()
    rf = torch.as_tensor(ref).detach().float().cpu().flatten()
    cos = torch.nn.functional.cosine_similarity(yf.double(), rf.double(), dim=0).item()
    return (yf - rf).abs().max().item(), rf.abs().max().item(), cos
