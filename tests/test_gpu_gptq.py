"""GPTQ solver on the B200: Hessian on the tensor cores (samq_syrk_f32_fwd) and
the blocked rounding loop as a kernel (samq_gptq_block_fwd), pinned to the outputs of the
REFERENCE's GPTQ.add_batch / fasterquant (tests/golden/gptq_*.npz, written by
tests/golden/make_gptq_fixtures.py from /root/reference/gptq.py on the CPU).

Tolerance as for the CPU solver (tests/test_gptq_solver.py): H rtol 1e-5 at the fixtures' 150-200
tokens (1e-4 of the largest entry at 4096+ tokens: tensor-core fp32 accumulation); rounded weights may differ
from the fixture only where a different summation order moves a value across a rounding boundary:
<= 0.5 % of the entries, each by one grid step."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from sam_quantization_b200 import _lib, ops
from sam_quantization_b200 import gptq as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("C,tokens,dtype", [(128, 150, torch.float32), (192, 200, torch.float32),
                                            (1280, 4096, torch.float16), (1280, 4900, torch.float32)])
def test_hessian_on_the_tensor_cores(cuda_device, C, tokens, dtype):
    g = torch.Generator().manual_seed(C + tokens)
    x = torch.randn(tokens, C, generator=g).to(dtype)
    ref = torch.zeros(C, C, dtype=torch.float64)
    H = torch.zeros(C, C, device=cuda_device)
    for n in range(2):                                        # two calls: exercises beta
        xs = x * (n + 1)
        ref = ref * (n / (n + 1)) + (2.0 / (n + 1)) * xs.double().t() @ xs.double()
        ops.hessian_accumulate(H, xs.to(cuda_device), 2.0 / (n + 1), n / (n + 1))
    err = (H.cpu().double() - ref).abs().max().item()
    # fp32 accumulation inside the tensor core over up to 4900 products per entry (not an IEEE
    # sequential sum): measured 2e-5 of the largest entry at 4096 tokens, 1e-6 at 200
    assert err <= (1e-4 if tokens > 1000 else 2e-6) * ref.abs().max().item(), (err, ref.abs().max().item())
    assert torch.equal(H, H.t()) or (H - H.t()).abs().max().item() <= 1e-6 * ref.abs().max().item()


@pytest.mark.parametrize("name", ["g64", "g32_inblock", "perrow_b3", "actorder_b8"])
def test_device_solver_matches_the_reference_fasterquant(cuda_device, name):
    f = dict(np.load(os.path.join(GOLD, f"gptq_{name}.npz")))
    bits, groupsize, blocksize, actorder, sym = (int(v) for v in f["cfg"])
    rows, cols = f["W0"].shape
    lin = nn.Linear(cols, rows, bias=False)
    lin.weight.data = torch.from_numpy(f["W0"]).clone()
    lin = lin.to(cuda_device)
    s = G.GPTQ(lin)
    assert s._on_device()
    s.quantizer = G.Quantizer()
    s.quantizer.configure(bits, perchannel=True, sym=bool(sym), mse=False)
    launches = _lib.launch_count()
    for x in torch.from_numpy(f["X"]):
        s.add_batch(x.unsqueeze(0).to(cuda_device))
    assert _lib.launch_count() - launches == 3 * len(f["X"])     # fp32 input: hi.hi, hi.lo, lo.hi per call
    assert torch.allclose(s.H.cpu(), torch.from_numpy(f["H"]), rtol=1e-5, atol=1e-5 * float(np.abs(f["H"]).max()))
    launches = _lib.launch_count()
    scale, zero = s.fasterquant(blocksize=blocksize, percdamp=0.01, groupsize=groupsize, actorder=bool(actorder))
    assert _lib.launch_count() - launches == -(-cols // blocksize)   # one kernel per column block
    Q, Qr = lin.weight.data.cpu().numpy(), f["Q"]
    step = float(f["scale"].max())
    diff = np.abs(Q - Qr)
    assert (diff > 1e-6).mean() <= 5e-3 and diff.max() <= 1.01 * step, ((diff > 1e-6).mean(), diff.max(), step)
    assert np.allclose(scale.cpu().numpy(), f["scale"], rtol=1e-3, atol=1e-6)
    assert (np.abs(zero.cpu().numpy() - f["zero"]) > 0).mean() <= 2e-2
    assert (s.g_idx is not None) == bool(actorder and groupsize != -1)


def test_device_solver_on_a_vith_layer_equals_the_host_solver(cuda_device):
    """ViT-H proj-sized layer (1280 -> 1280), one image's 4096 tokens in fp16 (the calibration
    protocol of gptq4sam.py on a half model): the device solver against this package's host solver
    (itself pinned to the reference by tests/test_gptq_solver.py)."""
    torch.manual_seed(0)
    C = 1280
    w = torch.randn(C, C) * 0.02
    x = (torch.randn(4096, C) * torch.linspace(0.2, 2.0, C)).half()
    out = {}
    for dev in ("cpu", cuda_device):
        lin = nn.Linear(C, C, bias=False)
        lin.weight.data = w.clone()
        lin = lin.to(dev)
        s = G.GPTQ(lin)
        s.quantizer = G.Quantizer()
        s.quantizer.configure(4, perchannel=True, sym=False, mse=False)
        s.add_batch(x.unsqueeze(0).to(dev))
        H = s.H.clone().cpu()
        scale, zero = s.fasterquant(blocksize=128, percdamp=0.01, groupsize=128)
        out[str(dev)] = (H, lin.weight.data.cpu(), scale.cpu(), s.error)
    (Hc, Qc, sc, ec), (Hg, Qg, sg, eg) = out["cpu"], out[str(cuda_device)]
    assert torch.allclose(Hg, Hc, rtol=1e-4, atol=1e-4 * Hc.abs().max().item())
    flips = ((Qc - Qg).abs() > 1e-6).float().mean().item()
    assert flips <= 5e-3, flips
    assert (Qc - Qg).abs().max().item() <= 1.01 * sc.max().item()
    assert abs(ec - eg) <= 1e-2 * abs(ec)
