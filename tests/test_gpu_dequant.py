"""samq_unpack_dequant (through the C ABI) vs the oracle: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import quant as oq
from sam_quantization_b200 import ops
from gpu_util import dev, rand_packed

pytestmark = pytest.mark.gpu


def same_bits(t, ref):
    return np.array_equal(t.cpu().numpy().view(np.uint16), np.ascontiguousarray(ref).view(np.uint16))


def test_golden_dequant_fixture(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "dequant_b4.npz"))
    w = ops.unpack_dequant(dev(g["qweight"], cuda_device), dev(g["qzeros"], cuda_device), dev(g["scales"], cuda_device),
                           4, int(g["groupsize"]))
    assert same_bits(w, g["w"])


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("g_idx", [False, True])
@pytest.mark.parametrize("transposed", [False, True])
def test_dequant_bit_exact(cuda_device, bits, g_idx, transposed):
    K, N, gs = 512, 384, 128
    qw, qz, sc, gi = rand_packed(K, N, bits, gs, seed=10 + bits, g_idx=g_idx, scale_lo=1e-4)
    ref = oq.dequant(qw, qz, sc, bits, gs, gi)
    w = ops.unpack_dequant(dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), bits, gs,
                           dev(gi, cuda_device), transposed=transposed)
    assert same_bits(w.t().contiguous() if transposed else w, ref)


@pytest.mark.parametrize("K,N,gs", [(1280, 3840, 128), (5120, 1280, 128), (1280, 5120, -1)])
def test_dequant_vith_layers_bit_exact(cuda_device, K, N, gs):
    g = K if gs == -1 else gs
    qw, qz, sc, _ = rand_packed(K, N, 4, g, seed=3)
    ref = oq.dequant(qw, qz, sc, 4, gs)
    w = ops.unpack_dequant(dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), 4, gs)
    assert same_bits(w, ref)


def test_dequant_subnormal_scales_and_zero_quirk(cuda_device):
    """tiny scales (fp16 subnormal products) and all-ones qzeros words (the zero-1 == -1 case)."""
    K, N, gs = 256, 128, 128
    qw, qz, sc, _ = rand_packed(K, N, 4, gs, seed=5, scale_lo=1e-7, scale_hi=1e-5)
    qz[:] = -1
    ref = oq.dequant(qw, qz, sc, 4, gs)
    w = ops.unpack_dequant(dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), 4, gs)
    assert same_bits(w, ref)


def test_dequant_rejects_bad_bits(cuda_device):
    with pytest.raises(NotImplementedError):
        ops.unpack_dequant(torch.zeros(8, 64, dtype=torch.int32, device=cuda_device),
                           torch.zeros(1, 8, dtype=torch.int32, device=cuda_device),
                           torch.zeros(1, 64, dtype=torch.float16, device=cuda_device), 5, 64)
