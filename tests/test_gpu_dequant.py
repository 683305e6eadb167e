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


@pytest.mark.parametrize("tag", ["golden_g128", "golden_nogroups"])
@pytest.mark.parametrize("transposed", [False, True])
def test_dequant_equals_the_reference_triton_kernel(cuda_device, golden_dir, tag, transposed):
    """Bit-equality with the reference's own kernel: the fixture is triton_matmul4(gs, I_K, ...) run
    on a B200 by oracle/ref_gpu.py (quant_linear.py:355-437).  Both unpack kernels (generic and the
    transposed int4 fast path that feeds the dense GEMM) are held to it."""
    g = np.load(os.path.join(golden_dir, "dequant_triton_b4.npz"))
    w = ops.unpack_dequant(dev(g[f"{tag}_qweight"], cuda_device), dev(g[f"{tag}_qzeros"], cuda_device),
                           dev(g[f"{tag}_scales"], cuda_device), 4, int(g[f"{tag}_groupsize"]), transposed=transposed)
    assert same_bits(w.t().contiguous() if transposed else w, g[f"{tag}_w_triton"])


@pytest.mark.parametrize("tag", ["golden_g128", "golden_nogroups"])
def test_fused_gemm_operand_equals_the_reference_triton_kernel(cuda_device, golden_dir, tag, samq_env):
    """The same identity-matrix extraction through OUR fused in-SM dequant GEMM (the TMEM A operand):
    qlinear(I_K) must return the Triton kernel's weights bit for bit."""
    samq_env.set("SAMQ_GEMM", "fused")
    g = np.load(os.path.join(golden_dir, "dequant_triton_b4.npz"))
    qw, gs = g[f"{tag}_qweight"], int(g[f"{tag}_groupsize"])
    K = qw.shape[0] * 8
    eye = torch.eye(K, dtype=torch.float16, device=cuda_device)
    w = ops.qlinear(eye, dev(qw, cuda_device), dev(g[f"{tag}_qzeros"], cuda_device), dev(g[f"{tag}_scales"], cuda_device),
                    4, gs)
    assert same_bits(w, g[f"{tag}_w_triton"])


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
