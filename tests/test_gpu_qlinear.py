"""samq_qlinear_fwd vs the oracle (oracle/quant.py::qlinear, fp32 on the stepwise-dequantised
weight).  Tolerance: the kernel accumulates in fp32 and rounds once to fp16, so
|y - ref| <= 2^-10 * max|ref| (one fp16 ulp at the output's magnitude) and cosine >= 0.99999."""
import numpy as np
import pytest
import torch

import sam_quantization_b200 as sq
from oracle import quant as oq
from sam_quantization_b200 import _lib, ops
from gpu_util import dev, rand_packed, report

pytestmark = pytest.mark.gpu

ULP = 2.0 ** -10


def run_case(device, M, K, N, bits, gs, seed, g_idx=False, epilogue="none", bias=True, residual=False):
    g = K if gs == -1 else gs
    qw, qz, sc, gi = rand_packed(K, N, bits, g, seed=seed, g_idx=g_idx)
    rng = np.random.default_rng(seed + 100)
    x = rng.standard_normal((M, K)).astype(np.float16)
    b = (rng.standard_normal(N)).astype(np.float16) if bias else None
    r = rng.standard_normal((M, N)).astype(np.float16) if residual else None
    y = ops.qlinear(dev(x, device), dev(qw, device), dev(qz, device), dev(sc, device), bits, gs, dev(b, device),
                    dev(gi, device), _lib.EPI_GELU if epilogue == "gelu" else _lib.EPI_NONE, dev(r, device))
    ref = oq.qlinear(x, qw, qz, sc, bits, gs, b, gi, epilogue, r)
    err, mag, cos = report(y, ref)
    assert not torch.isnan(y).any()
    assert err <= ULP * mag + 1e-6, (err, mag)
    assert cos >= 0.99999, cos
    return err, mag, cos


# BASELINE config 5 shapes at the sizes the CPU oracle finishes in seconds
@pytest.mark.parametrize("M", [196, 4096])
@pytest.mark.parametrize("K,N", [(1280, 3840), (1280, 5120), (5120, 1280)])
def test_int4_vith_shapes(cuda_device, M, K, N):
    run_case(cuda_device, M, K, N, 4, 128, seed=1)


@pytest.mark.parametrize("K,N", [(768, 2304), (768, 768), (3072, 768), (1024, 4096), (1280, 1280)])
def test_int4_vit_b_l_h_layer_shapes(cuda_device, K, N):
    run_case(cuda_device, 4900, K, N, 4, 128, seed=2)


@pytest.mark.parametrize("M", [1, 7, 191, 192, 193, 385])
def test_ragged_m(cuda_device, M):
    run_case(cuda_device, M, 256, 256, 4, 128, seed=3)


def test_empty_batch(cuda_device):
    qw, qz, sc, _ = rand_packed(256, 256, 4, 128)
    y = ops.qlinear(torch.zeros(0, 256, dtype=torch.float16, device=cuda_device), dev(qw, cuda_device),
                    dev(qz, cuda_device), dev(sc, cuda_device), 4, 128)
    assert y.shape == (0, 256)


@pytest.mark.parametrize("gs", [64, 128, 256, -1])
def test_groupsizes(cuda_device, gs):
    run_case(cuda_device, 300, 512, 384, 4, gs, seed=4)


@pytest.mark.parametrize("epilogue,bias,residual", [("none", False, False), ("gelu", True, False),
                                                     ("none", True, True), ("gelu", True, True)])
def test_epilogues(cuda_device, epilogue, bias, residual):
    run_case(cuda_device, 1000, 1280, 1280, 4, 128, seed=5, epilogue=epilogue, bias=bias, residual=residual)


@pytest.mark.parametrize("bits", [2, 3, 8])
@pytest.mark.parametrize("g_idx", [False, True])
def test_other_bit_widths_and_act_order(cuda_device, bits, g_idx):
    """BASELINE config 4 (extension: parity unpinned by the reference, checked against the oracle)."""
    run_case(cuda_device, 500, 1280, 1280, bits, 128, seed=6, g_idx=g_idx)


def test_int4_act_order(cuda_device):
    run_case(cuda_device, 500, 1280, 3840, 4, 128, seed=7, g_idx=True)


@pytest.mark.parametrize("K,N", [(1280, 3840), (1280, 5120), (5120, 1280)])
@pytest.mark.parametrize("variant", ["auto", "fused"])
def test_full_size_m32768_against_fp32_gemm_of_exact_weights(cuda_device, samq_env, K, N, variant):
    """At BASELINE's largest M the oracle is replaced by a size-independent identity:
    the fused kernel must equal an fp32 GEMM (torch, on the GPU) of the bit-exact dequantised
    weight (itself pinned to the oracle by test_gpu_dequant)."""
    M = 32768
    if variant == "fused":
        samq_env.set("SAMQ_GEMM", "fused")
    else:
        samq_env.unset("SAMQ_GEMM")
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=8)
    tq, tz, ts = dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device)
    x = torch.randn(M, K, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(0)).half()
    y = ops.qlinear(x, tq, tz, ts, 4, 128)
    w = ops.unpack_dequant(tq, tz, ts, 4, 128).float()
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = x.float() @ w
    err = (y.float() - ref).abs().max().item()
    mag = ref.abs().max().item()
    assert err <= ULP * mag, (err, mag)
    # linearity in x (size-independent property): f(2x) == 2 f(x) exactly in fp16 (power-of-two
    # scaling commutes with every rounding) wherever the output is a normal fp16 number;
    # subnormal outputs may differ by one subnormal step (2^-24)
    y2 = ops.qlinear((x * 2).half(), tq, tz, ts, 4, 128)
    normal = y.abs() >= 2.0 ** -13
    assert torch.equal(y2[normal], (y * 2).half()[normal])
    assert (y2.float() - 2 * y.float()).abs().max().item() <= 2.0 ** -23


@pytest.mark.parametrize("K,N,epilogue,residual", [(1280, 3840, "none", False), (1280, 1280, "none", True),
                                                   (1280, 5120, "gelu", False), (5120, 1280, "none", True)])
def test_the_four_vith_linears_at_the_benchmarked_batch(cuda_device, K, N, epilogue, residual):
    """The bench's own GEMMs: M = 131072 rows (batch 32), each ViT-H layer shape with its epilogue.
    (1) 2048-row slices (first, a middle one, last) recomputed alone -- same kernel path -- give the
    same bits: no tile depends on another.  (2) The same slices against an fp32 GEMM of the bit-exact
    dequantised weight with the epilogue applied in fp32 (one fp16 ulp of the output's magnitude;
    GELU is the exact erf form)."""
    M = 131072
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=14)
    tq, tz, ts = dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device)
    g = torch.Generator(cuda_device).manual_seed(3)
    x = torch.randn(M, K, device=cuda_device, generator=g).half()
    b = torch.randn(N, device=cuda_device, generator=g).half()
    r = torch.randn(M, N, device=cuda_device, generator=g).half() if residual else None
    epi = _lib.EPI_GELU if epilogue == "gelu" else _lib.EPI_NONE
    y = ops.qlinear(x, tq, tz, ts, 4, 128, b, epilogue=epi, residual=r)
    w = ops.unpack_dequant(tq, tz, ts, 4, 128).float()
    torch.backends.cuda.matmul.allow_tf32 = False
    for lo in (0, 61440 + 256, M - 2048):
        sl = slice(lo, lo + 2048)
        alone = ops.qlinear(x[sl].contiguous(), tq, tz, ts, 4, 128, b, epilogue=epi,
                            residual=None if r is None else r[sl].contiguous())
        assert torch.equal(y[sl], alone), lo
        ref = x[sl].float() @ w + b.float()
        if epilogue == "gelu":
            ref = torch.nn.functional.gelu(ref)
        if r is not None:
            ref = ref + r[sl].float()
        err = (y[sl].float() - ref).abs().max().item()
        mag = ref.abs().max().item()
        assert err <= ULP * mag + 1e-6, (lo, err, mag)


def test_module_forward_and_reference_entry_point(cuda_device):
    """QuantLinear.forward and triton_matmul4 (the reference's public entry, quant_linear.py:355)."""
    K, N, gs = 1280, 1280, 128
    qw, qz, sc, _ = rand_packed(K, N, 4, gs, seed=9)
    m = sq.QuantLinear(4, gs, K, N, True)
    m.qweight, m.qzeros, m.scales = torch.from_numpy(qw), torch.from_numpy(qz), torch.from_numpy(sc)
    m.bias = torch.randn(N).half()
    m = m.to(cuda_device)
    x = torch.randn(2, 14, 14, K, device=cuda_device).half()
    y = m(x)
    assert y.shape == (2, 14, 14, N) and y.dtype == torch.float16
    y2 = sq.triton_matmul4(gs, x, m.qweight, m.scales, m.qzeros, m.bias)
    assert torch.equal(y, y2)
    ref = oq.qlinear(x.cpu().numpy().reshape(-1, K), qw, qz, sc, 4, gs, m.bias.cpu().numpy())
    err, mag, cos = report(y.view(-1, N), ref)
    assert err <= ULP * mag and cos >= 0.99999
    with pytest.raises(AssertionError):
        m(x[..., :640].contiguous())
    with pytest.raises(AssertionError):
        m(x.transpose(1, 2))          # non-contiguous, quant_linear.py:381


@pytest.mark.parametrize("K,N,M", [(1280, 3840, 4900), (5120, 1280, 4096), (1280, 5120, 777)])
@pytest.mark.parametrize("epilogue", ["none", "gelu"])
def test_cta_pair_kernel_equals_single_cta_kernel(cuda_device, samq_env, K, N, M, epilogue):
    """The cta_group::2 kernel (SAMQ_GEMM=2cta) must give bit-identical results to the default
    single-CTA kernel (same operands, same fp32 accumulation order per output)."""
    if not _lib.has_ablations():
        pytest.skip("the cta_group::2 fused kernel is only in `make ABLATIONS=1` builds (SAMQ_LIB=...)")
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=12)
    tq, tz, ts = dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device)
    x = torch.randn(M, K, device=cuda_device).half()
    b = torch.randn(N, device=cuda_device).half()
    r = torch.randn(M, N, device=cuda_device).half()
    epi = _lib.EPI_GELU if epilogue == "gelu" else _lib.EPI_NONE
    samq_env.unset("SAMQ_GEMM")
    y1 = ops.qlinear(x, tq, tz, ts, 4, 128, b, epilogue=epi, residual=r)
    samq_env.set("SAMQ_GEMM", "2cta")
    y2 = ops.qlinear(x, tq, tz, ts, 4, 128, b, epilogue=epi, residual=r)
    assert torch.equal(y1, y2)
    ref = oq.qlinear(x.cpu().numpy(), qw, qz, sc, 4, 128, b.cpu().numpy(), None, epilogue, r.cpu().numpy())
    err, mag, cos = report(y2, ref)
    assert err <= ULP * mag + 1e-6 and cos >= 0.99999


@pytest.mark.parametrize("K,N,gs", [(1280, 1280, 128), (320, 256, 64), (64, 512, 64), (192, 256, 64)])
def test_size_dispatch_two_kernel_path_equals_fused(cuda_device, samq_env, K, N, gs):
    """M >= 2048 takes unpack-once + dense GEMM; it must equal the fused kernel bit for bit.
    K = 320, 64, 192: short reductions (fewer k-blocks than pipeline stages, odd counts)."""
    M = 12288 + 77
    qw, qz, sc, _ = rand_packed(K, N, 4, gs, seed=13)
    tq, tz, ts = dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device)
    x = torch.randn(M, K, device=cuda_device).half()
    b = torch.randn(N, device=cuda_device).half()
    samq_env.unset("SAMQ_GEMM")
    y_auto = ops.qlinear(x, tq, tz, ts, 4, gs, b, epilogue=_lib.EPI_GELU)
    samq_env.set("SAMQ_GEMM", "fused")
    y_fused = ops.qlinear(x, tq, tz, ts, 4, gs, b, epilogue=_lib.EPI_GELU)
    assert torch.equal(y_auto, y_fused)
    # and the fast transposed int4 dequant kernel is bit-exact against the oracle
    wt = ops.unpack_dequant(tq, tz, ts, 4, gs, transposed=True)
    ref = oq.dequant(qw, qz, sc, 4, gs)
    assert np.array_equal(wt.t().contiguous().cpu().numpy().view(np.uint16), ref.view(np.uint16))


def test_dense_ablation_path_equals_fused(cuda_device):
    K, N = 1280, 3840
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=11)
    tq, tz, ts = dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device)
    x = torch.randn(4900, K, device=cuda_device).half()
    y = ops.qlinear(x, tq, tz, ts, 4, 128)
    wt = ops.unpack_dequant(tq, tz, ts, 4, 128, transposed=True)
    yd = ops.dense_linear(x, wt)
    assert torch.equal(y, yd)      # same fp16 operands, same fp32 accumulation order per tile


@pytest.mark.parametrize("B,H,W,K,N,variant", [(1, 64, 64, 1280, 1280, "auto"), (3, 64, 64, 256, 256, "auto"),
                                               (2, 20, 30, 128, 128, "auto"), (3, 64, 64, 1280, 1280, "dense"),
                                               (2, 64, 64, 768, 768, "2cta")])
def test_proj_with_fused_unpartition_and_residual(cuda_device, samq_env, B, H, W, K, N, variant):
    """samq_qlinear_unpartition_fwd == shortcut + window_unpartition(x @ W + bias)
    (image_encoder.py:201-204, 309-333), on every GEMM kernel variant."""
    from oracle import encoder as oe
    if variant == "2cta" and not _lib.has_ablations():
        pytest.skip("the cta_group::2 fused kernel is only in `make ABLATIONS=1` builds")
    if variant == "auto":
        samq_env.unset("SAMQ_GEMM")
    else:
        samq_env.set("SAMQ_GEMM", variant)
    ws = 14
    nH, nW = (H + ws - 1) // ws, (W + ws - 1) // ws
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=21)
    g = torch.Generator().manual_seed(5)
    xw = torch.randn(B * nH * nW, ws, ws, K, generator=g).half()
    sc_t = torch.randn(B, H, W, N, generator=g).half()
    b = torch.randn(N, generator=g).half()
    y = ops.qlinear_unpartition(xw.to(cuda_device), dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device),
                                4, 128, b.to(cuda_device), sc_t.to(cuda_device), ws)
    lin = torch.from_numpy(oq.qlinear(xw.numpy().reshape(-1, K), qw, qz, sc, 4, 128, b.numpy())).view(-1, ws, ws, N)
    ref = sc_t.float() + oe.window_unpartition(lin.half().float(), ws, (nH * ws, nW * ws), (H, W))
    err, mag, cos = report(y, ref)
    assert y.shape == (B, H, W, N)
    assert err <= 2 * ULP * mag and cos >= 0.99999


@pytest.mark.parametrize("B,H,W,K,N,variant", [(3, 64, 64, 1280, 3840, "auto"), (1, 64, 64, 256, 768, "auto"),
                                               (2, 20, 30, 128, 384, "auto"), (2, 28, 28, 128, 256, "auto"),
                                               (3, 64, 64, 256, 768, "dense"), (2, 64, 64, 256, 768, "2cta")])
@pytest.mark.parametrize("with_bias", [True, False])
def test_qkv_with_fused_partition(cuda_device, samq_env, B, H, W, K, N, variant, with_bias):
    """samq_qlinear_partition_fwd == window_partition(x) @ W + bias (image_encoder.py:196-198,
    282-306) on every GEMM kernel variant -- bit-identical to partitioning first, because the real
    rows see the same dot products and a zero-padding row's result is exactly fp16(0 + bias)."""
    from oracle import encoder as oe
    if variant == "2cta" and not _lib.has_ablations():
        pytest.skip("the cta_group::2 fused kernel is only in `make ABLATIONS=1` builds")
    if variant == "auto":
        samq_env.unset("SAMQ_GEMM")
    else:
        samq_env.set("SAMQ_GEMM", variant)
    ws = 14
    qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=44)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, H, W, K, generator=g).half()
    b = torch.randn(N, generator=g).half() if with_bias else None
    packed = (dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), 4, 128)
    bd = b.to(cuda_device) if with_bias else None
    y = ops.qlinear_partition(x.to(cuda_device), *packed, bd, ws)
    xw, _ = oe.window_partition(x.float(), ws)
    xw = xw.half().contiguous()
    ref = ops.qlinear(xw.to(cuda_device), *packed, bd)
    assert y.shape == ref.shape
    assert torch.equal(y, ref)


@pytest.mark.parametrize("bits", [2, 3, 8])
@pytest.mark.parametrize("gs", [128, 64, -1])
def test_fused_in_sm_unpack_of_every_format_is_bit_exact(cuda_device, samq_env, bits, gs):
    """The fused kernel's in-register unpackers (qlinear_common.cuh::unpack_kblock<2|3|8>): the
    identity-matrix extraction qlinear(I_K) must return the oracle's dequantised weight bit for bit,
    exactly as for int4 (which is pinned to the Triton kernel in test_gpu_dequant)."""
    samq_env.set("SAMQ_GEMM", "fused")
    K, N = 512, 384
    g = K if gs == -1 else gs
    qw, qz, sc, _ = rand_packed(K, N, bits, g, seed=20 + bits, scale_lo=1e-4)
    eye = torch.eye(K, dtype=torch.float16, device=cuda_device)
    launches = _lib.launch_count()
    w = ops.qlinear(eye, dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), bits, gs)
    assert _lib.launch_count() - launches == 1            # one kernel: no unpack pass
    ref = oq.dequant(qw, qz, sc, bits, g)
    assert np.array_equal(w.cpu().numpy().view(np.uint16), ref.view(np.uint16))


@pytest.mark.parametrize("bits", [2, 3, 8])
@pytest.mark.parametrize("M,K,N", [(196, 1280, 3840), (4096, 1280, 1280), (1000, 5120, 1280)])
def test_fused_and_unpack_once_paths_agree_for_every_format(cuda_device, samq_env, bits, M, K, N):
    """BASELINE config 4/5 formats on ViT-H layer shapes: the fused in-SM dequant GEMM against the
    oracle, and against the unpack-once + dense GEMM path on the same operands (same k order in the
    tensor core: identical bits)."""
    samq_env.set("SAMQ_GEMM", "fused")
    run_case(cuda_device, M, K, N, bits, 128, seed=30 + bits)
    qw, qz, sc, _ = rand_packed(K, N, bits, 128, seed=40 + bits)
    x = torch.randn(M, K, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(1)).half()
    b = torch.randn(N, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(2)).half()
    args = (dev(qw, cuda_device), dev(qz, cuda_device), dev(sc, cuda_device), bits, 128, b)
    y_fused = ops.qlinear(x, *args, epilogue=_lib.EPI_GELU)
    samq_env.set("SAMQ_GEMM", "dense")
    y_dense = ops.qlinear(x, *args, epilogue=_lib.EPI_GELU)
    assert torch.equal(y_fused, y_dense)


@pytest.mark.parametrize("bits", [3, 4, 8])
def test_act_order_layer_takes_gather_plus_fused_kernel(cuda_device, samq_env, bits):
    """QuantLinear with a g_idx at short M: rows of qweight sorted by group once (sorted_pack), x's
    columns gathered (samq_gather_cols_fwd), fused in-SM dequant kernel on contiguous groups.  Same
    products as the g_idx-aware unpack + dense path, summed in another order: both within the GEMM
    tolerance of the oracle, and within two output ulps of each other."""
    samq_env.unset("SAMQ_GEMM")
    K, N, gs, M = 1280, 1280, 128, 600
    qw, qz, sc, gi = rand_packed(K, N, bits, gs, seed=50 + bits, g_idx=True)
    layer = sq.QuantLinear(bits, gs, K, N, bias=True)
    rng = np.random.default_rng(5)
    bias = rng.standard_normal(N).astype(np.float16)
    layer.qweight, layer.qzeros, layer.scales = (torch.from_numpy(a) for a in (qw, qz, sc))
    layer.bias = torch.from_numpy(bias)
    layer.g_idx = torch.from_numpy(gi)
    layer = layer.to(cuda_device)
    x = rng.standard_normal((M, K)).astype(np.float16)
    xd = dev(x, cuda_device)
    launches = _lib.launch_count()
    y = layer(xd)
    assert _lib.launch_count() - launches == 2            # gather + fused GEMM (no unpack pass)
    assert torch.equal(ops.gather_cols(xd, layer.sorted_pack()[0]), xd[:, layer.sorted_pack()[0].long()])
    ref = oq.qlinear(x, qw, qz, sc, bits, gs, bias, gi)
    err, mag, cos = report(y, ref)
    assert err <= ULP * mag + 1e-6 and cos >= 0.99999
    y_dense = ops.qlinear(xd, layer.qweight, layer.qzeros, layer.scales, bits, gs, layer.bias, layer.g_idx)
    assert (y.float() - y_dense.float()).abs().max().item() <= 2 * ULP * mag
