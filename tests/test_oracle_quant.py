"""Pin the numpy oracle (oracle/quant.py) to the reference's own outputs (fixtures made by
tests/golden/make_golden.py from /root/reference) and check its internal identities."""
import os

import numpy as np
import pytest

from oracle import quant as oq


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_rtn_matches_reference_quantizer(golden_dir, bits):
    """find_params / quantize == gptq.py Quantizer (perchannel, asym, no mse) per group."""
    g = np.load(os.path.join(golden_dir, f"pack_b{bits}.npz"))
    wfake, scale, zero = oq.rtn_quantize(g["weight16"].astype(np.float32), bits, int(g["groupsize"]))
    assert bits_equal(scale, g["scale"])
    assert bits_equal(zero, g["zero"])
    assert bits_equal(wfake.astype(np.float16), g["wfake16"])


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_pack_bit_identical_to_reference_pack_linear(golden_dir, bits):
    g = np.load(os.path.join(golden_dir, f"pack_b{bits}.npz"))
    p = oq.pack(g["wfake16"], g["scale"], g["zero"], bits, int(g["groupsize"]))
    assert bits_equal(p["qweight"], g["qweight"])
    assert bits_equal(p["qzeros"], g["qzeros"])
    assert bits_equal(p["scales"], g["scales"])


def test_zero_minus_one_quirk_is_reproduced(golden_dir):
    """SURVEY trap 8: a group with zero == 0 stores -1, which sign-fills its qzeros word."""
    g = np.load(os.path.join(golden_dir, "pack_b4.npz"))
    rows, groups = np.nonzero(g["zero"] == 0)
    assert len(rows) >= 1
    n, grp = int(rows[0]), int(groups[0])
    word = int(g["qzeros"][grp, n // 8].view(np.uint32))
    # all fields from n%8 upwards are 0xF (sign fill of -1)
    for j in range(n % 8, 8):
        assert (word >> (4 * j)) & 0xF == 0xF
    z = oq.unpack_qzeros(g["qzeros"], 4, g["qzeros"].shape[1] * 8)
    assert z[grp, n] == 15   # (-1) & 0xF: dequant then uses z+1 == 16, not 0


@pytest.mark.parametrize("tag", ["golden_g128", "golden_nogroups"])
def test_dequant_fma_form_is_the_reference_triton_kernel(golden_dir, tag):
    """The fixture is the output of the reference's OWN kernel on a B200:
    triton_matmul4(gs, I_K, qweight, scales, qzeros) (quant_linear.py:355-437) run by
    oracle/ref_gpu.py -- grouped (g128) and NO_GROUPS code paths.  The oracle's default form must
    reproduce it bit for bit; the un-contracted PyTorch form must not (it is a different rounding)."""
    g = np.load(os.path.join(golden_dir, "dequant_triton_b4.npz"))
    args = (g[f"{tag}_qweight"], g[f"{tag}_qzeros"], g[f"{tag}_scales"], 4, int(g[f"{tag}_groupsize"]))
    assert bits_equal(oq.dequant(*args), g[f"{tag}_w_triton"])
    assert bits_equal(oq.dequant(*args, form="fma"), g[f"{tag}_w_triton"])
    step = oq.dequant(*args, form="stepwise")
    assert np.mean(step.view(np.uint16) != g[f"{tag}_w_triton"].view(np.uint16)) > 0.05


def test_dequant_stepwise_is_the_literal_torch_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "dequant_b4.npz"))
    w = oq.dequant(g["qweight"], g["qzeros"], g["scales"], 4, int(g["groupsize"]), form="stepwise")
    assert bits_equal(w, g["w"])


def test_dequant_forms_differ_by_at_most_one_ulp_of_the_product(golden_dir):
    """SURVEY trap 7: the three rounding forms are different definitions (many elements
    differ) but never by more than one fp16 ulp of the larger intermediate (16 * scale)."""
    g = np.load(os.path.join(golden_dir, "dequant_b4.npz"))
    a = oq.dequant(g["qweight"], g["qzeros"], g["scales"], 4, int(g["groupsize"]), form="stepwise")
    bound = 2.0 ** -10 * 16 * float(g["scales"].astype(np.float32).max())
    for form in ("fma", "single"):
        b = oq.dequant(g["qweight"], g["qzeros"], g["scales"], 4, int(g["groupsize"]), form=form)
        assert np.max(np.abs(a.astype(np.float32) - b.astype(np.float32))) <= bound
        assert np.mean(a.view(np.uint16) != b.view(np.uint16)) > 0.05   # they are NOT the same definition


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("act_order", [False, True])
def test_unpack_pack_identity(bits, act_order):
    """Extension formats (3-bit, g_idx) have no reference pin: parity unpinned by reference;
    checked through the pack -> unpack round trip on the integer grid."""
    rng = np.random.default_rng(bits)
    n, k, gs = 96, 256, 64
    w = (rng.standard_normal((n, k)) * 0.02).astype(np.float32)
    g_idx = None
    if act_order:
        perm = rng.permutation(k)
        inv = np.empty(k, dtype=np.int64)
        inv[perm] = np.arange(k)
        g_idx = (inv // gs).astype(np.int32)
        wp, scale, zero = oq.rtn_quantize(w[:, perm], bits, gs)
        wf = np.empty_like(wp)
        wf[:, perm] = wp
    else:
        wf, scale, zero = oq.rtn_quantize(w, bits, gs)
    zero = np.maximum(zero, 1)   # keep zero-1 >= 0 (the quirk has its own test)
    gi = oq.default_g_idx(k, gs) if g_idx is None else g_idx
    wf = (scale[:, gi] * (np.clip(np.round(w / scale[:, gi]) + zero[:, gi], 0, 2**bits - 1) - zero[:, gi])).astype(np.float32)
    p = oq.pack(wf, scale, zero, bits, gs, g_idx)
    q = oq.unpack_qweight(p["qweight"], bits, k)
    z = oq.unpack_qzeros(p["qzeros"], bits, n)
    expect_q = np.round(wf.T / scale.T[gi] + zero.T[gi]).astype(np.int64)
    assert np.array_equal(q.astype(np.int64), expect_q)
    assert np.array_equal(z.astype(np.int64) + 1, zero.T.astype(np.int64))
    assert p["qweight"].shape == (k * bits // 32, n) and p["qzeros"].shape == (k // gs, n * bits // 32)
    # dequantised weight reproduces the fake-quantised one up to fp16 rounding of scale
    wd = oq.dequant(p["qweight"], p["qzeros"], p["scales"], bits, gs, g_idx).astype(np.float32)
    assert np.max(np.abs(wd.T - wf)) < 1e-3 * (2**bits)


def test_three_bit_layout_matches_quant3linear_stream():
    """quant.py:160-180: v0..v9 @3i in word0, v10 low 2 bits @30; v10 bit2 @0 of word1 ..."""
    vals = np.arange(32, dtype=np.int64)[:, None] % 8
    words = oq._pack_fields(vals, 3).view(np.uint32)[:, 0]
    w0 = sum(int(vals[j, 0]) << (3 * j) for j in range(10)) | ((int(vals[10, 0]) & 3) << 30)
    w1 = ((int(vals[10, 0]) >> 2) & 1) | sum(int(vals[11 + j, 0]) << (3 * j + 1) for j in range(10)) | ((int(vals[21, 0]) & 1) << 31)
    w2 = ((int(vals[21, 0]) >> 1) & 3) | sum(int(vals[22 + j, 0]) << (3 * j + 2) for j in range(10))
    assert [int(w) for w in words] == [w0 & 0xFFFFFFFF, w1 & 0xFFFFFFFF, w2 & 0xFFFFFFFF]
    assert np.array_equal(oq._unpack_fields(words[:, None].view(np.int32), 3, 32)[:, 0], vals[:, 0])


def test_empty_and_shape_edges():
    # groupsize -1 == one group over all of K
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((32, 64)) * 0.02).astype(np.float32)
    wf, s, z = oq.rtn_quantize(w, 4, -1)
    assert s.shape == (32, 1)
    p = oq.pack(wf, s, z, 4, -1)
    assert p["qzeros"].shape == (1, 4) and p["scales"].shape == (1, 32)
    x = np.zeros((0, 64), dtype=np.float32)   # empty batch
    y = oq.qlinear(x, p["qweight"], p["qzeros"], p["scales"], 4, -1)
    assert y.shape == (0, 32)
