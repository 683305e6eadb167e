"""The C-ABI library loads on a machine without a GPU and exports exactly the symbols
include/samq.h declares; argument validation works before any device call."""
import ctypes
import os
import re

import pytest

from sam_quantization_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built_library():
    """A fresh checkout has no libsamq.so (build artefacts are git-ignored): build it (nvcc
    cross-compiles sm_100a without a GPU) instead of failing on a missing file."""
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "samq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(samq_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert getattr(lib, name) is not None, name


def test_abi_version_and_error_strings():
    lib = _lib.load()
    assert lib.samq_abi_version() == 4
    assert lib.samq_has_ablations() == 0     # the shipped library holds the product kernels only
    # validation happens before any CUDA call, so these run on a GPU-less box
    rc = lib.samq_unpack_dequant(1, 1, 1, None, 1, 64, 64, 5, 64, 0, None)
    assert rc == _lib.SAMQ_ERR_UNSUPPORTED_BITS and "bits" in _lib.last_error()
    rc = lib.samq_qlinear_fwd(16, 16, 16, 16, None, None, None, 16, None, 4, 100, 128, 4, 128, 0, None)
    assert rc == _lib.SAMQ_ERR_BAD_SHAPE and "K=100" in _lib.last_error()
    # weight prefetch: int4 only; and a GEMM call without packed pointers needs the prefetched scratch
    rc = lib.samq_qlinear_prefetch(16, 16, 16, 16, 128, 128, 3, 128, None)
    assert rc == _lib.SAMQ_ERR_UNSUPPORTED_BITS
    rc = lib.samq_qlinear_prefetch(16, 16, 16, 16, 100, 128, 4, 128, None)
    assert rc == _lib.SAMQ_ERR_BAD_SHAPE and "K=100" in _lib.last_error()
    rc = lib.samq_qlinear_fwd(16, None, None, None, None, None, None, 16, None, 4, 128, 128, 4, 128, 0, None)
    assert rc == _lib.SAMQ_ERR_BAD_ARG and "prefetched" in _lib.last_error()
    rc = lib.samq_attn_relpos_fwd(16, 16, 16, 16, 1, 32, 32, 4, 80, 0.1, 0, None)
    assert rc == _lib.SAMQ_ERR_BAD_SHAPE
    rc = lib.samq_layernorm_fwd(None, None, None, None, 4, 64, 1e-6, None)
    assert rc == _lib.SAMQ_ERR_BAD_ARG
    # the re-layout entry points of the stem / neck
    assert lib.samq_im2col3x3_fwd(None, None, 1, 4, 4, 8, None) == _lib.SAMQ_ERR_BAD_ARG
    assert lib.samq_im2col3x3_fwd(16, 16, 1, 4, 4, 12, None) == _lib.SAMQ_ERR_BAD_SHAPE and "C=12" in _lib.last_error()
    assert lib.samq_im2col3x3_fwd(16, 24, 1, 4, 4, 8, None) == _lib.SAMQ_ERR_BAD_ARG          # 16-byte alignment
    assert lib.samq_patchify_fwd(16, 16, 1, 3, 60, 64, 16, None) == _lib.SAMQ_ERR_BAD_SHAPE    # H not a multiple of P
    # the window size of the fused partition / unpartition attention is fixed
    rc = lib.samq_attn_relpos_unpartition_fwd(16, 16, 16, 16, 1, 64, 64, 7, 16, 80, 0.1, 0, None)
    assert rc != _lib.SAMQ_OK and "14" in _lib.last_error()


def test_status_maps_to_reference_exception_types():
    lib = _lib.load()
    with pytest.raises(NotImplementedError):
        _lib.check(lib.samq_unpack_dequant(1, 1, 1, None, 1, 64, 64, 5, 64, 0, None))
    with pytest.raises(AssertionError):
        _lib.check(lib.samq_qlinear_fwd(16, 16, 16, 16, None, None, None, 16, None, 4, 100, 128, 4, 128, 0, None))
    with pytest.raises(ValueError):
        _lib.check(lib.samq_layernorm_fwd(None, None, None, None, 4, 64, 1e-6, None))
