"""ctypes binding of libsamq.so (the C ABI declared in include/samq.h).

There is NO fallback: if the shared library is missing, or a call returns a non-zero
status, an exception is raised.  The product never routes through PyTorch eager ops
or the CPU oracle for the hot path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
#: SAMQ_LIB points at another build of the library (e.g. lib/libsamq_ablations.so from `make ABLATIONS=1`)
LIB_PATH = os.environ.get("SAMQ_LIB") or os.path.join(_HERE, "lib", "libsamq.so")

SAMQ_OK = 0
SAMQ_ERR_BAD_SHAPE = -1
SAMQ_ERR_UNSUPPORTED_BITS = -2
SAMQ_ERR_UNSUPPORTED_ARCH = -3
SAMQ_ERR_LAUNCH = -4
SAMQ_ERR_BAD_ARG = -5

EPI_NONE = 0
EPI_GELU = 1

RELW_REFERENCE = 0
RELW_UPSTREAM = 1

# every symbol include/samq.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "samq_abi_version": (c_int, []),
    "samq_last_error": (c_char_p, []),
    "samq_device_check": (c_int, []),
    "samq_launch_count": (c_uint64, []),
    "samq_config_reload": (None, []),
    "samq_has_ablations": (c_int, []),
    "samq_unpack_dequant": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    ),
    "samq_attn_small_fwd": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "samq_small_linear_fwd": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "samq_gelu_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "samq_syrk_f32_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_float, c_void_p]),
    "samq_gptq_block_fwd": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
         c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "samq_gather_cols_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "samq_qlinear_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p],
    ),
    "samq_qlinear_unpartition_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    ),
    "samq_qlinear_partition_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    ),
    "samq_qlinear_prefetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "samq_dense_linear_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p],
    ),
    "samq_attn_relpos_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    ),
    "samq_attn_relpos_unpartition_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
         c_void_p],
    ),
    "samq_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "samq_layernorm_partition_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    ),
    "samq_unpartition_residual": (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "samq_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "samq_patchify_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "samq_im2col3x3_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


class SamqError(RuntimeError):
    """A libsamq call failed (CUDA launch/driver error)."""


def load() -> ctypes.CDLL:
    """Load libsamq.so (once) and bind every declared symbol.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C sam_quantization_b200/csrc`). There is no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


#: host-side developer switches, resolved once at import / by reload_config() -- never per call
#:   SAMQ_GEMM=fused|dense   force one int4 QuantLinear path (the library reads the same variable)
#:   SAMQ_PAD_SKIP=0         windowed blocks partition first and multiply the zero-padding rows
#:   SAMQ_NECK_CONV=cudnn    neck conv3x3 through cuDNN instead of im2col + tcgen05 GEMM
#:   SAMQ_PREFETCH=0         no weight prefetch: every QuantLinear unpacks its own weight right before its GEMM
OPTIONS = {}


def _read_options() -> None:
    OPTIONS["gemm"] = os.environ.get("SAMQ_GEMM", "")
    OPTIONS["pad_skip"] = os.environ.get("SAMQ_PAD_SKIP", "1") != "0"
    OPTIONS["neck_cudnn"] = os.environ.get("SAMQ_NECK_CONV", "") == "cudnn"
    OPTIONS["prefetch"] = os.environ.get("SAMQ_PREFETCH", "1") != "0"


_read_options()


def reload_config() -> None:
    """Re-read the SAMQ_* developer switches (host side and library side).  Tests / A-B scripts
    call this after changing the environment; the product never does."""
    _read_options()
    load().samq_config_reload()


def has_ablations() -> bool:
    return bool(load().samq_has_ablations())


def last_error() -> str:
    return load().samq_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Map a samq_status to the exception type the reference raises for that failure
    (SURVEY 8(b) 'Error conventions')."""
    if status == SAMQ_OK:
        return
    msg = last_error()
    if status == SAMQ_ERR_BAD_SHAPE:
        raise AssertionError(msg)                      # quant_linear.py:378-399
    if status == SAMQ_ERR_UNSUPPORTED_BITS:
        raise NotImplementedError(msg)                 # quant_linear.py:72-73
    if status == SAMQ_ERR_UNSUPPORTED_ARCH:
        raise RuntimeError(msg)                        # fused_attention.py:314-318
    if status == SAMQ_ERR_BAD_ARG:
        raise ValueError(msg)
    raise SamqError(msg)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: sam_quantization_b200 has no CPU path "
            f"(got device {t.device})"
        )


_device_checked = set()


def device_check(device: torch.device) -> None:
    """Fail loudly unless `device` is an sm_100 GPU (once per device)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_checked:
        return
    with torch.cuda.device(idx):
        check(load().samq_device_check())
    _device_checked.add(idx)


def launch_count() -> int:
    return int(load().samq_launch_count())
