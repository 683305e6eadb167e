"""sam_quantization_b200 -- B200-native GPTQ-quantized SAM image-encoder hot path.

Drop-in for the reference's ``gptq_triton`` package (same names and call signatures,
/root/reference/gptq_triton/__init__.py:8-12) backed by hand-written sm_100a CUDA
kernels reached through the C ABI in ``include/samq.h``.
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
