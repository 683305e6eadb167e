"""Recipe: stage the UNMODIFIED reference sources of the hot path under ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (no reference source enters the
history) but travels to the GPU box with the gpurun snapshot, where ``/root/reference`` does
not exist.  Only ``tests/``, ``bench.py --impl reference`` / its ``cpu_baseline`` leg and the
harness ``oracle/ref_gpu.py`` import from it; nothing under ``sam_quantization_b200/`` does.

What is staged (copied byte for byte, no edits):
  gptq_triton/            the Triton QuantLinear / fused attention the product replaces
  segment_anything/       the host encoder the operators are swapped into (+ build_sam configs)
  gptq.py                 Quantizer / GPTQ solver (fixtures, f-2)

    python oracle/make_ref.py            # in the build container (needs /root/reference)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SAMQ_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
ITEMS = ("gptq_triton", "segment_anything", "gptq.py")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DST, "gptq_triton"))


def make(verbose: bool = True) -> bool:
    """Copy the reference's hot-path sources into oracle/_ref/.  Returns False (and does
    nothing) when /root/reference is absent -- e.g. on the GPU box, which uses the staged copy."""
    if not os.path.isdir(REF_SRC):
        return False
    os.makedirs(REF_DST, exist_ok=True)
    for item in ITEMS:
        src, dst = os.path.join(REF_SRC, item), os.path.join(REF_DST, item)
        if os.path.isdir(src):
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(src, dst)
        if verbose:
            print(f"staged {src} -> {dst}")
    return True


def add_to_path() -> str:
    """Put oracle/_ref first on sys.path (harness / tests only) and return it."""
    if not available():
        raise ImportError("oracle/_ref is empty: run `python oracle/make_ref.py` in the build container")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    return REF_DST


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref staged" if ok else f"{REF_SRC} not found: nothing staged")
