"""The GEMM epilogue's exact-erf GELU uses Abramowitz-Stegun 7.1.26 for erfc
(sam_quantization_b200/csrc/qlinear.cu::gelu_erf).  This re-evaluates the same formula in fp32
numpy and bounds its error against scipy's erf: it must be invisible after fp16 rounding."""
import math

import numpy as np
from scipy.special import erf


def gelu_as(x):
    x = x.astype(np.float32)
    ax = np.abs(x)
    t = (np.float32(1) / (ax * np.float32(0.3275911 * 0.70710678118654752440) + np.float32(1))).astype(np.float32)
    poly = np.float32(0.5 * 1.061405429)
    for c in (-1.453152027, 1.421413741, -0.284496736, 0.254829592):
        poly = (poly * t + np.float32(0.5 * c)).astype(np.float32)
    poly = (poly * t).astype(np.float32)
    # the epilogue's packed-fp32 form (qlinear_common.cuh::EpiBlock::finish):
    #   e = 2^((x x) (-log2(e)/2)),  g = poly e,  relu = 0.5 |x| + 0.5 x (exact),  gelu = fma(|x|, -g, relu)
    e = np.exp2(((x * x).astype(np.float32) * np.float32(-0.72134752044448170368)).astype(np.float32)).astype(np.float32)
    g = (poly * e).astype(np.float32)
    relu = (ax * np.float32(0.5) + x * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(relu, np.maximum(x, 0))
    return (relu.astype(np.float64) - ax.astype(np.float64) * g).astype(np.float32)


def test_gelu_formula_error_is_below_fp16_resolution():
    x = np.linspace(-12, 12, 400001).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / math.sqrt(2)))
    err = np.abs(gelu_as(x) - ref)
    assert err.max() < 5e-7
    # relative to one fp16 ulp of the result (2^-11 |y|, floor at the subnormal step 2^-24)
    ulp16 = np.maximum(np.abs(ref) * 2.0 ** -11, 2.0 ** -24)
    ratio = err / ulp16
    assert ratio.max() < 1.0                       # never more than one fp16 ulp (far negative tail, |y| < 2e-4)
    assert ratio[np.abs(ref) >= 1e-2].max() < 0.05  # invisible everywhere the output is not tiny
    assert abs(0.72134752044448170368 - math.log2(math.e) / 2) < 1e-15
