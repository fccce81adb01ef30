"""CPU: the one-`ex2` forward GELU of the SIMT kernels (uno_b200/csrc/tc_common.cuh gelu_fwd_fast) restated in numpy fp32 with the
coefficients parsed from the CUDA source, against the exact-erf GELU the reference computes (torch.nn.functional.gelu,
integral_operators.py:247 / darcy_flow_uno2d.py:105) in fp64.  Guards the polynomial against an accidental edit: the GPU parity tests
would only see a change of this size as a slightly larger error."""
import math
import os
import re

import numpy as np

from conftest import ROOT


def _coefficients():
    src = open(os.path.join(ROOT, "uno_b200", "csrc", "tc_common.cuh")).read()
    body = src[src.index("float gelu_fwd_fast(float x)"):]
    body = body[:body.index("return")]
    nums = [float(m) for m in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body)]
    assert len(nums) == 7, nums
    # source order: r = fmaf(z, c6, c5); then c4, c3, c2, c1, c0
    return nums[::-1]            # c0 .. c6


def test_gelu_fwd_fast_matches_the_erf_gelu():
    c = np.array(_coefficients(), np.float32)
    rng = np.random.default_rng(0)
    x = np.concatenate([np.linspace(-12, 12, 400001), rng.normal(size=200000) * 2]).astype(np.float32)
    a = np.abs(x)
    z = np.minimum(a, np.float32(6.0))
    r = np.full_like(z, c[6])
    for k in range(5, -1, -1):
        r = (r * z + c[k]).astype(np.float32)
    e = np.exp2((z * r - np.float32(1.0)).astype(np.float32)).astype(np.float32)
    got = (np.maximum(x, np.float32(0)) - a * e).astype(np.float32)
    x64 = x.astype(np.float64)
    ref = 0.5 * x64 * (1.0 + np.vectorize(math.erf)(x64 / math.sqrt(2.0)))
    err = np.abs(got - ref)
    assert err.max() < 5e-7, (float(err.max()), float(x[err.argmax()]))
    # relative to the magnitude of the input (the parity tolerance is 2e-5 of the tensor's largest value)
    assert (err / np.maximum(np.abs(x64), 1.0)).max() < 2e-7
    # the negative tail has no cancellation: gelu(-5) ~ -1.4e-6 is reproduced to a few percent, not flushed to zero
    i = np.argmin(np.abs(x + 5.0))
    assert abs(got[i] / ref[i] - 1.0) < 5e-2
