"""CPU: the closed-form approximations the kernels use, re-evaluated in numpy fp32 with the constants read from
csrc/common.cuh, so that a changed coefficient cannot slip through without a GPU:
  * ex2_poly2 -- 2^x for the softmax exponentials without the MUFU pipe (round-down magic add, degree-3 polynomial on
    the fraction, integer part added to the exponent field);
  * gelu_fast / gelu2x_fast2 -- erf-GELU as max(v, 0) - |v|/2 * 2^(-Q(|v|)) with a fitted cubic Q."""
import os
import re

import numpy as np

from tests.conftest import ROOT

SRC = open(os.path.join(ROOT, 'vit-ed_b200', 'csrc', 'common.cuh')).read()
f32 = np.float32


def _body(name):
    i = SRC.index(name + '(')
    j = SRC.index('\n}\n', i)
    return SRC[i:j]


def _floats(text):
    return [float(m) for m in re.findall(r'(-?\d+\.\d*(?:e-?\d+)?)f\b', text)]


def test_ex2_poly2_matches_exp2_to_a_sixth_of_an_fp16_step():
    body = _body('uint64_t ex2_poly2')
    c = _floats(body)
    # clamp, magic, -magic, c3, c2, c1, 1 -- each constant appears twice (both halves of the pair)
    assert c[:2] == [-126.0, -126.0] and 12582912.0 in c and -12582912.0 in c
    c3, c2, c1 = [f32(v) for v in (0.07706724107265472, 0.22764497995376587, 0.6951166391372681)]
    for v in (c3, c2, c1):
        assert float(v) in [float(f32(x)) for x in c], 'polynomial coefficient changed: update this test with the new fit'
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(200000) * 130, -rng.random(100000) * 8, rng.random(1000) * 8,      # (the lazy rescale
                        [0.0, -1.0, -126.0, -127.5, -0.5, -1e-7, 7.999]]).astype(f32)                   # lets x reach +8)
    xc = np.maximum(x, f32(-126))
    xf = np.floor(xc.astype(np.float64) + 12582912.0).astype(f32)          # add.rm: ulp is 1 at this magnitude
    fr = (xc - (xf - f32(12582912.0)).astype(f32)).astype(f32)
    assert fr.min() >= 0 and fr.max() < 1
    p = ((c3 * fr + c2).astype(f32) * fr + c1).astype(f32)
    p = (p * fr + f32(1)).astype(f32)
    assert p.min() >= 1 and p.max() < 2                                     # the mantissa never carries into the exponent
    bits = (p.view(np.uint32).astype(np.uint64) + ((xf.view(np.uint32).astype(np.uint64) << np.uint64(23)) & np.uint64(0xFFFFFFFF)))
    got = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(f32)
    ref = 2.0 ** np.maximum(x.astype(np.float64), -126)
    assert np.abs(got / ref - 1).max() < 9e-5                               # fp16 rounding of P is 4.9e-4
    assert got[x == 0][0] == 1.0


def test_gelu_cubic_stays_within_1e_4_of_erf_gelu():
    from scipy.special import erfc
    body = _body('float gelu_fast')
    c = _floats(body)
    c3, c2, c1 = f32(c[0]), f32(c[1]), f32(c[2])
    assert c3 < 0 and c2 < 0 and c1 < 0, '-Q(|v|) must be monotone decreasing: all three coefficients negative'
    assert [float(f32(v)) for v in _floats(_body('uint64_t gelu2x_fast2'))[:3]] == [float(c3), float(c2), float(c1)], \
        'gelu2x_fast2 (fused MLP) and gelu_fast (GEMM epilogue) must use the same fit'
    v = np.linspace(-12, 12, 400001).astype(f32)
    a = np.abs(v)
    q = ((a * c3 + c2).astype(f32) * a + c1).astype(f32) * a
    e = np.exp2(q.astype(np.float64))
    gelu = np.maximum(v, 0) - 0.5 * a * e
    exact = np.maximum(v.astype(np.float64), 0) - 0.5 * a.astype(np.float64) * erfc(a.astype(np.float64) / np.sqrt(2))
    assert np.abs(gelu - exact).max() < 1e-4
    # the fused MLP hands 2 * gelu to fc2 and halves the accumulator: v + |v| (1 - e) is exactly twice the value above
    twice = v.astype(np.float64) + a * (1 - e)
    assert np.abs(twice - 2 * gelu).max() < 1e-6
