"""pmath.h (the portable libm shared by the CUDA kernels and the oracle's pmath mode): error against mpmath and
agreement with glibc.  Runs on CPU through the oracle library's probe entry."""
import math

import mpmath as mp
import numpy as np
import pytest

SIN, COS, TAN, ACOS, LOG, ATAN2, EXPF = range(7)


def _ulp(x):
    x = abs(float(x))
    if x == 0.0:
        return 5e-324
    return math.ldexp(1.0, math.frexp(x)[1] - 53)


def _max_ulp_err(fn_id, mpf, xs, orc, ys=None):
    pm, libm = orc.pmath_eval(fn_id, xs, ys)
    mp.mp.dps = 40
    worst_pm = worst_gl = 0.0
    for k in range(xs.size):
        t = mpf(mp.mpf(float(xs[k]))) if ys is None else mpf(mp.mpf(float(ys[k])), mp.mpf(float(xs[k])))
        u = _ulp(float(t))
        worst_pm = max(worst_pm, float(abs(mp.mpf(float(pm[k])) - t)) / u)
        worst_gl = max(worst_gl, float(abs(mp.mpf(float(libm[k])) - t)) / u)
    agree = float(np.mean(pm == libm))
    return worst_pm, worst_gl, agree


@pytest.mark.parametrize("fn,mpf,lo,hi,bound", [
    (SIN, mp.sin, -1.7, 7.0, 1.0), (COS, mp.cos, -1.7, 7.0, 1.0), (TAN, mp.tan, 0.0, 1.5707, 2.5),
    (ACOS, mp.acos, -1.0, 1.0, 1.0), (ACOS, mp.acos, 1 - 1e-7, 1.0, 0.6), (LOG, mp.log, 1e-9, 60.0, 1.0),
])
def test_pmath_accuracy(orc, fn, mpf, lo, hi, bound):
    rng = np.random.default_rng(fn * 17 + 3)
    xs = rng.uniform(lo, hi, 3000)
    e_pm, e_gl, agree = _max_ulp_err(fn, mpf, xs, orc)
    assert e_pm < bound, f"pmath fn {fn}: max error {e_pm:.3f} ulp (glibc {e_gl:.3f})"
    assert agree > (0.5 if fn == TAN else 0.9), f"agreement with glibc only {agree:.3f}"


def test_pmath_atan2_accuracy(orc):
    rng = np.random.default_rng(99)
    xs, ys = rng.uniform(-3.2, 3.2, 3000), rng.uniform(-3.2, 3.2, 3000)
    e_pm, e_gl, agree = _max_ulp_err(ATAN2, lambda y, x: mp.atan2(y, x), xs, orc, ys)
    assert e_pm < 1.6 and agree > 0.75


def test_pmath_expf_matches_float_rounding(orc):
    xs = np.linspace(-30.0, 0.0, 4001).astype(np.float32).astype(np.float64)
    pm, libm = orc.pmath_eval(EXPF, xs)
    ref = np.exp(xs).astype(np.float32).astype(np.float64)
    assert np.max(np.abs(pm - ref) / np.maximum(ref, 1e-30)) < 1.3e-7      # within one float ulp
    assert np.mean(pm == libm) > 0.99


def test_pmath_special_values(orc):
    pi = math.pi
    x = np.array([0.0, pi / 2, pi, 90 * (pi / 180.0), -0.0])
    pm, libm = orc.pmath_eval(SIN, x)
    assert pm[0] == 0.0 and pm[1] == 1.0 and pm[3] == 1.0
    pm, libm = orc.pmath_eval(COS, x)
    assert pm[0] == 1.0 and pm[1] == libm[1]          # cos(pi/2 rounded) = 6.123e-17, same bits as glibc
    pm, _ = orc.pmath_eval(ACOS, np.array([1.0, -1.0, 0.0]))
    assert pm[0] == 0.0 and pm[1] == pi and pm[2] == pi / 2
    pm, _ = orc.pmath_eval(ATAN2, np.array([1.0, 0.0, -1.0, 0.0]), np.array([0.0, 1.0, 0.0, -1.0]))   # (x, y)
    assert pm[0] == 0.0 and pm[1] == pi / 2 and pm[2] == pi and pm[3] == -pi / 2
    pm, _ = orc.pmath_eval(LOG, np.array([1.0, math.e]))
    assert pm[0] == 0.0 and abs(pm[1] - 1.0) < 3e-16


def test_local_coord_squared_residual_bound():
    """geom.cuh decides LOCAL_COORD's "zres > zresmax" (src/mod_bilin_2d.f90:741, zres = SQRT(s), zresmax = 0.1 in REAL(4))
    on s: 0x1.47ae151eb8520p-7 must be the largest double whose correctly rounded square root is <= REAL(0.1,4)."""
    import math
    from fractions import Fraction as Fr
    import re
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sosie_b200", "csrc", "geom.cuh")).read()
    m = re.search(r"zres2max = (0x[0-9a-fp.+-]+);", src)
    assert m, "constant not found in geom.cuh"
    s0 = float.fromhex(m.group(1))
    c = float(np.float32(0.1))
    bound = (Fr(c) + Fr(math.ulp(c)) / 2) ** 2      # sqrt(s) rounds to <= c  <=>  s <= (c + ulp/2)^2 (tie to even: c is even)
    assert Fr(s0) <= bound < Fr(math.nextafter(s0, 1.0))
    y_dn, y_up = s0, math.nextafter(s0, 1.0)
    for _ in range(200):                             # math.sqrt is correctly rounded
        assert not (math.sqrt(y_dn) > c) and (math.sqrt(y_up) > c)
        y_dn, y_up = math.nextafter(y_dn, 0.0), math.nextafter(y_up, 1.0)
