"""Proof by enumeration of the squared-distance thresholds used by the kernel (v0_step.cuh kSq*).

For s a double and d = sqrt(s) correctly rounded (numpy/libm sqrt is), the kernel replaces
    d <= 1.0 by s <= nextafter(1, inf);  d <= 2.0 by s <= nextafter(4, inf);  d > 12.0 by s > 144.0;
    d < r by s <= sqrt_less_than_bound(r).
sqrt is monotone, so it is enough to check the doubles around each bound.
"""
import math

import numpy as np


def neighbours(x, k=64):
    out, lo, hi = [x], x, x
    for _ in range(k):
        lo = math.nextafter(lo, -math.inf)
        hi = math.nextafter(hi, math.inf)
        out += [lo, hi]
    return sorted(out)


def test_fixed_thresholds():
    le1, le2 = float.fromhex("0x1.0000000000001p+0"), float.fromhex("0x1.0000000000001p+2")
    assert le1 == math.nextafter(1.0, math.inf) and le2 == math.nextafter(4.0, math.inf)
    for s in neighbours(1.0) + neighbours(le1):
        assert (math.sqrt(s) <= 1.0) == (s <= le1)
        assert (float(np.sqrt(np.float64(s))) <= 1.0) == (s <= le1)
    for s in neighbours(4.0) + neighbours(le2):
        assert (math.sqrt(s) <= 2.0) == (s <= le2)
        assert (math.sqrt(s) > 2.0) == (s > le2)
    for s in neighbours(144.0):
        assert (math.sqrt(s) > 12.0) == (s > 144.0)
    # strict bounds of the intercept rule (v0_step.cuh kSqLt1 / kSqLt4): d < 1 <=> s < 1, d < 4 <=> s < 16
    for s in neighbours(1.0):
        assert (math.sqrt(s) < 1.0) == (s < 1.0)
    for s in neighbours(16.0):
        assert (math.sqrt(s) < 4.0) == (s < 16.0)
    # v1 speed clamps (v1_step.cuh clamp_sq_*): sqrt(s) > max <=> s > largest s whose root is <= max
    for mx in (10.0, 25.0):
        from tests.test_v0_step_host import sqrt_less_than_bound
        bound = sqrt_less_than_bound(math.nextafter(mx, math.inf))
        for s in neighbours(bound) + neighbours(mx * mx):
            assert (math.sqrt(s) > mx) == (s > bound)


def test_reach_bound_for_many_speeds():
    from tests.test_v0_step_host import sqrt_less_than_bound
    rng = np.random.default_rng(0)
    for speed in [12.0, 9.5, 14.0, 1.0, 0.3, 100.0] + list(rng.uniform(0.1, 50.0, 200)):
        r = 0.1 * speed
        b = sqrt_less_than_bound(r)
        for s in neighbours(b, 16):
            if s >= 0:
                assert (math.sqrt(s) < r) == (s <= b)
    assert sqrt_less_than_bound(0.0) == -1.0


def test_constant_products():
    """v0_step.cuh writes these float64 products of reference constants as literals."""
    assert 68 * 0.2 == 13.600000000000001 and 68 * 0.8 == 54.400000000000006 and 105 * 0.1 == 10.5
    assert 105 * 0.6 == 63.0 and 105 * 0.75 == 78.75 and 68 * 0.5 == 34.0
    assert -50 * 0.3 == -15.0 and 60 * 0.3 == 18.0 and 30 * 0.3 == 9.0 and 0.9 / (1.0 - 2.0) == -0.9


def test_uniform_thresholds_as_integer_compares():
    """v0_step.cuh kU*: random() = k * 2**-24 (exact in float64); the reference's fixed thresholds as compares on k."""
    for k in list(range(15099494 - 70, 15099494 + 70)) + list(range(838860 - 70, 838860 + 70)) + list(range(13421773 - 70, 13421773 + 70)) + [0, 1, (1 << 24) - 1]:
        u = k * 2.0 ** -24
        assert (u < 0.9) == (k <= 15099494) and (u < 0.05) == (k <= 838860) and (u > 0.8) == (k >= 13421773)
