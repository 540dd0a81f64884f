"""div_rcp (epoch_b200/csrc/epb_internal.h): a / b as a * rb followed by one residual correction through two
FMAs, rb = 1.0 / b.  The moment and collision kernels use it for quotients by loop constants, because the
compiler's IEEE division ends in a slow-path branch that keeps neighbouring divisions from overlapping.  This is
the arithmetic claim the kernels' comments make, checked with exact rational arithmetic: the result is the
correctly rounded quotient except in rare cases, and never more than one unit in the last place away."""
from fractions import Fraction

import numpy as np


def _fma(a, b, c):
    # float(Fraction) rounds to nearest even, so this is the exactly rounded fused multiply-add
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def div_rcp(a, b, rb):
    q = a * rb
    return _fma(_fma(-q, b, a), rb, q)


def test_div_rcp_is_the_ieee_quotient_up_to_rare_last_place_cases():
    rng = np.random.default_rng(11)
    c, m0 = 2.99792458e8, 9.10938291e-31
    divisors = [c, m0, m0 * c, np.sqrt(m0), 1836.2 * m0, 4.0e-8, 1.0 / 3.0e6] + list(rng.uniform(1.0, 2.0, 8)) + \
        list(10.0 ** rng.uniform(-35, 12, 8))
    n_wrong = n = 0
    for b in divisors:
        b = float(b)
        rb = 1.0 / b
        for a in rng.standard_normal(800) * 10.0 ** rng.uniform(-30, 3, 800):
            a = float(a)
            want = a / b
            got = div_rcp(a, b, rb)
            n += 1
            if got != want:
                n_wrong += 1
                assert abs(got - want) <= np.spacing(abs(want)), (a, b, got, want)
    assert n_wrong <= n // 1000, (n_wrong, n)


def test_div_rcp_exact_cases():
    for a, b in ((6.0, 3.0), (1.0, 4.0), (-7.5, 2.5), (0.0, 3.0), (1.0e-300, 1.0e-10)):
        assert div_rcp(a, b, 1.0 / b) == a / b
