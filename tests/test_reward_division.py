"""The reward's life / 100.0 (zombsole/gym/reward.py:37-41) is evaluated on the device without a division:
q0 = RN(a * RN(1/100)); q = RN(q0 + RN(a - 100 * q0) * RN(1/100)) with fused multiply-adds.  This test proves, in
exact rational arithmetic, that the result equals the correctly rounded quotient (what Python's a / 100.0 gives)
for every integer the kernel sends down that path (|a| <= 200,000; larger values take the real division)."""
from fractions import Fraction


def rn(x):
    """nearest double of an exact rational (int / int true division is correctly rounded in CPython)"""
    return 0.0 if x == 0 else x.numerator / x.denominator


def test_newton_step_division_by_100_is_exact():
    r = Fraction(0.01)
    for a in range(-200000, 200001):
        q0 = rn(a * r)
        e = rn(a - 100 * Fraction(q0))   # fma(-100, q0, a)
        q = rn(Fraction(e) * r + Fraction(q0))  # fma(e, r, q0)
        assert q == a / 100.0, a
