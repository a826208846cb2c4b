"""Host-side check of a device arithmetic identity (CPU, exact rational arithmetic): the terminal-equality sweep kernel
(cddp-cpp_b200/csrc/ipddp_teq_small.cuh, `div_shared`) forms the quotients by one LDL^T pivot from ONE correctly rounded
reciprocal r = RN(1/d):   q0 = RN(x r),  e = x - d q0 (exact in an FMA),  q = RN(q0 + e r)   (Markstein).
The reference divides (Eigen's `dst.row(i) /= vecD(i)`), so the substitution is only legitimate if q is the correctly rounded
quotient RN(x/d): checked here bit for bit on random operands over sixteen decades, and bounded by one ulp on the known
hard case (divisors whose significand is all ones, where RN(1/d) is least accurate)."""
import math
import random
from fractions import Fraction as F


def fma(a, b, c):
    return float(F(a) * F(b) + F(c))  # float(Fraction) rounds to nearest even: an exact fused multiply-add


def div_shared(x, d):
    r = 1.0 / d
    q = x * r
    return fma(fma(-d, q, x), r, q)


def test_random_operands_give_the_correctly_rounded_quotient():
    rng = random.Random(20261017)
    bad = 0
    for _ in range(60000):
        x = math.copysign(10.0 ** rng.uniform(-8, 8) * rng.uniform(1.0, 10.0), rng.choice((-1.0, 1.0)))
        d = math.copysign(10.0 ** rng.uniform(-8, 8) * rng.uniform(1.0, 10.0), rng.choice((-1.0, 1.0)))
        bad += div_shared(x, d) != x / d
    assert bad == 0, f"{bad} of 60000 quotients differ from IEEE division"


def test_all_ones_significands_stay_within_one_ulp():
    rng = random.Random(7)
    worst = 0.0
    for k in range(2000):
        d = math.ldexp(2.0 - math.ldexp(1.0, -52 + (k % 3)), rng.randrange(-20, 20))  # 1.11...1, 1.11...10, 1.11...100
        x = rng.uniform(-1e3, 1e3)
        q, ref = div_shared(x, d), x / d
        worst = max(worst, abs(q - ref) / math.ulp(ref))
    assert worst <= 1.0, worst


def test_zero_numerator_and_signs():
    for d in (3.0, -7.5, 1e-200, 1e200):
        assert div_shared(0.0, d) == 0.0
        assert div_shared(d, d) == 1.0
        assert div_shared(-d, d) == -1.0
