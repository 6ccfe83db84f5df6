"""csrc/rng.cuh forms 2 * generate_canonical(first, second) - 1 of a polar attempt in three fp64 instructions
(canonical_2x_minus_1: two FMAs and a clamp) instead of the six of the literal restatement of libstdc++
(/usr/include/c++/13/bits/random.tcc:3346-3381 + 1826-1833). The claim is that the BITS are the same: second * 2^32 and
the scalings by 2^-64 and 2 are exact, so each FMA rounds exactly where an addition did. Checked here against exact
rational arithmetic (Fraction -> float is correctly rounded), clamp included."""
import random
from fractions import Fraction


def literal(first, second):
    prod = float(second) * 4294967296.0                      # exact
    s = float(Fraction(first) + Fraction(prod))              # __dadd_rn
    r = s * 5.42101086242752217e-20                          # 2^-64, exact
    if r >= 1.0:
        r = 0.99999999999999989                              # nextafter(1, 0)
    return float(Fraction(2.0 * r) - 1)                      # __dmul_rn (exact), __dsub_rn


def fused(first, second):
    s = float(Fraction(second) * 4294967296 + Fraction(first))            # fma
    x = float(Fraction(s) * Fraction(1.0842021724855044e-19) - 1)         # fma with 2^-63
    return 0.99999999999999978 if s >= 18446744073709551616.0 else x      # 1 - 2^-52


def test_constants():
    assert 1.0842021724855044e-19 == 2.0 ** -63 and 5.42101086242752217e-20 == 2.0 ** -64
    assert 0.99999999999999978 == 1.0 - 2.0 ** -52 and 0.99999999999999989 == 1.0 - 2.0 ** -53


def test_fused_canonical_has_the_same_bits():
    rnd = random.Random(20260002)
    cases = [(0, 0), (0xffffffff, 0xffffffff), (0xfffffc00, 0xffffffff), (0xfffffbff, 0xffffffff), (0xfffff800, 0xffffffff),
             (1, 0), (0, 1), (0xffffffff, 0), (0x80000000, 0x7fffffff), (0x400, 0xffffffff), (0x3ff, 0xffffffff)]
    cases += [(rnd.getrandbits(32), rnd.getrandbits(32)) for _ in range(30000)]
    cases += [(rnd.getrandbits(32), 0xffffffff) for _ in range(5000)]      # the neighbourhood of the clamp
    cases += [(rnd.getrandbits(32), rnd.getrandbits(3)) for _ in range(5000)]
    for first, second in cases:
        assert literal(first, second) == fused(first, second), (hex(first), hex(second))
