"""The sampler's quotient RN(p / acc) is computed as q0 = RN(p y), r = RN(p - acc q0), q = RN(q0 + r y) with
y = RN(1 / acc) (common_b200/csrc/msb_kernels.cuh, dart_walk).  Checked here against exact rational arithmetic
(util.hpp:125-156 divides with the IEEE operator, so the sequence must round identically)."""
import random
import struct
from fractions import Fraction

import numpy as np


def rn32(fr):
    """nearest binary32 (ties to even, gradual underflow) of a Fraction"""
    if fr == 0:
        return np.float32(0.0)
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    if Fraction(2) ** (e + 1) <= a:
        e += 1
    e = max(e, -126)
    ulp = Fraction(2) ** (e - 23)
    n = a / ulp
    fl = n.numerator // n.denominator
    rem = n - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return np.float32(sign * float(Fraction(fl) * ulp))


def F(x):
    return Fraction(float(x))


def fma(a, b, c):
    return rn32(F(a) * F(b) + F(c))


def markstein(p, acc):
    y = rn32(1 / F(acc))
    q0 = rn32(F(p) * F(y))
    r = fma(-acc, q0, p)
    return fma(r, y, q0)


def bits(e, m):
    return np.frombuffer(struct.pack("<I", (e << 23) | m), np.float32)[0]


def test_markstein_sequence_rounds_like_the_ieee_division():
    rng = random.Random(7)
    edge = [0, 1, 2, 0x7FFFFF, 0x7FFFFE, 0x400000, 0x3FFFFF]
    checked = 0
    for i in range(20000):
        mode = i % 4
        if mode == 0:
            p, acc = np.float32(rng.random()), np.float32(1 + rng.random() * 300)
        elif mode == 1:
            p, acc = np.float32(2.0 ** rng.uniform(-100, 0)), np.float32(2.0 ** rng.uniform(0, 24))
        elif mode == 2:  # mantissa extremes
            p = bits(rng.randint(27, 126), rng.choice(edge + [rng.getrandbits(23)]))
            acc = bits(rng.randint(127, 150), rng.choice(edge + [rng.getrandbits(23)]))
        else:            # acc as a rounded sum of probabilities, p one of them
            terms = [np.float32(rng.random() ** 8) for _ in range(10)]
            acc = np.float32(1 + float(sum(terms)))
            p = terms[0]
        if not (p >= np.float32(2.0 ** -100)):
            continue
        assert markstein(p, acc) == rn32(F(p) / F(acc)), (p, acc)
        checked += 1
    assert checked > 15000
    assert markstein(np.float32(0.0), np.float32(3.5)) == 0.0
