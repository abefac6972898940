"""The exactness claim behind ``div_magic`` (csrc/rectify_common.cuh): for every 32-bit ``n`` and every
divisor ``d >= 2``, ``(n * ceil(2**64 / d)) >> 64 == n // d`` -- K1 uses it for quad-row, tile-column and
tile-row indices instead of generic 64-bit divisions.  Checked here with Python integers (no GPU): the
edge values of ``n`` around every multiple boundary for a sweep of divisors, and random pairs."""

import random


def _magic(d: int) -> int:
    return 0 if d <= 1 else ((2**64 - 1) // d) + 1  # div_magic_of: floor((2^64 - 1) / d) + 1 == ceil(2^64 / d), d >= 2


def _div(n: int, m: int) -> int:
    return ((n * m) >> 64) if m else n  # div_magic: __umul64hi(n, magic), or n itself for d == 1


def test_magic_is_the_ceiling_of_two_to_the_64_over_d():
    for d in list(range(2, 5000)) + [2**k for k in range(1, 32)] + [2**k - 1 for k in range(2, 33)] + [2**31 + 1, 2**32 - 1]:
        assert _magic(d) == -(-(2**64) // d), d
    assert _magic(1) == 0 and _div(123456, _magic(1)) == 123456


def test_quotients_are_exact_at_every_multiple_boundary():
    top = 2**32 - 1
    divisors = list(range(2, 300)) + [511, 512, 513, 1000, 4090, 4091, 4864, 4865, 7992, 10980, 36000, 65535, 65536, 65537,
                                      2**20 - 1, 2**20 + 1, 2**31 - 1, 2**31, 2**31 + 1, top]
    for d in divisors:
        m = _magic(d)
        ks = {0, 1, 2, top // d, top // d - 1, (top // d) // 2}
        for k in ks:
            for n in (k * d - 1, k * d, k * d + 1, k * d + d - 1):
                if 0 <= n <= top:
                    assert _div(n, m) == n // d, (n, d)
        assert _div(top, m) == top // d


def test_random_pairs():
    rng = random.Random(7)
    for _ in range(200000):
        d = rng.randrange(2, 2**32)
        n = rng.randrange(0, 2**32)
        assert _div(n, _magic(d)) == n // d
    for _ in range(100000):  # small divisors: many multiples inside the 32-bit range
        d = rng.randrange(2, 70000)
        n = rng.randrange(0, 2**32)
        assert _div(n, _magic(d)) == n // d
