"""Known-answer tests of the 256-bit field arithmetic against Python ints.

The reference's arithmetic is AbstractAlgebra GF(p) over BigInt (un-vendored, no KATs of its own):
Python's arbitrary-precision ints are the independent ground truth for both the oracle's
shift-subtract code (CPU) and the device Montgomery code (GPU, through ecne_fr_batch)."""
import ctypes as C
import random

import numpy as np
import pytest

from ecneproject_b200 import _abi
from helpers import P, from_limbs, to_limbs
import oracle_lib

R = 1 << 256
EDGE = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, R % P, (R * R) % P, 2 ** 253, 2 ** 253 - 1,
        2 ** 128, 2 ** 64, 2 ** 64 - 1, 2 ** 192 + 12345, 0x43e1f593f0000000, P - 2 ** 64]


def vectors(n, seed):
    rng = random.Random(seed)
    a = EDGE + [rng.randrange(P) for _ in range(n)]
    b = list(reversed(EDGE)) + [rng.randrange(P) for _ in range(n)]
    # pair every edge value with every other edge value as well
    for x in EDGE:
        for y in EDGE:
            a.append(x)
            b.append(y)
    return a, b


def expected(op, a, b):
    if op == 0:
        return [(x + y) % P for x, y in zip(a, b)]
    if op == 1:
        return [(x - y) % P for x, y in zip(a, b)]
    if op == 2:
        return [(x * y) % P for x, y in zip(a, b)]
    if op == 3:
        return [pow(x, -1, P) if x else 0 for x in a]
    if op == 4:
        return [(-x) % P for x in a]
    if op == 5:
        return [((-x) * pow(y, -1, P)) % P if y else 0 for x, y in zip(a, b)]


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5])
def test_oracle_field_ops(op):
    a, b = vectors(2000, 1234 + op)
    A, B = to_limbs(a), to_limbs(b)
    out = np.zeros_like(A)
    oracle_lib.lib().ecne_oracle_fr(op, len(a), A.ctypes.data_as(_abi.u64p), B.ctypes.data_as(_abi.u64p),
                                    out.ctypes.data_as(_abi.u64p))
    assert from_limbs(out) == expected(op, a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5])
def test_device_field_ops(op):
    a, b = vectors(10000, 99 + op)
    if op == 5:  # the device divexact is only called with a non-zero divisor (the caller raises)
        b = [y if y else 1 for y in b]
    A, B = to_limbs(a), to_limbs(b)
    out = np.zeros_like(A)
    lib = _abi.engine_lib()
    st = lib.ecne_fr_batch(op, len(a), A.ctypes.data_as(_abi.u64p), B.ctypes.data_as(_abi.u64p),
                           out.ctypes.data_as(_abi.u64p))
    assert st == 0, lib.ecne_last_error()
    assert from_limbs(out) == expected(op, a, b)


@pytest.mark.gpu
def test_device_integer_helpers_of_case5():
    """The plain 256-bit integer divisibility / product-compare used by the mixed-radix rule (:1266-1274)."""
    import random
    rng = random.Random(7)
    a, b = [], []
    for i in range(6000):
        kind = i % 6
        if kind == 0:      # exact multiples with nearby lengths (what a sorted coefficient chain produces)
            y = rng.getrandbits(rng.randint(1, 200)) | 1
            x = y * rng.randint(1, 1 << rng.randint(0, 50))
        elif kind == 1:    # near misses
            y = rng.getrandbits(rng.randint(2, 200)) | 2
            x = y * rng.randint(1, 1 << 40) + rng.randint(1, y - 1)
        elif kind == 2:    # powers of two
            y = 1 << rng.randint(0, 250)
            x = rng.getrandbits(253) & ~((1 << rng.randint(0, 252)) - 1)
        elif kind == 3:    # random field-size values
            y = rng.randrange(1, P)
            x = rng.randrange(0, P)
        elif kind == 4:    # divisor longer than the dividend, zero dividend
            y = rng.getrandbits(200) | (1 << 199)
            x = rng.choice([0, rng.getrandbits(100)])
        else:              # equal values, one
            y = rng.choice([1, rng.randrange(1, P)])
            x = y if rng.random() < 0.5 else rng.randrange(0, P)
        a.append(x % (1 << 256))
        b.append(y)
    A, B = to_limbs(a), to_limbs(b)
    lib = _abi.engine_lib()
    out = np.zeros_like(A)
    assert lib.ecne_fr_batch(6, len(a), A.ctypes.data_as(_abi.u64p), B.ctypes.data_as(_abi.u64p),
                             out.ctypes.data_as(_abi.u64p)) == 0, lib.ecne_last_error()
    assert from_limbs(out) == [1 if x % y == 0 else 0 for x, y in zip(a, b)]
    out = np.zeros_like(A)
    assert lib.ecne_fr_batch(7, len(a), A.ctypes.data_as(_abi.u64p), B.ctypes.data_as(_abi.u64p),
                             out.ctypes.data_as(_abi.u64p)) == 0, lib.ecne_last_error()
    assert from_limbs(out) == [(x * y > P) - (x * y < P) + 1 for x, y in zip(a, b)]
