"""The residue ("Ozaki scheme II") arithmetic of csrc/crt_gemm.cuh on the host: the library's own constants
(reach_crt_constants, no GPU needed) against exact integer arithmetic, and the kernel's reconstruction recipe replayed in
NumPy float64 on random integer dot products.  The GPU kernels are checked against exact products in tests/test_int8_gpu.py."""
import ctypes as C
import math
from fractions import Fraction

import numpy as np
import pytest

import reach_b200  # noqa: F401
from reach_b200 import _lib


def constants(N):
    lib = _lib.load()
    p = (C.c_int32 * N)()
    f1 = (C.c_double * N)()
    f2 = (C.c_double * N)()
    P, a, b = C.c_double(), C.c_int32(), C.c_int32()
    assert lib.reach_crt_constants(N, p, f1, f2, C.byref(P), C.byref(a), C.byref(b)) == 0
    return list(p), list(f1), list(f2), P.value, a.value, b.value


@pytest.mark.parametrize("N", range(10, 17))
def test_constants_match_exact_integer_arithmetic(N):
    p, f1, f2, Pd, alpha, beta = constants(N)
    assert all(2 <= q <= 256 for q in p) and len(set(p)) == N
    assert all(math.gcd(p[i], p[j]) == 1 for i in range(N) for j in range(i))       # pairwise coprime
    P = math.prod(p)
    assert Pd == float(P)
    # |sum x'_k y'_k| <= ||x'|| ||y'|| <= 2^(alpha+beta) must stay below P/4 (the reconstruction drops the integer part of
    # sum u_i f_i with ONE rounding and needs |C'/P| < 1/4 + 2^-28)
    assert 2 ** (alpha + beta) <= P // 4 and 2 ** (alpha + beta + 1) > P // 4 and 0 <= alpha - beta <= 1
    for i in range(N):
        Mi = P // p[i]
        w = Mi * pow(Mi, -1, p[i])                                                    # CRT weight, < P
        assert w % p[i] == 1 and all(w % p[j] == 0 for j in range(N) if j != i)
        # f1 = the first 40 bits of w / P, exactly; f2 = the next 80 bits rounded to FP64
        assert Fraction(f1[i]) == Fraction((w << 40) // P, 1 << 40)
        rest = Fraction(w, P) - Fraction(f1[i])
        assert 0 <= rest < Fraction(1, 1 << 40)
        assert abs(Fraction(f2[i]) - rest) <= Fraction(1, 1 << 92)
    # exactness of sum u_i f1_i in FP64: N residues below 256, multiples of 2^-40 -> below 2^12 with 52 bits
    assert N * 255 < 2 ** 12


def test_invalid_modulus_counts_are_refused():
    lib = _lib.load()
    assert lib.reach_crt_constants(9, None, None, None, None, None, None) != 0
    assert lib.reach_crt_constants(17, None, None, None, None, None, None) != 0


@pytest.mark.parametrize("N", [10, 13, 14, 16])
def test_reconstruction_recipe_recovers_integer_dot_products(N):
    """What the reconstruction warps do, in NumPy float64: u_i = C' mod p_i in [0, p_i) (any INT32 accumulator congruent to
    C' gives the same u_i), s1 = sum u_i f1_i (exact), s2 = sum u_i f2_i, C'/P = (s1 - rint(s1)) + s2."""
    p, f1, f2, Pd, alpha, beta = constants(N)
    rng = np.random.default_rng(N)
    bound = 2 ** (alpha + beta)
    worst = 0.0
    for trial in range(2000):
        shift = int(rng.integers(0, alpha + beta))
        Cp = int(rng.integers(-2 ** 62, 2 ** 62)) * int(rng.integers(1, 2 ** 62)) % (bound >> shift)
        if rng.integers(0, 2):
            Cp = -Cp
        u = np.array([Cp % q for q in p], dtype=np.float64)                          # Python's % is non-negative
        s1 = 0.0
        s2 = 0.0
        for i in range(N):
            prod = u[i] * f1[i]
            assert Fraction(prod) == Fraction(int(u[i])) * Fraction(f1[i])           # exact product
            nxt = s1 + prod
            assert Fraction(nxt) == Fraction(s1) + Fraction(prod)                     # exact sum
            s1 = nxt
            s2 += u[i] * f2[i]
        t = (s1 - np.rint(s1)) + s2
        got = t * Pd
        err = abs(Fraction(got) - Cp) / bound
        worst = max(worst, float(err))
        if abs(Cp) > 2 ** 60:                                                        # relative accuracy of the value itself
            assert abs(Fraction(got) - Cp) <= abs(Cp) * Fraction(1, 2 ** 51) + bound * Fraction(1, 2 ** 70)
    assert worst <= 2.0 ** -52


@pytest.mark.parametrize("p,c", [(256, 12345678), (255, -987654321), (253, 2 ** 30 - 1), (193, -(2 ** 30) + 1), (251, 0), (211, -1)])
def test_residue_reduction_of_the_accumulators(p, c):
    """crt_mod: cu = c + off (off a multiple of p, >= 2^30), q = umulhi(cu, floor(2^32 / p)), r = cu - q p in [0, 2p), one
    conditional subtraction."""
    off = ((1 << 30) // p + 1) * p
    magic = (1 << 32) // p
    cu = (c + off) & 0xFFFFFFFF
    q = (cu * magic) >> 32
    r = cu - q * p
    assert 0 <= r < 2 * p
    r = min(r, (r - p) & 0xFFFFFFFF)
    assert r == c % p
