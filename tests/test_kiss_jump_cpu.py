"""The O(1) jump-ahead of the McICA KISS generator (csrc/kiss_jump.cpp builds the jump entries on the host,
csrc/mcica.cuh applies them on the device) against the sequential generator of SH/cloud_subcol_gen.F90:546-607.

The table builder is a plain host function of the shipped library and is called here as it is; the few lines that
apply an entry on the device (Kiss::jump, mwc_jump) are mirrored in Python integers.  Covers the fixed points and the
transient of the multiply-with-carry lanes (0, m, m+1) that the closed form has to treat apart."""
import ctypes as C

import numpy as np
import pytest

M32 = 0xffffffff


class KissJump(C.Structure):       # csrc/common.cuh
    _fields_ = [("lcg_a", C.c_uint32), ("lcg_c", C.c_uint32), ("mwc3", C.c_uint32), ("mwc4", C.c_uint32),
                ("n", C.c_uint32), ("pad", C.c_uint32 * 3), ("xs", C.c_uint32 * 32)]


@pytest.fixture(scope="module")
def jump_table():
    from geosradiation_gridcomp_b200 import host
    fn = getattr(host.lib(), "_ZN6rrtmgx15kiss_jump_tableEiibPNS_8KissJumpE")
    fn.restype = None
    fn.argtypes = [C.c_int, C.c_int, C.c_bool, C.POINTER(KissJump)]

    def build(nsub, nlay, inhomo):
        out = (KissJump * (2 * nsub))()
        fn(nsub, nlay, inhomo, out)
        return out
    return build


def step(s):
    s1, s2, s3, s4 = s
    s1 = (69069 * s1 + 1327217885) & M32
    s2 ^= (s2 << 13) & M32
    s2 ^= s2 >> 17
    s2 ^= (s2 << 5) & M32
    s3 = (18000 * (s3 & 65535) + (s3 >> 16)) & M32
    s4 = (30903 * (s4 & 65535) + (s4 >> 16)) & M32
    return s1, s2, s3, s4


def mwc_jump(y, n, mult, a):       # csrc/mcica.cuh mwc_jump<A>
    m = a * 65536 - 1
    mw = lambda v: (a * (v & 65535) + (v >> 16)) & M32
    if n == 0: return y
    y = mw(y)
    if n == 1: return y
    y = mw(y)
    if n == 2: return y
    if y == 0 or y == m: return y
    if y == m + 1:
        y = mw(y)
        mult = (mult * 65536) % m
    return (mult * y) % m


def apply_jump(J, s):              # csrc/mcica.cuh Kiss::jump
    s1, s2, s3, s4 = s
    s1 = (J.lcg_a * s1 + J.lcg_c) & M32
    r = 0
    for j in range(32):
        if (s2 >> j) & 1:
            r ^= J.xs[j]
    return s1, r, mwc_jump(s3, J.n, J.mwc3, 18000), mwc_jump(s4, J.n, J.mwc4, 30903)


@pytest.mark.parametrize("nlay,inhomo", [(72, True), (72, False), (181, True), (1, False)])
def test_jump_entries_reproduce_the_sequential_generator(jump_table, nlay, inhomo):
    nsub = 140
    tabl = jump_table(nsub, nlay, inhomo)
    stride = (4 if inhomo else 2) * nlay
    rng = np.random.default_rng(11)
    m3, m4 = 18000 * 65536 - 1, 30903 * 65536 - 1
    seeds = [tuple(int(v) for v in rng.integers(0, 1 << 32, 4)) for _ in range(6)]
    # states whose multiply-with-carry lanes sit on or next to the fixed points of y -> a*y mod m
    seeds += [(1, 1, 0, 0), (5, 7, m3, m4), (5, 7, m3 + 1, m4 + 1), (M32, M32, M32, M32), (0, 1, 65536, 65535),
              (3, 9, (m3 + 1) & M32, 1)]
    for s0 in seeds:
        s, k = s0, 0
        for i in range(nsub):
            for which, n in ((0, i * stride), (1, i * stride + 2 * nlay)):
                while k < n:
                    s = step(s); k += 1
                J = tabl[2 * i + which]
                assert J.n == n
                assert apply_jump(J, s0) == s, (s0, n)


def test_sequential_step_is_the_reference_generator():
    """The `step` above against the int32 formulation the oracle tests pin (tests/test_oracle_cpu.py kiss_python)."""
    from test_oracle_cpu import kiss_python
    for seeds in [(123456789, 362436069, 521288629, 916191069), (-5, 77, -2147483648, 2147483647)]:
        want = kiss_python(seeds, 50)
        s = tuple(v & M32 for v in seeds)
        got = []
        for _ in range(50):
            s = step(s)
            k = (s[0] + s[1] + ((s[2] << 16) & M32) + s[3]) & M32
            k = k - (1 << 32) if k >= (1 << 31) else k
            got.append(k * 2.328306e-10 + 0.5)
        np.testing.assert_array_equal(np.array(got), want)
