"""The device KISS generator on its own (rrtmgx_debug_kiss): the draws against rng_kiss executed from the reference's
source text (tests/golden/rrtmg_refexec_golden.npz, keys kiss/*), the real*4 scaling against the range the reference
records (SH/cloud_subcol_gen.F90:597-604: the only expected value the path holds), and the O(1) jump-ahead the McICA
kernel uses against replaying the sequence draw by draw on the device."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refexec_cases as rc   # noqa: E402

pytestmark = pytest.mark.gpu


def test_device_draws_equal_the_reference_source_draws(rx):
    g = np.load(rc.GOLDEN)
    seeds, ran = g["kiss/seeds"].astype(np.int32), g["kiss/ran_num"]
    o = rx.debug_kiss(seeds, ndraw=ran.shape[1])
    np.testing.assert_array_equal(o["ran8"], ran)
    # the integer behind every draw, recovered exactly: ran_num is monotone in kiss and 2.328306e-10 * 2^32 < 1
    np.testing.assert_array_equal(o["kiss"].astype(np.float64) * 2.328306e-10 + 0.5, ran)
    # real*4 scaling of the same integers = IEEE single arithmetic (int -> real*4 rounds, then product, then sum)
    want4 = o["kiss"].astype(np.float32) * np.float32(2.328306e-10) + np.float32(0.5)
    np.testing.assert_array_equal(o["ran4"], want4)


def test_real4_scaling_has_the_range_the_reference_records(rx):
    """`mini * 2.328306e-10 + 0.5` and `maxi * 2.328306e-10 + 0.5` in real*4: 8.9406967E-08 and 0.9999999, the output
    of the reference's rng_test.f90 under ifort (SH/cloud_subcol_gen.F90:583-604)."""
    o = rx.debug_kiss([[1, 2, 3, 4]], values=[-2 ** 31, 2 ** 31 - 1, 0])
    assert "%.7E" % o["val4"][0] == "8.9406967E-08"
    assert "%.7f" % o["val4"][1] == "0.9999999"
    assert o["val4"][2] == np.float32(0.5)
    assert 0.0 < o["val8"][0] < 1e-7 and 0.9999999 < o["val8"][1] < 1.0   # real*8: strictly inside (0, 1) too


@pytest.mark.parametrize("nsub,nlay,inhomo", [(140, 72, 1), (112, 72, 1), (140, 181, 1), (112, 72, 0)])
def test_device_jump_ahead_equals_replaying_the_draws(rx, nsub, nlay, inhomo):
    rng = np.random.default_rng(nsub + nlay)
    seeds = np.concatenate([rng.integers(-2 ** 31, 2 ** 31, size=(29, 4)),
                            [[0, 1, 0, 0], [-1, -1, -1, -1], [7, 123, 65535, 65536]]]).astype(np.int32)
    o = rx.debug_kiss(seeds, jump_table=(nsub, nlay, inhomo))
    assert o["jumped"].shape == (32, 2 * nsub, 4)
    np.testing.assert_array_equal(o["jumped"], o["replayed"])
    # entry 0 jumps nothing: the state is the seeds; later entries move it
    np.testing.assert_array_equal(o["jumped"][:, 0, :], seeds.view(np.uint32))
    assert (o["jumped"][:, 2, 0] != o["jumped"][:, 0, 0]).all()
