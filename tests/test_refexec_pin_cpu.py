"""THE PIN of the oracle: the C restatement (oracle/*.c) against numbers produced by the REFERENCE'S OWN SOURCE TEXT.

tests/golden/rrtmg_refexec_golden.npz was written by tests/golden/make_golden_from_refexec.py, which executes the
Fortran of /root/reference (RRTMG LW + SW, McICA, NRLSSI2: 91 files, read where they lie) through the translator
oracle/refexec/f90py.py - no Fortran compiler exists in this image, so this is how the reference runs here.  The
cases (tests/golden/refexec_cases.py) cover the GEOS defaults at L72 and L181, every LW and SW cloud-optics option,
every solar-variability mode, the three condensate inhomogeneity options, non-default decorrelation lengths, ragged
partitions, and every intermediate the oracle taps (jp / jt / jt1 / indices, interpolation factors, the McICA
sub-column mask, optical depths, Planck fractions, Rayleigh, solar source).

Bar: integers (indices, masks, clear counts) bit for bit; reals within 1e-12 relative (what is seen is bit-identical
or last-place: gcc's libm `exp` / `log` / `pow` against CPython's, and numpy's pairwise `sum`)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refexec_cases as rc   # noqa: E402

TOL = 1e-12


@pytest.fixture(scope="module")
def golden():
    assert os.path.exists(rc.GOLDEN), "tests/golden/rrtmg_refexec_golden.npz is part of the repository"
    g = np.load(rc.GOLDEN)
    assert int(g["real_bytes"]) == 8
    return g


def test_golden_file_holds_every_case(golden):
    keys = set(golden.files)
    for name, c in rc.CASES.items():
        if c["lw"] is not None:
            for k in rc.LW_OUT + (rc.LW_TAPS if c["taps"] else ()):
                assert f"{name}/lw/{k}" in keys, (name, k)
        if c["sw"] is not None:
            for k in rc.SW_OUT + (tuple(rc.SW_TAPS) if c["taps"] else ()):
                assert f"{name}/sw/{k}" in keys, (name, k)
    # the cases are not trivially cloud-free: cloudy sub-columns exist, and the clear-sky stream differs from the total
    assert (golden["default/lw/clearCounts"][:, 0] < 140).sum() >= 10
    assert np.abs(golden["default/lw/uflx"] - golden["default/lw/uflxc"]).max() > 10.0
    assert np.abs(golden["default/sw/swdflx"] - golden["default/sw/swdflxc"]).max() > 0.1
    assert golden["taps/lw/cldymc"].sum() > 1000 and golden["taps/sw/cldymc"].sum() > 1000


@pytest.mark.parametrize("name", list(rc.CASES))
def test_oracle_equals_reference_source_output(oracle, golden, name):
    got = rc.run_case(oracle, name, lambda ih, corr: oracle.set_mcica(ih, corr), lambda: oracle.set_mcica(1))
    n, same, worst = rc.check_case(got, golden, name, TOL)
    assert n >= 9
    print(f"{name}: {n} arrays, {same} bit-identical, worst relative difference {worst:.2e}")


def test_oracle_kiss_equals_reference_source_draws(oracle, golden):
    """rng_kiss (SH/cloud_subcol_gen.F90:546-576) executed from the reference text, 4 seed quadruples x 600 draws
    (with the extreme seeds of the 32-bit range): the C restatement draws the same numbers, bit for bit."""
    seeds, ran = golden["kiss/seeds"], golden["kiss/ran_num"]
    assert ran.shape == (4, 600) and 0.0 < ran.min() and ran.max() < 1.0
    for i in range(seeds.shape[0]):
        r, _ = oracle.rng_kiss(seeds[i].astype(np.int32), ran.shape[1])
        np.testing.assert_array_equal(np.asarray(r), ran[i], err_msg=str(seeds[i]))


@pytest.mark.skipif(not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")), reason="no reference tree on this machine")
def test_golden_is_reproducible_from_the_reference_tree(golden):
    """One small case regenerated on the spot from /root/reference: the committed file is what the reference's text
    yields today (runs in the build container only; about a minute)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_from_refexec as gen
    name = "lw_no_dudts"
    fresh = gen.run_case(name, rc.CASES[name])
    for k, v in fresh.items():
        np.testing.assert_array_equal(np.asarray(v, dtype=golden[k].dtype), golden[k], err_msg=k)
