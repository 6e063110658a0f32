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


@pytest.mark.skipif(not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")), reason="no reference tree on this machine")
@pytest.mark.parametrize("seed", [101, 202])
def test_oracle_against_the_reference_source_live(oracle, seed):
    """Not a stored vector: the reference's text is translated and EXECUTED in this test run on columns and options
    drawn from `seed` (cloud-optics options, solar-variability mode, partition sizes, inhomogeneity), and the C
    restatement is held to what it returns.  Build container only (needs /root/reference); 5 columns per case."""
    from geosradiation_gridcomp_b200.synthetic import make_columns
    from oracle.refexec import run
    rng = np.random.default_rng(seed)
    ih = int(rng.integers(0, 3))
    s = make_columns(5, 72, seed=seed)
    lw_opt = dict(psize=int(rng.integers(1, 6)), iceflg=int(rng.choice([0, 1, 2, 3, 4])), dudTs=bool(rng.integers(0, 2)))
    isolvar = int(rng.choice([-1, 0, 2, 3]))
    sw_opt = dict(rpart=int(rng.integers(0, 4)), iceflg=int(rng.choice([1, 2, 3, 4])), isolvar=isolvar,
                  normFlx=int(rng.integers(0, 2)), iaer=int(rng.choice([0, 10])), do_drfband=True)
    if isolvar == 2:
        sw_opt["indsolvar"] = [float(rng.uniform(0.14, 0.17)), float(rng.uniform(0.0, 2000.0))]
    if isolvar == 3:
        sw_opt["bndscl"] = rng.uniform(0.9, 1.1, 14)
    oracle.set_mcica(ih)
    try:
        ref_lw, got_lw = run.rrtmg_lw(s, ih=ih, **lw_opt), oracle.rrtmg_lw(s, **lw_opt)
        ref_sw, got_sw = run.rrtmg_sw(s, ih=ih, **sw_opt), oracle.rrtmg_sw(s, **sw_opt)
    finally:
        oracle.set_mcica(1)
    assert got_lw["rc"] == 0 and got_sw["rc"] == 0 and ref_sw["ret"][-1] == 0
    for ref, got, keys in ((ref_lw, got_lw, rc.LW_OUT), (ref_sw, got_sw, rc.SW_OUT)):
        for k in keys:
            if k == "clearCounts":
                np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{seed} {k}")
            else:
                assert rc.rel_err(got[k], ref[k]) <= TOL, (seed, k, lw_opt, sw_opt)
