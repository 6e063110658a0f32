"""THE PIN of the oracle: the C restatement (oracle/*.c) against numbers produced by the REFERENCE'S OWN SOURCE TEXT.

tests/golden/rrtmg_refexec_golden.npz was written by tests/golden/make_golden_from_refexec.py, which executes the
Fortran of /root/reference (RRTMG LW + SW, McICA, NRLSSI2: 91 files, read where they lie) through the translator
oracle/refexec/f90py.py - no Fortran compiler exists in this image, so this is how the reference runs here.  The
cases (tests/golden/refexec_cases.py) cover the GEOS defaults at L72 and L181, every LW and SW cloud-optics option,
every solar-variability mode, the three condensate inhomogeneity options, non-default decorrelation lengths, ragged
partitions, the SOLAR_RADVAL build of rrtmg_sw (its 120 extra diagnostics), and every intermediate the oracle taps (jp / jt / jt1 / indices, interpolation factors, the McICA
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
            for k in rc.SW_OUT + (tuple(rc.SW_TAPS) if c["taps"] else ()) + (("radval",) if c["sw"].get("radval") else ()):
                assert f"{name}/sw/{k}" in keys, (name, k)
    # the cases are not trivially cloud-free: cloudy sub-columns exist, and the clear-sky stream differs from the total
    assert (golden["default/lw/clearCounts"][:, 0] < 140).sum() >= 10
    assert np.abs(golden["default/lw/uflx"] - golden["default/lw/uflxc"]).max() > 10.0
    assert np.abs(golden["default/sw/swdflx"] - golden["default/sw/swdflxc"]).max() > 0.1
    assert golden["taps/lw/cldymc"].sum() > 1000 and golden["taps/sw/cldymc"].sum() > 1000
    # the SOLAR_RADVAL cases: every one of the 120 diagnostics is non-zero somewhere, liquid and ice, all super-layers
    rv = np.concatenate([golden[f"{n}/sw/radval"] for n in rc.CASES if n.startswith("radval")])
    assert rv.shape[1] == 120 and (np.abs(rv).max(axis=0) > 0).all()


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


# ---- the drivers' Run-phase glue (IRR:3238-3371, 3487-3533; SOL:6116-6219, 6395-6454) ------------------------------------
def _refresh_state():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_from_refexec as gen
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    return gen, make_native_state(gen.REFRESH_NCOL, 72, seed=gen.REFRESH_SEED)


def test_oracle_refresh_equals_the_reference_text_refresh(oracle, golden):
    """A whole refresh of each driver - native GEOS state in, native exports out - as the reference's text computes it
    (its glue LINES around its RRTMG sources, all executed through oracle/refexec; golden keys refresh/*) against the C
    restatement's chain glue.c -> lw.c / sw.c -> glue.c: the prepared RRTMG arguments bit for bit, the exports within
    1e-12, cloud fractions and the MAPL_UNDEF pattern of the optical thicknesses exactly."""
    gen, n = _refresh_state()
    s = oracle.irrad_prepare(n)
    for k in gen.IRR_PREPARED:
        np.testing.assert_array_equal(s["tauaer" if k == "tauaer_lw" else k], golden[f"refresh/irr_prepared/{k}"], err_msg=k)
    assert (s["cloudLM"], s["cloudMH"]) == (int(golden["refresh/irr_prepared/cloudLM"]), int(golden["refresh/irr_prepared/cloudMH"]))
    f = oracle.irrad_finish(n, oracle.rrtmg_lw(s))
    for k in gen.IRR_EXPORTS:
        ref = golden[f"refresh/irr/{k}"]
        if k.startswith("cld"):
            np.testing.assert_array_equal(f[k], ref, err_msg=k)
        else:
            assert rc.rel_err(f[k], ref) <= TOL, k
    s = oracle.solar_prepare(n)
    for k in gen.SOL_PREPARED:
        np.testing.assert_array_equal(s[k], golden[f"refresh/sol_prepared/{k}"], err_msg=k)
    f = oracle.solar_finish(n, oracle.rrtmg_sw(s))
    for k in gen.SOL_EXPORTS:
        ref = golden[f"refresh/sol/{k}"]
        if k.startswith("cld"):
            np.testing.assert_array_equal(f[k], ref, err_msg=k)
        elif k.startswith("cot"):
            np.testing.assert_array_equal(f[k] == n["undef"], ref == n["undef"], err_msg=k)
            assert rc.rel_err(np.where(ref == n["undef"], 0.0, f[k]), np.where(ref == n["undef"], 0.0, ref)) <= TOL, k
        else:
            assert rc.rel_err(f[k], ref) <= TOL, k


@pytest.mark.skipif(not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")), reason="no reference tree on this machine")
@pytest.mark.parametrize("iceflg,liqflg", [(3, 1), (0, 0), (1, 1), (2, 1), (4, 1)])
def test_oracle_glue_equals_the_reference_lines_live(oracle, iceflg, liqflg):
    """The glue line ranges of the two drivers, read from the reference files in this test run and executed: every radius
    clamp of every cloud-optics option, with and without aerosols, CO2 as a field or as the fixed scalar - bit for bit."""
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    from oracle.refexec import glue
    n = make_native_state(20, 72, seed=71)
    variants = [n, {k: (None if k in ("taua_lw", "ssaa_lw", "taua_sw", "ssaa_sw", "asya_sw", "co2") else v) for k, v in n.items()}]
    for m in variants:
        r, o = glue.irrad_prepare(m, iceflg, liqflg), oracle.irrad_prepare(m, iceflg, liqflg)
        for k in r:
            np.testing.assert_array_equal(np.asarray(r[k]), np.asarray(o[k]), err_msg=f"irrad {k}")
        r, o = glue.solar_prepare(m, iceflg, liqflg), oracle.solar_prepare(m, iceflg, liqflg)
        for k in r:
            if k in o:
                np.testing.assert_array_equal(np.asarray(r[k]), np.asarray(o[k]), err_msg=f"solar {k}")
        assert r["adjes"] == m["dist"]
    s = oracle.irrad_prepare(n)
    lw = oracle.rrtmg_lw(s)
    r, o = glue.irrad_finish(n, lw), oracle.irrad_finish(n, lw)
    for k in r:
        np.testing.assert_array_equal(r[k], o[k], err_msg=f"irrad finish {k}")
    s = oracle.solar_prepare(n)
    sw = oracle.rrtmg_sw(s)
    r, o = glue.solar_finish(n, sw), oracle.solar_finish(n, sw)
    for k in r:
        np.testing.assert_array_equal(r[k], o[k], err_msg=f"solar finish {k}")


@pytest.mark.skipif(not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")), reason="no reference tree on this machine")
def test_oracle_irrad_update_equals_the_reference_lines_live(oracle):
    """The between-refresh update of the LW exports (GEOS_IrradGridComp.F90:3604, 3606, 3861 and the USE_RRTMG branch
    :3932-3992, read from the file and executed with its export pointers associated): all thirteen exports the library's
    rrtmgx_irrad_update produces, bit for bit, plus the MAPL_UNDEF rule of the cloud-free composites."""
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    from oracle.refexec import glue
    n = make_native_state(16, 72, seed=28)
    f = oracle.irrad_finish(n, oracle.rrtmg_lw(oracle.irrad_prepare(n)))
    rng = np.random.default_rng(5)
    ts_int = np.asfortranarray(n["ts"])
    tsinst = np.asfortranarray(n["ts"] + rng.normal(0, 1.5, n["ncol"]))
    r = glue.irrad_update(f, ts_int, tsinst, undef=n["undef"], cldtt=f["cldtt"])
    o = oracle.irrad_update(f, ts_int, tsinst)
    assert len(o) == 13
    for k in o:
        np.testing.assert_array_equal(r[k], o[k], err_msg=k)
    np.testing.assert_array_equal(r["dsfdts"], -f["dfdts"][:, -1])
    clear = f["cldtt"] <= 0.05
    assert clear.any() and (~clear).any()
    np.testing.assert_array_equal(r["olcc5"][clear], r["olc"][clear])
    assert (r["olcc5"][~clear] == n["undef"]).all()


def test_heating_rate_formula_equals_the_reference_lines(golden):
    """RADLW / RADSW of the parent component (GEOS_RadiationGridComp.F90:811, 813-814, executed from the file by
    oracle/refexec/glue.py on the golden refresh: keys refresh/rad/*, K/s, top-down) against the restatement the heating
    rate tests compare rrtmgx_heating_rate with (K/day, surface first): the bar of north_star, 1e-6 K/day - what is
    seen is the rounding of the unit conversions, ~1e-14 K/day."""
    import make_golden_from_refexec as gen
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(gen.REFRESH_NCOL, 72, seed=gen.REFRESH_SEED)
    fnet, plev = rc.heating_inputs(golden, n)
    for k in ("radlw", "radsw"):
        ref = golden[f"refresh/rad/{k}"][:, ::-1] * 86400.0      # K/day, surface first
        got = rc.heating_formula(fnet[k], plev, gen.RAD_GRAV, gen.RAD_CP)
        assert np.abs(ref).max() > 10.0
        assert np.abs(got - ref).max() <= 1e-6, k
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), k
    # DTDT (:801-802) is the sum of the two flux divergences times g / cp
    d = (golden["refresh/rad/radlw"] + golden["refresh/rad/radsw"]) * (np.asarray(n["ple"])[:, 1:] - np.asarray(n["ple"])[:, :-1])
    assert rc.rel_err(d, golden["refresh/rad/dtdt"]) <= 1e-12


@pytest.mark.skipif(not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")), reason="no reference tree on this machine")
def test_heating_rate_golden_is_what_the_reference_lines_give_live(golden):
    import make_golden_from_refexec as gen
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    from oracle.refexec import glue
    n = make_native_state(gen.REFRESH_NCOL, 72, seed=gen.REFRESH_SEED)
    flw = golden["refresh/irr/flxu"] + golden["refresh/irr/flxd"]
    hr = glue.heating_rates(n["ple"], flw, gen.rad_fsw(n, golden["refresh/sol/fsw"]), gen.RAD_GRAV, gen.RAD_CP)
    for k in gen.RAD_EXPORTS:
        np.testing.assert_array_equal(hr[k], golden[f"refresh/rad/{k}"], err_msg=k)
