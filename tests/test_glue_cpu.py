"""Oracle restatement of the Run-phase glue (oracle/glue.c; GEOS_IrradGridComp.F90:3237-3371, :3486-3533,
GEOS_SolarGridComp.F90:6113-6223, :6395-6447) checked by properties: the native state is the synthetic
RRTMG state run backwards through the glue, so the glue must reproduce that state."""
import numpy as np

from geosradiation_gridcomp_b200.synthetic import make_columns, make_native_state


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def test_irrad_prepare_inverts_the_native_state(oracle):
    n = make_native_state(200, 72, seed=3)
    ref = make_columns(200, 72, seed=3)
    s = oracle.irrad_prepare(n)
    for k in ("play", "plev", "clwp", "ciwp"):
        assert rel(s[k], ref[k]) < 1e-14, k
    np.testing.assert_array_equal(s["tlay"], ref["tlay"])
    np.testing.assert_array_equal(s["tsfc"], ref["tsfc"])
    np.testing.assert_array_equal(s["emis"], ref["emis"])
    assert np.abs(s["tlev"][:, 1:] - ref["tlev"][:, 1:]).max() < 1e-11   # pressure-weighted interpolation
    np.testing.assert_array_equal(s["tlev"][:, 0], n["t2m"])
    assert (s["cloudLM"], s["cloudMH"]) == (ref["cloudLM"], ref["cloudMH"])
    # planted negatives are cleaned up, everything else survives the q <-> vmr round trip
    for k, nk in (("h2ovmr", "q"), ("o3vmr", "o3"), ("cldf", "fcld")):
        planted = n[nk][:, ::-1] < 0
        assert planted.any() and (s[k][planted] == 0).all(), k
        assert rel(s[k][~planted], ref[k][~planted]) < 1e-13, k
    # radius limits of iceflag 3 / liqflag 1
    assert s["rel"].min() >= 2.5 and s["rel"].max() <= 60.0 and (s["rel"] == 2.5).any()
    assert s["rei"].min() >= 5.0 and s["rei"].max() <= 140.0 and (s["rei"] == 5.0).any()
    # absorption optical depth = extinction - scattering
    assert rel(s["tauaer"], ref["tauaer_lw"]) < 1e-12
    # layer heights: relative, increasing, hydrostatic
    assert (s["zm"][:, 0] == 0).all() and (np.diff(s["zm"], axis=1) > 0).all()
    assert rel(s["zm"], ref["zm"]) < 1e-4     # the synthetic generator rounds RGAS / GRAV differently


def test_irrad_finish_conventions(oracle):
    n = make_native_state(64, 72, seed=4)
    s = oracle.irrad_prepare(n)
    o = oracle.rrtmg_lw(s)
    assert o["rc"] == 0
    f = oracle.irrad_finish(n, o)
    np.testing.assert_array_equal(f["flxu"], -o["uflx"][:, ::-1])
    np.testing.assert_array_equal(f["flxd"], o["dflx"][:, ::-1])
    np.testing.assert_array_equal(f["dfdtsc"], -o["duflxc_dTs"][:, ::-1])
    assert (f["flxu"] <= 0).all() and (f["flxd"] >= 0).all()
    np.testing.assert_array_equal(f["cldtt"], 1.0 - o["clearCounts"][:, 0] / 140.0)
    assert ((f["cldtt"] >= f["cldhi"] - 1e-15) & (f["cldtt"] >= f["cldlo"] - 1e-15)).all()
    # emitted = upward - reflected downward, positive downward convention
    np.testing.assert_allclose(f["sfcem"], -(o["uflx"][:, 0] - o["dflx"][:, 0] * (1 - n["emis"])), rtol=0, atol=0)
    assert (f["sfcem"] < 0).all()


def test_solar_glue(oracle):
    n = make_native_state(96, 72, seed=6)
    ref = make_columns(96, 72, seed=6)
    s = oracle.solar_prepare(n)
    for k in ("play", "plev", "clwp", "ciwp"):
        assert rel(s[k], ref[k]) < 1e-14, k
    lit = ref["tauaer_sw"] > 0
    assert rel(s["tauaer"], ref["tauaer_sw"]) == 0
    assert rel(s["ssaaer"][lit], ref["ssaaer"][lit]) < 1e-14 and (s["ssaaer"][~lit] == 0).all()
    assert rel(s["asmaer"][lit], ref["asmaer"][lit]) < 1e-14 and (s["asmaer"][~lit] == 0).all()
    assert s["rel"].min() >= 2.5 and s["rei"].min() >= 5.0
    assert (s["cloudLM"], s["cloudMH"]) == (ref["cloudLM"], ref["cloudMH"])
    o = oracle.rrtmg_sw(s)
    assert o["rc"] == 0
    f = oracle.solar_finish(n, o)
    np.testing.assert_array_equal(f["fswu"], o["swuflx"][:, ::-1])
    np.testing.assert_array_equal(f["fsw"], (o["swdflx"] - o["swuflx"])[:, ::-1])
    np.testing.assert_array_equal(f["fsc"], (o["swdflxc"] - o["swuflxc"])[:, ::-1])
    cloudy = (o["cotntp"] > 0) & (o["cotdtp"] > 0)
    assert cloudy.any() and (~cloudy).any()
    assert (f["cottp"][~cloudy] == n["undef"]).all()
    np.testing.assert_array_equal(f["cottp"][cloudy], o["cotntp"][cloudy] / o["cotdtp"][cloudy])
    np.testing.assert_array_equal(f["cldts"], 1.0 - o["clearCounts"][:, 0] / 112.0)
