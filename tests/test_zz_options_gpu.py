"""Options no earlier GPU test reaches (file sorts last): non-default McICA decorrelation lengths (initialize_cloud_subcol_gen, SH/cloud_subcol_gen.F90:108-129) through the
C ABI against the oracle: masks and clear counts bit-exact, fluxes within the flux tolerance."""
import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu

CORR = (0.6, 3.0, 5.0, -30.0, 0.3, 1.1, 10.0, 35.0)


def test_custom_correlation_lengths(rx, oracle):
    s = make_columns(200, 72, seed=61)
    oracle.set_mcica(1, CORR)
    rx.initialize_cloud_subcol_gen(*CORR)
    try:
        o_lw, o_sw = oracle.rrtmg_lw(s), oracle.rrtmg_sw(s)
        g_lw, g_sw = rx.run_lw(s), rx.run_sw(s)
    finally:
        oracle.set_mcica(1)
        rx._mcica["corr"] = None
        rx._apply_mcica()
    assert o_lw["rc"] == 0 and o_sw["rc"] == 0
    np.testing.assert_array_equal(g_lw["clearCounts"], o_lw["clearCounts"])
    np.testing.assert_array_equal(g_sw["clearCounts"], o_sw["clearCounts"])
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))
    for k in ("uflx", "dflx"):
        assert rel(g_lw[k], o_lw[k]) <= 1e-9, k
    for k in ("swuflx", "swdflx"):
        assert rel(g_sw[k], o_sw[k]) <= 1e-9, k
    # and the lengths matter: the default ones give other clear counts somewhere
    d_lw = oracle.rrtmg_lw(s)
    assert (d_lw["clearCounts"] != o_lw["clearCounts"]).any()


@pytest.mark.parametrize("iceflg", [0, 1, 2, 4])
def test_lw_ice_parameterisations(rx, oracle, iceflg):
    """The LW ice options other than the GEOS default 3 (LW/src/rrtmg_lw_cldprmc.F90:66-268)."""
    s = make_columns(192, 72, seed=67)
    o = oracle.rrtmg_lw(s, iceflg=iceflg, taps=("taucmc",))
    assert o["rc"] == 0
    g = rx.run_lw(s, iceflg=iceflg, taps=("taucmc",))
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))
    assert (o["taucmc"] > 0).sum() > 1000
    assert rel(g["taucmc"], o["taucmc"]) <= 1e-13
    np.testing.assert_array_equal(g["clearCounts"], o["clearCounts"])
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert rel(g[k], o[k]) <= 1e-9, k


@pytest.mark.parametrize("iceflg,liqflg", [(0, 0), (1, 1), (2, 1), (4, 1)])
def test_glue_radius_limits_of_every_option(rx, oracle, iceflg, liqflg):
    """The drivers clamp the effective radii differently for every cloud-optics option (IRR:3277-3296, SOL:6140-6170,
    and not the same way in the two drivers for liqflag 0); the refresh tests only use the GEOS defaults (3, 1)."""
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(300, 72, seed=71)
    for prep_o, prep_g in ((oracle.irrad_prepare, rx.irrad_prepare), (oracle.solar_prepare, rx.solar_prepare)):
        o = prep_o(n, iceflg=iceflg, liqflg=liqflg)
        g = prep_g(n, iceflg=iceflg, liqflg=liqflg)
        for k in ("rei", "rel", "ciwp", "clwp", "play", "tlay"):
            np.testing.assert_array_equal(g[k], o[k], err_msg=f"{prep_o.__name__} {k}")
