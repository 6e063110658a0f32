"""Oracle restatement of the Run-phase glue (oracle/glue.c; GEOS_IrradGridComp.F90:3237-3371, :3486-3533,
GEOS_SolarGridComp.F90:6113-6223, :6395-6447) checked by properties: the native state is the synthetic
RRTMG state run backwards through the glue, so the glue must reproduce that state."""
import numpy as np
import pytest

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


# ---- an independent restatement of the Irrad Run-phase glue ---------------------------------------------------------
# GEOS_IrradGridComp.F90:3237-3371 in vectorised numpy, written from the Fortran without reference to oracle/glue.c:
# the vertical flip, TLEV, water paths, radius clamps of every ice/liquid option, unit conversions, aerosol
# absorption, ZM and the clean-up of negatives.  Same IEEE operations in the same order: bit-exact.
def _irrad_prepare_np(n, iceflg, liqflg):
    lm = n["lm"]
    ple, pl, t = n["ple"], n["pl"], n["t"]                       # (ncol, LM+1) / (ncol, LM), level 1 at the model top
    dp = ple[:, 1:] - ple[:, :-1]                                # DP(1..LM)
    tlev = np.empty_like(ple)                                    # TLEV(1..LM+1) -> columns 0..LM
    tlev[:, 1:lm] = (t[:, :-1] * dp[:, 1:] + t[:, 1:] * dp[:, :-1]) / (dp[:, :-1] + dp[:, 1:])
    tlev[:, lm] = n["t2m"]
    tlev[:, 0] = tlev[:, 1]
    fl = lambda a: np.asfortranarray(a[:, ::-1])                 # K = 1..LM  <-  LV = LM..1
    xx = 1.02 * 100 * dp
    out = dict(clwp=fl(xx * n["qliq"]), ciwp=fl(xx * n["qice"]))
    rel, rei = fl(n["rliq"]), fl(n["rice"])
    if liqflg == 0: rel = np.minimum(np.maximum(rel, 5.0), 10.0)
    elif liqflg == 1: rel = np.minimum(np.maximum(rel, 2.5), 60.0)
    lim = {0: (10.0, 30.0), 1: (13.0, 130.0), 2: (5.0, 131.0), 3: (5.0, 140.0)}
    if iceflg in lim: rei = np.minimum(np.maximum(rei, lim[iceflg][0]), lim[iceflg][1])
    elif iceflg == 4: rei = np.minimum(np.maximum(rei * 2., 1.0), 200.0)
    out.update(rel=rel, rei=rei)
    plev = np.empty_like(ple); tl = np.empty_like(ple)
    plev[:, :lm] = (ple[:, 1:] / 100.)[:, ::-1]                  # PLE_R(K-1) = PLE(LV)/100
    tl[:, :lm] = tlev[:, 1:][:, ::-1]                            # TLEV_R(K-1) = TLEV(LV+1)
    plev[:, lm] = ple[:, 0] / 100.
    tl[:, lm] = tlev[:, 0]
    out.update(plev=plev, tlev=tl, play=fl(pl / 100.), tlay=fl(t))
    pos = lambda a: np.where(a < 0., 0., a)
    out["h2ovmr"] = pos(fl(n["q"] / (1. - n["q"]) * (n["airmw"] / n["h2omw"])))
    out["o3vmr"] = pos(fl(n["o3"] * (n["airmw"] / n["o3mw"])))
    for dst, src in (("ch4vmr", "ch4"), ("n2ovmr", "n2o"), ("co2vmr", "co2"), ("cfc11vmr", "cfc11"),
                     ("cfc12vmr", "cfc12"), ("cfc22vmr", "hcfc22"), ("cldf", "fcld")):
        out[dst] = pos(fl(n[src]))
    out["o2vmr"] = np.full_like(out["play"], n["o2"])
    out["ccl4vmr"] = np.full_like(out["play"], n["ccl4"])
    out["tauaer"] = np.maximum(n["taua_lw"] - n["ssaa_lw"], 0.)[:, ::-1, :]
    zm = np.zeros_like(out["play"])
    for k in range(1, lm):
        # Fortran layer K is column k = K-1 here; PLE_R / TLEV_R are dimensioned (0:LM), so level K-1 is column k
        zm[:, k] = zm[:, k - 1] + n["rgas"] * tl[:, k] / n["grav"] * (out["play"][:, k - 1] - out["play"][:, k]) / plev[:, k]
    out.update(zm=zm, tsfc=n["ts"], alat=n["lats"], emis=np.repeat(n["emis"][:, None], 16, axis=1),
               cloudLM=lm - n["lcldlm"] + 1, cloudMH=lm - n["lcldmh"] + 1)
    return out


@pytest.mark.parametrize("iceflg,liqflg", [(3, 1), (0, 0), (1, 1), (2, 1), (4, 1)])
def test_irrad_prepare_against_independent_numpy(oracle, iceflg, liqflg):
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(300, 72, seed=31)
    got = oracle.irrad_prepare(n, iceflg=iceflg, liqflg=liqflg)
    mine = _irrad_prepare_np(n, iceflg, liqflg)
    assert (n["q"] < 0).any() and (n["fcld"] < 0).any()
    for k, v in mine.items():
        if k in ("cloudLM", "cloudMH"):
            assert got[k] == v, k
        else:
            np.testing.assert_array_equal(got[k], v, err_msg=k)


# GEOS_SolarGridComp.F90:6113-6223 (aerosol normalisation, flip, clamps, TLEV with TS at the surface, ZL) and the two
# drivers' epilogues (GEOS_IrradGridComp.F90:3486-3533, GEOS_SolarGridComp.F90:6395-6453), same treatment.
def _solar_prepare_np(n, iceflg, liqflg):
    lm = n["lm"]
    taua, ssaa, asya = n["taua_sw"], n["ssaa_sw"], n["asya_sw"]
    good = (taua > 0.) & (ssaa > 0.)
    with np.errstate(divide="ignore", invalid="ignore"):
        asy = np.where(good, asya / ssaa, 0.)
        ssa = np.where(good, ssaa / taua, 0.)
    tau = np.where(good, taua, 0.)
    ple, pl, t = n["ple"], n["pl"], n["t"]
    dpr = ple[:, 1:] - ple[:, :-1]
    fl = lambda a: np.asfortranarray(a[:, ::-1])
    out = dict(ciwp=(1.02 * 100 * fl(dpr)) * fl(n["qice"]), clwp=(1.02 * 100 * fl(dpr)) * fl(n["qliq"]))
    rei, rel = fl(n["rice"]), fl(n["rliq"])
    lim = {0: (10., 30.), 1: (13., 130.), 2: (5., 131.), 3: (5., 140.)}
    if iceflg in lim: rei = np.minimum(np.maximum(rei, lim[iceflg][0]), lim[iceflg][1])
    elif iceflg == 4: rei = np.minimum(np.maximum(rei * 2., 1.), 200.)
    if liqflg == 0: rel = np.minimum(np.maximum(rel, 10.), 30.)
    elif liqflg == 1: rel = np.minimum(np.maximum(rel, 2.5), 60.)
    tlev = np.empty_like(ple)
    tlev[:, 1:lm] = (t[:, :-1] * dpr[:, 1:] + t[:, 1:] * dpr[:, :-1]) / (dpr[:, :-1] + dpr[:, 1:])
    tlev[:, lm] = n["ts"]
    tlev[:, 0] = tlev[:, 1]
    plev, tl = fl(ple) / 100., fl(tlev)
    play = fl(pl) / 100.
    pos = lambda a: np.where(a < 0., 0., a)
    zl = np.zeros_like(play)
    for k in range(1, lm):                       # Fortran k = 2..LM, all arrays 1-based: index k here is k+1 there
        zl[:, k] = zl[:, k - 1] + n["rgas"] * tl[:, k] / n["grav"] * (play[:, k - 1] - play[:, k]) / plev[:, k]
    out.update(rei=rei, rel=rel, plev=plev, play=play, tlay=fl(t), zm=zl,
               h2ovmr=pos(fl(n["q"]) / (1. - fl(n["q"])) * (n["airmw"] / n["h2omw"])),
               o3vmr=pos(fl(n["o3"]) * (n["airmw"] / n["o3mw"])), ch4vmr=pos(fl(n["ch4"])),
               co2vmr=np.full_like(play, n["co2_fixed"]), o2vmr=np.full_like(play, n["o2"]), cld=pos(fl(n["fcld"])),
               tauaer=tau[:, ::-1, :], ssaaer=ssa[:, ::-1, :], asmaer=asy[:, ::-1, :])
    return out


@pytest.mark.parametrize("iceflg,liqflg", [(3, 1), (0, 0), (1, 1), (2, 1), (4, 1)])
def test_solar_prepare_against_independent_numpy(oracle, iceflg, liqflg):
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(300, 72, seed=32)
    n["taua_sw"][:7, 3:9, :] = 0.0                          # no aerosol: all three properties zeroed, no 0/0
    got = oracle.solar_prepare(n, iceflg=iceflg, liqflg=liqflg)
    for k, v in _solar_prepare_np(n, iceflg, liqflg).items():
        np.testing.assert_array_equal(got[k], v, err_msg=k)
    assert got["cloudLM"] == 72 - n["lcldlm"] + 1 and got["cloudMH"] == 72 - n["lcldmh"] + 1


def test_driver_epilogues_against_independent_numpy(oracle):
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(200, 72, seed=33)
    lw = oracle.rrtmg_lw(oracle.irrad_prepare(n))
    got = oracle.irrad_finish(n, lw)
    fl = lambda a: a[:, ::-1]
    for k, src, sign in (("flxu", "uflx", -1.), ("flxd", "dflx", 1.), ("flcu", "uflxc", -1.), ("flcd", "dflxc", 1.),
                         ("dfdts", "duflx_dTs", -1.), ("dfdtsc", "duflxc_dTs", -1.)):
        np.testing.assert_array_equal(got[k], sign * fl(lw[src]), err_msg=k)
    np.testing.assert_array_equal(got["sfcem"], -(lw["uflx"][:, 0] - lw["dflx"][:, 0] * (1. - n["emis"])))
    for j, k in enumerate(("cldtt", "cldhi", "cldmd", "cldlo")):
        np.testing.assert_array_equal(got[k], 1.0 - lw["clearCounts"][:, j] / float(140))
    sw = oracle.rrtmg_sw(oracle.solar_prepare(n))
    got = oracle.solar_finish(n, sw)
    np.testing.assert_array_equal(got["fsw"], fl(sw["swdflx"]) - fl(sw["swuflx"]))
    np.testing.assert_array_equal(got["fsc"], fl(sw["swdflxc"]) - fl(sw["swuflxc"]))
    np.testing.assert_array_equal(got["fswu"], fl(sw["swuflx"]))
    np.testing.assert_array_equal(got["fscu"], fl(sw["swuflxc"]))
    for j, k in enumerate(("cldts", "cldhs", "cldms", "cldls")):
        np.testing.assert_array_equal(got[k], 1. - sw["clearCounts"][:, j] / float(112))
    for k, nn, dd in (("cottp", "cotntp", "cotdtp"), ("cothp", "cotnhp", "cotdhp"), ("cotmp", "cotnmp", "cotdmp"),
                      ("cotlp", "cotnlp", "cotdlp")):
        ok = (sw[nn] > 0.) & (sw[dd] > 0.)
        with np.errstate(divide="ignore", invalid="ignore"):
            np.testing.assert_array_equal(got[k], np.where(ok, sw[nn] / sw[dd], n["undef"]), err_msg=k)
        assert ok.any() and (~ok).any()


def test_irrad_update_against_independent_numpy(oracle):
    """GEOS_IrradGridComp.F90:3604-3606, :3861, :3929-3990: the linear update between refreshes."""
    n = make_native_state(150, 72, seed=34)
    f = oracle.irrad_finish(n, oracle.rrtmg_lw(oracle.irrad_prepare(n)))
    rng = np.random.default_rng(3)
    ts_int = np.ascontiguousarray(n["ts"])
    tsinst = ts_int + rng.normal(0., 2., ts_int.shape)
    got = oracle.irrad_update(f, ts_int, tsinst)
    delt = (tsinst - ts_int)[:, None]
    flx_int, flc_int = f["flxd"] + f["flxu"], f["flcd"] + f["flcu"]
    lm = 72
    mine = dict(flx=flx_int + f["dfdts"] * delt, flc=flc_int + f["dfdtsc"] * delt,
                flxu=f["flxu"] + f["dfdts"] * delt, flcu=f["flcu"] + f["dfdtsc"] * delt, flxd=f["flxd"], flcd=f["flcd"],
                olr=-(flx_int[:, 0] + f["dfdts"][:, 0] * delt[:, 0]), olc=-(flc_int[:, 0] + f["dfdtsc"][:, 0] * delt[:, 0]),
                sfcem=f["sfcem"] - f["dfdts"][:, lm] * delt[:, 0], lws=flx_int[:, lm] + f["sfcem"],
                lcs=flc_int[:, lm] + f["sfcem"], flns=flx_int[:, lm] + f["dfdts"][:, lm] * delt[:, 0],
                flnsc=flc_int[:, lm] + f["dfdtsc"][:, lm] * delt[:, 0])
    for k, v in mine.items():
        np.testing.assert_array_equal(got[k], v, err_msg=k)
