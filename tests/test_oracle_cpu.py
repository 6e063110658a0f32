"""CPU tests that pin the oracle (oracle/, the fp64 restatement of the reference Fortran).

The reference ships no golden vectors or known-answer tests for this path, so the
pins are: an independent restatement of the KISS generator, physical/algebraic properties that
hold for the reference algorithm, and the committed fixture tests/golden/rrtmg_golden_L72.npz."""
import os

import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

HERE = os.path.dirname(os.path.abspath(__file__))


def _i32(x):
    return ((x + 2**31) % 2**32) - 2**31


def kiss_python(seeds, n):
    """SH/cloud_subcol_gen.F90:568-575 in plain Python integers with explicit int32 wraparound."""
    s1, s2, s3, s4 = (int(np.int32(v)) for v in seeds)
    out = []
    for _ in range(n):
        s1 = _i32(69069 * s1 + 1327217885)
        u = s2 & 0xffffffff
        u ^= (u << 13) & 0xffffffff
        u ^= u >> 17
        u ^= (u << 5) & 0xffffffff
        s2 = _i32(u)
        s3 = _i32(18000 * (s3 & 65535) + ((s3 & 0xffffffff) >> 16))
        s4 = _i32(30903 * (s4 & 65535) + ((s4 & 0xffffffff) >> 16))
        kiss = _i32(s1 + s2 + _i32((s3 & 0xffffffff) << 16) + s4)
        out.append(kiss * 2.328306e-10 + 0.5)
    return np.array(out)


def test_kiss_matches_independent_restatement(oracle):
    for seeds in ([1, 2, 3, 4], [123456789, 362436069, 521288629, 916191069], [2147483646, 7, 65535, 65536]):
        got, _ = oracle.rng_kiss(seeds, 500)
        np.testing.assert_array_equal(got, kiss_python(seeds, 500))
        assert got.min() > 0.0 and got.max() < 1.0


def test_mcica_statistics_and_counts(oracle):
    nlay, ncol, nsub = 10, 6, 4000
    play = np.tile(np.linspace(1000.0, 100.0, nlay)[:, None], (1, ncol)) + 0.0123 * np.arange(ncol)[None, :]
    zmid = np.tile(np.linspace(100.0, 15000.0, nlay)[:, None], (1, ncol))
    cld = np.zeros((nlay, ncol))
    fr = np.array([0.1, 0.3, 0.5, 0.7, 0.95, 1.0])
    cld[4, :] = fr
    ciwp = np.zeros((nlay, ncol))
    clwp = np.where(cld > 0, 50.0, 0.0)
    mask, ci, cw = oracle.generate_stochastic_clouds(zmid, np.zeros(ncol), 172, play, cld, ciwp, clwp, nsub)
    assert mask[:4].sum() == 0 and mask[5:].sum() == 0
    got = mask[4].mean(axis=0)
    assert np.all(np.abs(got - fr) < 4.0 * np.sqrt(fr * (1 - fr) / nsub) + 1e-12)
    # inhomogeneous condensate (beta table): the scaling factor has mean ~ 1
    inc = cw[4][mask[4] > 0]
    assert abs(inc.mean() / 50.0 - 1.0) < 0.08
    assert ci.sum() == 0.0
    cc = oracle.clear_counts(mask, 3, 6)
    assert cc.min() >= 0 and cc.max() <= nsub
    np.testing.assert_array_equal(cc[0], nsub - mask.any(axis=0).sum(axis=0))
    np.testing.assert_array_equal(cc[2], cc[0])                  # cloud only in the middle super-layer
    np.testing.assert_array_equal(cc[1], nsub)
    np.testing.assert_array_equal(cc[3], nsub)


def test_mcica_maximum_overlap_of_adjacent_layers(oracle):
    nlay, ncol, nsub = 6, 3, 1500
    play = np.tile(np.linspace(900.0, 400.0, nlay)[:, None], (1, ncol)) + 0.0377 * np.arange(ncol)[None, :]
    zmid = np.tile((1000.0 + 1e-3 * np.arange(nlay))[:, None], (1, ncol))     # dz -> 0: alpha -> 1
    cld = np.zeros((nlay, ncol))
    cld[2:4, :] = 0.4
    clwp = np.where(cld > 0, 20.0, 0.0)
    oracle.set_mcica(0)
    try:
        mask, _, _ = oracle.generate_stochastic_clouds(zmid, np.zeros(ncol), 10, play, cld, np.zeros_like(cld),
                                                       clwp, nsub)
    finally:
        oracle.set_mcica(1)
    np.testing.assert_array_equal(mask[2], mask[3])


@pytest.fixture(scope="module")
def state():
    return make_columns(96, 72, seed=20260118)


def test_lw_properties(oracle, state):
    s = dict(state)
    s["emis"] = np.asfortranarray(np.ones_like(s["emis"]))
    taps = ("jp", "jt", "jt1", "fac00", "fac01", "fac10", "fac11", "pfracs", "laytrop", "cldymc")
    o = oracle.rrtmg_lw(s, taps=taps)
    assert o["rc"] == 0
    assert o["jp"].min() >= 1 and o["jp"].max() <= 58
    assert o["jt"].min() >= 1 and o["jt"].max() <= 4 and o["jt1"].min() >= 1 and o["jt1"].max() <= 4
    np.testing.assert_allclose(o["fac00"] + o["fac01"] + o["fac10"] + o["fac11"], 1.0, rtol=0, atol=1e-12)
    # Planck fractions of a band sum to 1 (summed unweighted by cmbgb, rrtmg_lw_init.F90:425-436)
    ngs = [0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140]
    for b in range(16):
        tot = o["pfracs"][:, ngs[b]:ngs[b + 1], :].sum(axis=1)
        ok = tot > 0          # bands 12, 15 have no upper-atmosphere source
        assert np.all(np.abs(tot[ok] - 1.0) < 2e-3), b + 1
    # black surface: the upward flux at the surface is the Planck emission 10-3250 cm-1 of tsfc
    sigma_t4 = 5.670374e-8 * s["tsfc"] ** 4
    assert np.all(np.abs(o["uflx"][:, 0] / sigma_t4 - 1.0) < 0.02)
    assert np.all(o["dflx"][:, -1] == 0.0) and np.all(o["dflxc"][:, -1] == 0.0)
    # columns without cloud: all-sky == clear-sky bit for bit, every subcolumn counted clear
    clear = ~(s["cldf"] > 0).any(axis=1)
    assert clear.sum() > 10
    for a, c in (("uflx", "uflxc"), ("dflx", "dflxc"), ("duflx_dTs", "duflxc_dTs")):
        np.testing.assert_array_equal(o[a][clear], o[c][clear])
    np.testing.assert_array_equal(o["clearCounts"][clear], 140)
    assert o["clearCounts"].min() >= 0 and o["clearCounts"].max() <= 140
    # clouds reduce the outgoing longwave radiation
    assert np.all(o["uflx"][:, -1] <= o["uflxc"][:, -1] + 1e-9)
    # dF/dTs is positive and decays upward
    assert np.all(o["duflx_dTs"] >= 0) and np.all(np.diff(o["duflx_dTs"], axis=1) <= 1e-12)


def test_sw_properties(oracle, state):
    s = state
    o = oracle.rrtmg_sw(s, normFlx=0, do_drfband=True, taps=("jp", "jt", "jt1", "fac00", "fac01", "fac10", "fac11"))
    assert o["rc"] == 0
    np.testing.assert_allclose(o["fac00"] + o["fac01"] + o["fac10"] + o["fac11"], 1.0, rtol=0, atol=1e-12)
    toa = o["swdflx"][:, -1]
    # NRLSSI2 mean-cycle irradiance scaled to scon (rrtmg_sw_rad.F90:1050-1060)
    np.testing.assert_allclose(toa, s["scon"] * s["adjes"] * s["coszen"], rtol=2e-5)
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc"):
        assert np.all(o[k] >= 0) and np.all(o[k] <= toa[:, None] * (1 + 1e-12))
    # energy closure: the atmosphere absorbs, never emits
    absorbed = (o["swdflx"][:, -1] - o["swuflx"][:, -1]) - (o["swdflx"][:, 0] - o["swuflx"][:, 0])
    assert np.all(absorbed > 0)
    # surface components add up to the surface fluxes
    tot = o["nirr"] + o["nirf"] + o["parr"] + o["parf"] + o["uvrr"] + o["uvrf"]
    np.testing.assert_allclose(tot, o["swdflx"][:, 0], rtol=1e-12)
    np.testing.assert_allclose(o["fswband"].sum(axis=1), o["swdflx"][:, 0] - o["swuflx"][:, 0], rtol=1e-12)
    np.testing.assert_allclose((o["drband"] + o["dfband"]).sum(axis=1), o["swdflx"][:, 0], rtol=1e-12)
    clear = ~(s["cldf"] > 0).any(axis=1)
    np.testing.assert_array_equal(o["swuflx"][clear], o["swuflxc"][clear])
    np.testing.assert_array_equal(o["swdflx"][clear], o["swdflxc"][clear])
    np.testing.assert_array_equal(o["clearCounts"][clear], 112)
    assert np.all(o["cotntp"][clear] == 0) and np.all(o["cotdtp"][clear] == 0)
    cloudy = o["clearCounts"][:, 0] < 112
    assert np.all(o["cotdtp"][cloudy] > 0) and np.all(o["cotntp"][cloudy] > 0)
    # normalised fluxes are the same fluxes divided by the TOA downward flux (:1769-1798)
    n = oracle.rrtmg_sw(s, normFlx=1)
    np.testing.assert_allclose(n["swdflx"][:, -1], 1.0, rtol=1e-15)
    np.testing.assert_allclose(n["swuflx"] * toa[:, None], o["swuflx"], rtol=1e-13)


def test_results_do_not_depend_on_partition_size(oracle):
    s = make_columns(37, 72, seed=5)
    a, b = oracle.rrtmg_lw(s, psize=4), oracle.rrtmg_lw(s, psize=9)
    for k in ("uflx", "dflx", "uflxc", "dflxc", "clearCounts"):
        np.testing.assert_array_equal(a[k], b[k])
    a, b = oracle.rrtmg_sw(s, rpart=0), oracle.rrtmg_sw(s, rpart=5)
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "clearCounts", "fswband"):
        np.testing.assert_array_equal(a[k], b[k])


def test_reduced_tables(oracle):
    rw = oracle.table("lw", "rwgt")
    assert rw is not None and rw.size == 256 and rw.min() > 0
    # band 1 of the LW keeps 10 of 16 g-points; fracrefa sums to 1 before and after the reduction
    fa = oracle.table("lw", "fracrefa", 1)
    assert fa.size == 10 and abs(fa.sum() - 1.0) < 1e-4
    for band, ng in ((16, 6), (17, 12), (24, 8), (29, 12)):
        sf = oracle.table("sw", "sfluxref", band)
        assert sf is not None and sf.size % ng == 0 and sf.min() >= 0


def test_oracle_matches_committed_golden_vectors(oracle):
    g = np.load(os.path.join(HERE, "golden", "rrtmg_golden_L72.npz"))
    s = make_columns(int(g["ncol"]), int(g["nlay"]), seed=int(g["seed"]))
    lw, sw = oracle.rrtmg_lw(s, taps=("jp", "jt", "jt1", "laytrop")), oracle.rrtmg_sw(s, taps=("jp", "jt", "jt1", "laytrop"))
    for pre, o in (("lw_", lw), ("sw_", sw)):
        for k in g.files:
            if not k.startswith(pre):
                continue
            ref, got = g[k], o[k[3:]]
            if ref.dtype.kind in "iu":
                np.testing.assert_array_equal(got, ref, err_msg=k)
            else:
                np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-300, err_msg=k)


# ---- an independent restatement of the SW two-stream and adding routines ----------------------------------
# Written straight from SW/src/rrtmg_sw_spcvmc.F90 (reftra_sw :1115-1370, vrtqdr_sw :1374-1588) in vectorised
# numpy, on purpose without looking at oracle/sw.c: two transcriptions that agree pin the formulas against
# slips of sign, index or branch (not the last bit: numpy's exp is not libm's).
def _reftra_np(zto1, zw, zg, prmuz):
    eps, od_lo, zwcrit = 1.e-08, 0.06, 0.9999995
    zg3 = 3. * zg
    zgamma1 = (8. - zw * (5. + zg3)) * 0.25
    zgamma2 = 3. * (zw * (1. - zg)) * 0.25
    zgamma3 = (2. - zg3 * prmuz) * 0.25
    zgamma4 = 1. - zgamma3
    zwo = zw / (1. - (1. - zw) * (zg / (1. - zg)) ** 2)
    # conservative scattering
    za = zgamma1 * prmuz
    za1 = za - zgamma3
    zgt = zgamma1 * zto1
    ze2c = np.exp(-np.minimum(zto1 / prmuz, 500.))
    ref_c = (zgt - za1 * (1. - ze2c)) / (1. + zgt)
    tra_c = 1. - ref_c
    refd_c = zgt / (1. + zgt)
    trad_c = 1. - refd_c
    one = ze2c == 1.
    ref_c, tra_c = np.where(one, 0., ref_c), np.where(one, 1., tra_c)
    refd_c, trad_c = np.where(one, 0., refd_c), np.where(one, 1., trad_c)
    # non-conservative scattering
    with np.errstate(invalid="ignore", divide="ignore"):
        za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3
        za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4
        zrk = np.sqrt(zgamma1 ** 2 - zgamma2 ** 2)
        zrp = zrk * prmuz
        zrp1, zrm1, zrk2 = 1. + zrp, 1. - zrp, 2. * zrk
        zrpp = 1. - zrp * zrp
        zrkg = zrk + zgamma1
        zr1 = zrm1 * (za2 + zrk * zgamma3)
        zr2 = zrp1 * (za2 - zrk * zgamma3)
        zr3 = zrk2 * (zgamma3 - za2 * prmuz)
        zr4 = zrpp * zrkg
        zr5 = zrpp * (zrk - zgamma1)
        zt1 = zrp1 * (za1 + zrk * zgamma4)
        zt2 = zrm1 * (za1 - zrk * zgamma4)
        zt3 = zrk2 * (zgamma4 + za1 * prmuz)
        zbeta = (zgamma1 - zrk) / zrkg
        ze1 = np.minimum(zrk * zto1, 5.)
        ze2 = np.minimum(zto1 / prmuz, 5.)
        zem1 = np.where(ze1 <= od_lo, 1. - ze1 + 0.5 * ze1 * ze1, np.exp(-ze1))
        zep1 = 1. / zem1
        zem2 = np.where(ze2 <= od_lo, 1. - ze2 + 0.5 * ze2 * ze2, np.exp(-ze2))
        zep2 = 1. / zem2
        zdenr = zr4 * zep1 + zr5 * zem1
        zdent = zr4 * zep1 + zr5 * zem1
        small = (zdenr >= -eps) & (zdenr <= eps)
        ref_n = np.where(small, eps, zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr)
        tra_n = np.where(small, zem2, zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent)
        zemm = zem1 * zem1
        zdend = 1. / ((1. - zbeta * zemm) * zrkg)
        refd_n = zgamma2 * (1. - zemm) * zdend
        trad_n = zrk2 * zem1 * zdend
    cons = zwo >= zwcrit
    return (np.where(cons, ref_c, ref_n), np.where(cons, refd_c, refd_n), np.where(cons, tra_c, tra_n),
            np.where(cons, trad_c, trad_n))


def _vrtqdr_np(pref, prefd, ptra, ptrad, pdbt, ptdbt):
    """Arrays [level or layer][...], index 0 = top (jk = 1); level arrays have klev+1 entries."""
    klev = pdbt.shape[0]
    prup, prupd = np.empty_like(pref), np.empty_like(pref)
    prup[klev], prupd[klev] = pref[klev], prefd[klev]
    zreflect = 1. / (1. - prefd[klev] * prefd[klev - 1])
    prup[klev - 1] = pref[klev - 1] + (ptrad[klev - 1] * ((ptra[klev - 1] - pdbt[klev - 1]) * prefd[klev] +
                                                            pdbt[klev - 1] * pref[klev])) * zreflect
    prupd[klev - 1] = prefd[klev - 1] + ptrad[klev - 1] * ptrad[klev - 1] * prefd[klev] * zreflect
    for jk in range(1, klev):                    # do jk = 1, klev-1
        ikp = klev + 1 - jk                      # 1-based
        ikx = ikp - 1
        zr = 1. / (1. - prupd[ikp - 1] * prefd[ikx - 1])
        prup[ikx - 1] = pref[ikx - 1] + (ptrad[ikx - 1] * ((ptra[ikx - 1] - pdbt[ikx - 1]) * prupd[ikp - 1] +
                                                            pdbt[ikx - 1] * prup[ikp - 1])) * zr
        prupd[ikx - 1] = prefd[ikx - 1] + ptrad[ikx - 1] * ptrad[ikx - 1] * prupd[ikp - 1] * zr
    ptdn, prdnd = np.empty_like(pref), np.empty_like(pref)
    ptdn[0], prdnd[0] = 1., 0.
    ptdn[1], prdnd[1] = ptra[0], prefd[0]
    for jk in range(2, klev + 1):                # do jk = 2, klev
        ikp = jk + 1
        zr = 1. / (1. - prefd[jk - 1] * prdnd[jk - 1])
        ptdn[ikp - 1] = ptdbt[jk - 1] * ptra[jk - 1] + (ptrad[jk - 1] * ((ptdn[jk - 1] - ptdbt[jk - 1]) +
                                                                        ptdbt[jk - 1] * pref[jk - 1] * prdnd[jk - 1])) * zr
        prdnd[ikp - 1] = prefd[jk - 1] + ptrad[jk - 1] * ptrad[jk - 1] * prdnd[jk - 1] * zr
    zr = 1. / (1. - prdnd * prupd)
    pfu = (ptdbt * prup + (ptdn - ptdbt) * prupd) * zr
    pfd = ptdbt + (ptdn - ptdbt + ptdbt * prup * prdnd) * zr
    return pfd, pfu


def test_two_stream_and_adding_against_independent_numpy(oracle):
    import ctypes as C
    L = oracle.lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(8)
    ncol, nlay, ng = 6, 40, 112
    shp = (nlay, ng, ncol)
    tau = np.asfortranarray(10.0 ** rng.uniform(-6, 2.5, shp))
    w = np.asfortranarray(np.where(rng.random(shp) < 0.25, 1.0 - 10.0 ** rng.uniform(-9, -5, shp), rng.uniform(0, 1, shp)))
    g = np.asfortranarray(np.where(rng.random(shp) < 0.3, 0.0, rng.uniform(0, 0.95, shp)))
    mu = np.ascontiguousarray(rng.uniform(0.02, 1.0, ncol))
    out = [np.zeros((nlay + 1, ng, ncol), order="F") for _ in range(4)]
    p = lambda a: a.ctypes.data_as(dp)
    assert L.oracle_reftra_sw(ncol, nlay, p(g), p(mu), p(tau), p(w), *[p(o) for o in out]) == 0
    ref = _reftra_np(tau, w, g, mu[None, None, :])
    # reflectances and transmittances are O(1); thin layers are differences of O(1) terms, so their last
    # bits (numpy exp vs libm exp) show up as ~1e-16 absolute, not as a relative error
    for o, r in zip(out, ref):
        np.testing.assert_allclose(o[:nlay], r, rtol=2e-12, atol=2e-15)
    zwo = w / (1. - (1. - w) * (g / (1. - g)) ** 2)
    assert (zwo >= 0.9999995).sum() > 1000 and (zwo < 0.9999995).sum() > 1000    # both branches exercised
    # adding method on those layers (plus a surface), top = index 0
    dbt = np.asfortranarray(np.exp(-tau / mu[None, None, :]))
    tdbt = np.ones((nlay + 1, ng, ncol), order="F")
    for k in range(nlay):
        tdbt[k + 1] = tdbt[k] * dbt[k]
    for o in out[:2]:
        o[nlay] = rng.uniform(0.03, 0.6, (ng, ncol))          # surface albedo: pref / prefd at klev+1
    out[2][nlay] = 0.
    out[3][nlay] = 0.
    pfd, pfu = np.zeros_like(tdbt), np.zeros_like(tdbt)
    assert L.oracle_vrtqdr_sw(ncol, nlay, *[p(o) for o in out], p(dbt), p(tdbt), p(pfd), p(pfu)) == 0
    rfd, rfu = _vrtqdr_np(out[0], out[1], out[2], out[3], dbt, tdbt)
    np.testing.assert_allclose(pfd, rfd, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(pfu, rfu, rtol=1e-12, atol=1e-15)
    assert (pfd[0] == 1.0).all()            # top boundary: unit downward flux, exactly


# ---- published-number sanity: AFGL mid-latitude summer, clear sky ------------------------------------------
# The reference ships no known-answer cases; the nearest thing to one is the ICRCCM / RRTMG validation
# literature for the AFGL standard atmospheres.  Line-by-line and RRTMG results for mid-latitude summer (MLS)
# clear sky cluster at OLR ~ 281-284 W/m2 and surface downward LW ~ 344-348 W/m2 (CO2 ~ 355-380 ppmv; today's
# CO2/CH4 lower the OLR by 1-2 W/m2), and the clear MLS atmosphere absorbs roughly a fifth of the incoming
# solar beam at 60 degrees.  The bounds below are 3 % wide: they do not pin bits, they catch a wrong band
# table, a unit slip in the column amounts or a broken continuum, which the property tests cannot see.
_MLS = np.array([  # z km, p hPa, T K, H2O ppmv  (AFGL-TR-86-0110, model 2)
    [0, 1013., 294.2, 18760.], [1, 902., 289.7, 13780.], [2, 802., 285.2, 9680.], [3, 710., 279.2, 5984.],
    [4, 628., 273.2, 3813.], [5, 554., 267.2, 2225.], [6, 487., 261.2, 1510.], [7, 426., 254.7, 1020.],
    [8, 372., 248.2, 646.], [9, 324., 241.7, 413.], [10, 281., 235.3, 247.], [11, 243., 228.8, 95.6],
    [12, 209., 222.3, 29.4], [13, 179., 215.8, 8.0], [14, 153., 215.7, 5.0], [15, 130., 215.7, 3.4],
    [16, 111., 215.7, 3.3], [17, 95., 215.7, 3.2], [18, 81.2, 216.8, 3.15], [19, 69.5, 217.9, 3.2],
    [20, 59.5, 219.2, 3.3], [21, 51., 220.4, 3.45], [22, 43.7, 221.6, 3.6], [23, 37.6, 222.8, 3.85],
    [24, 32.2, 223.9, 4.0], [25, 27.7, 225.1, 4.2], [27.5, 19.07, 228.45, 4.45], [30, 13.2, 233.7, 4.7],
    [32.5, 9.3, 239.0, 4.85], [35, 6.52, 245.2, 4.95], [37.5, 4.64, 251.3, 5.0], [40, 3.33, 257.5, 5.1],
    [42.5, 2.41, 263.7, 5.3], [45, 1.76, 269.9, 5.45], [47.5, 1.29, 275.2, 5.5], [50, 0.951, 275.7, 5.5]])


def _mls_state():
    from geosradiation_gridcomp_b200 import synthetic
    nlay = len(_MLS) - 1
    s = synthetic.make_columns(1, nlay=nlay, seed=1)
    f = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float64).reshape(1, -1))
    z, p, t, w = _MLS.T
    play = np.sqrt(p[:-1] * p[1:])                       # layer means (log-pressure midpoints)
    s.update(plev=f(p), play=f(play), tlev=f(t), tlay=f(0.5 * (t[:-1] + t[1:])), tsfc=f([t[0]]).reshape(1),
             h2ovmr=f(1e-6 * np.sqrt(w[:-1] * w[1:])),
             o3vmr=f(8e-6 * np.exp(-np.log(play / 10.0) ** 2 / 3.0) + 3e-8),
             zm=f(500.0 * (z[:-1] + z[1:])), emis=np.ones((1, 16), order="F"),
             cldf=f(np.zeros(nlay)), ciwp=f(np.zeros(nlay)), clwp=f(np.zeros(nlay)),
             tauaer_lw=np.zeros((1, nlay, 16), order="F"), tauaer_sw=np.zeros((1, nlay, 14), order="F"),
             coszen=np.array([0.5]), asdir=np.array([0.2]), asdif=np.array([0.2]), aldir=np.array([0.2]),
             aldif=np.array([0.2]), alat=np.array([0.7]), scon=1361.0, adjes=1.0)
    s["cloudLM"] = int(np.argmax(play < 700.0))
    s["cloudMH"] = int(np.argmax(play < 400.0))
    return s


def test_midlatitude_summer_clear_sky_against_published_ranges(oracle):
    s = _mls_state()
    lw = oracle.rrtmg_lw(s)
    assert lw["rc"] == 0
    olr, sfc_dn, sfc_up = lw["uflx"][0, -1], lw["dflx"][0, 0], lw["uflx"][0, 0]
    assert abs(sfc_up - 5.670373e-8 * 294.2 ** 4) < 0.5          # black surface at 294.2 K: 424.8 W/m2
    assert 272.0 < olr < 290.0, olr                               # literature 281-284 (older CO2/CH4)
    assert 338.0 < sfc_dn < 356.0, sfc_dn                         # literature 344-348
    np.testing.assert_array_equal(lw["uflx"], lw["uflxc"])       # no cloud: all-sky is clear-sky
    sw = oracle.rrtmg_sw(s, iaer=0, normFlx=0)
    assert sw["rc"] == 0
    toa_dn, toa_up = sw["swdflx"][0, -1], sw["swuflx"][0, -1]
    sfc_d, sfc_u = sw["swdflx"][0, 0], sw["swuflx"][0, 0]
    assert abs(toa_dn - 0.5 * 1361.0) < 0.05 * 0.5                # the 14 bands integrate to the solar constant
    absorbed = ((toa_dn - toa_up) - (sfc_d - sfc_u)) / toa_dn
    assert 0.16 < absorbed < 0.26, absorbed                       # clear MLS at 60 deg: about one fifth
    assert 0.68 < sfc_d / toa_dn < 0.80, sfc_d / toa_dn
    assert abs(sfc_u / sfc_d - 0.2) < 1e-9                        # Lambertian surface, albedo 0.2 in every band


# ---- an independent restatement of the LW radiative transfer ------------------------------------------------
# rtrnmc (LW/src/rrtmg_lw_rtrnmc.F90:22-390), the Planck-function interpolation and precipitable water of
# setcoef (LW/src/rrtmg_lw_setcoef.F90:204-395) and the transmittance tables of rrtmg_lw_ini
# (LW/src/rrtmg_lw_init.F90:96-114), written from the Fortran in numpy (g-points vectorised, layers looped)
# without reference to oracle/lw.c.  It is fed the oracle's own optical depths, Planck fractions and McICA
# cloud (taps), so what it pins is everything downstream of the gas optics.
def _lw_tables_np():
    ntbl, bpade, expeps = 10000, 1.0 / 0.278, 1.e-20
    tfn = np.arange(1, ntbl) / float(ntbl)
    tau = np.empty(ntbl + 1); ex = np.empty(ntbl + 1); tf = np.empty(ntbl + 1)
    tau[0], tau[ntbl], ex[0], ex[ntbl], tf[0], tf[ntbl] = 0.0, 1.e10, 1.0, expeps, 0.0, 1.0
    tau[1:ntbl] = bpade * tfn / (1. - tfn)
    ex[1:ntbl] = np.maximum(np.exp(-tau[1:ntbl]), expeps)
    t, e = tau[1:ntbl], ex[1:ntbl]
    with np.errstate(divide="ignore", invalid="ignore"):
        tf[1:ntbl] = np.where(t < 0.06, t / 6., 1. - 2. * ((1. / t) - (e / (1. - e))))
    return tau, ex, tf, bpade


def _lw_rt_np(s, taps, tab, dudTs=True):
    ncol, nlay = s["ncol"], s["nlay"]
    ngb = tab["lw.wvn.ngb"].astype(int)                                     # band of each g-point, 1-based
    totplnk, totplnkderiv = tab["lw.wvn.totplnk"], tab["lw.wvn.totplnkderiv"]
    wavenum1 = np.array([10., 350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600.])
    wavenum2 = np.array([350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600., 3250.])
    delwave, fluxfac, wtdiff = wavenum2 - wavenum1, np.pi * 2.e4, 0.5
    a0 = np.array([1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668] + [1.66] * 7)
    a1 = np.array([0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006] + [0.0] * 7)
    a2 = np.array([0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414] + [0.0] * 7)
    tau_tbl, exp_tbl, tfn_tbl, bpade = _lw_tables_np()
    tblint, amd, amw, avogad, grav = 10000.0, 28.9660, 18.0160, 6.02214199e+23, 9.8066
    ib = ngb - 1
    out = {k: np.zeros((ncol, nlay + 1)) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")}
    out["olrb"], out["dolrb_dTs"], out["pwvcm"] = np.zeros((16, ncol)), np.zeros((16, ncol)), np.zeros(ncol)

    def planck(t, tbl):                      # setcoef :278-347: table at 1 K steps from 160 K, linear in between
        i = int(t - 159.)
        i = min(max(i, 1), 180)
        fr = t - 159. - float(i)
        return tbl[i - 1] + fr * (tbl[i] - tbl[i - 1])           # all 16 bands

    for c in range(ncol):
        pz, h2o = s["plev"][c], s["h2ovmr"][c]
        amm = (1. - h2o) * amd + h2o * amw
        coldry = (pz[:-1] - pz[1:]) * 1.e3 * avogad / (1.e2 * grav * amm * (1. + h2o))
        amttl = wvttl = 0.
        for l in range(nlay):
            btemp = h2o[l] * coldry[l]
            amttl = amttl + coldry[l] + btemp
            wvttl = wvttl + btemp
        pwvcm = (amw * wvttl) / (amd * amttl) * (1.e3 * pz[0]) / (1.e2 * grav)
        out["pwvcm"][c] = pwvcm
        semiss = s["emis"][c]
        plankbnd = semiss * planck(s["tsfc"][c], totplnk)
        dplankbnd = semiss * planck(s["tsfc"][c], totplnkderiv)
        planklev = np.array([planck(t, totplnk) for t in s["tlev"][c]])       # (nlay+1, 16)
        planklay = np.array([planck(t, totplnk) for t in s["tlay"][c]])       # (nlay, 16)
        secdiff = np.clip(a0 + a1 * np.exp(a2 * pwvcm), 1.50, 1.80)
        for b in (1, 4, 10, 11, 12, 13, 14, 15, 16):
            secdiff[b - 1] = 1.66
        sec, sumfac = secdiff[ib], (wtdiff * delwave * fluxfac)[ib]
        taug, pfr, tauc = taps["taug"][c], taps["pfracs"][c], taps["taucmc"][c]     # (140, nlay)
        cloudy = taps["cldymc"][c].any(axis=0)                                       # (nlay,)
        agas, atot = np.zeros((nlay, 140)), np.zeros((nlay, 140))
        bbugas, bbutot = np.zeros((nlay, 140)), np.zeros((nlay, 140))
        radld, radclrd, diverge = np.zeros(140), np.zeros(140), False
        for l in range(nlay - 1, -1, -1):                                            # lev = nlay .. 1
            plfrac, blay = pfr[:, l], planklay[l][ib]
            dplankup, dplankdn = planklev[l + 1][ib] - blay, planklev[l][ib] - blay
            odepth = np.maximum(sec * taug[:, l], 0.)
            itgas = (tblint * (odepth / (bpade + odepth)) + 0.5).astype(int)
            agas[l] = 1. - exp_tbl[itgas]
            tfacgas = tfn_tbl[itgas]
            bbdgas = plfrac * (blay + tfacgas * dplankdn)
            bbugas[l] = plfrac * (blay + tfacgas * dplankup)
            cld = tauc[:, l] > 0.
            odtot = tau_tbl[itgas] + sec * tauc[:, l]
            ittot = (tblint * (odtot / (bpade + odtot)) + 0.5).astype(int)
            atot[l] = 1. - exp_tbl[ittot]
            tfactot = tfn_tbl[ittot]
            bbdtot = plfrac * (blay + tfactot * dplankdn)
            bbutot[l] = plfrac * (blay + tfactot * dplankup)
            radld = np.where(cld, radld + (bbdtot - radld) * atot[l], radld + (bbdgas - radld) * agas[l])
            diverge = diverge or bool(cloudy[l])
            radclrd = radclrd + (bbdgas - radclrd) * agas[l] if diverge else radld
            out["dflx"][c, l] = np.sum(sumfac * radld)
            out["dflxc"][c, l] = np.sum(sumfac * radclrd)
        rad0, drad0 = pfr[:, 0] * plankbnd[ib], pfr[:, 0] * dplankbnd[ib]
        reflect = 1. - semiss[ib]
        radlu, radclru = rad0 + reflect * radld, rad0 + reflect * radclrd
        dlu, dclru = drad0.copy(), drad0.copy()
        out["uflx"][c, 0], out["uflxc"][c, 0] = np.sum(sumfac * radlu), np.sum(sumfac * radclru)
        out["duflx_dTs"][c, 0], out["duflxc_dTs"][c, 0] = np.sum(sumfac * dlu), np.sum(sumfac * dclru)
        for l in range(nlay):
            cld = tauc[:, l] > 0.
            a = np.where(cld, atot[l], agas[l])
            radlu = radlu + (np.where(cld, bbutot[l], bbugas[l]) - radlu) * a
            dlu = dlu - dlu * a
            if diverge:
                radclru = radclru + (bbugas[l] - radclru) * agas[l]
                dclru = dclru - dclru * agas[l]
            else:
                radclru, dclru = radlu, dlu
            out["uflx"][c, l + 1], out["uflxc"][c, l + 1] = np.sum(sumfac * radlu), np.sum(sumfac * radclru)
            out["duflx_dTs"][c, l + 1], out["duflxc_dTs"][c, l + 1] = np.sum(sumfac * dlu), np.sum(sumfac * dclru)
        for b in range(16):
            if s["band_output"][b]:
                out["olrb"][b, c] = np.sum((sumfac * radlu)[ib == b])
                out["dolrb_dTs"][b, c] = np.sum((sumfac * dlu)[ib == b])
    return out


def test_lw_transfer_against_independent_numpy(oracle):
    from geosradiation_gridcomp_b200 import synthetic, tables
    tab = tables.load_tables()
    s = synthetic.make_columns(40, nlay=72, seed=77)
    s["tsfc"][:4] = [150.0, 159.5, 339.9, 345.0]                    # Planck table clamps at both ends
    s["tlev"][4, 0] = 341.0
    o = oracle.rrtmg_lw(s, taps=("taug", "pfracs", "taucmc", "cldymc", "pwvcm"))
    assert o["rc"] == 0
    tau_tbl, exp_tbl, tfn_tbl, _ = _lw_tables_np()
    for name, mine in (("tau_tbl", tau_tbl), ("exp_tbl", exp_tbl), ("tfn_tbl", tfn_tbl)):
        # tfn just above tau = 0.06 is a difference of nearly equal terms: one ulp of exp() shows as 4e-12
        rtol = {"tau_tbl": 4e-16, "exp_tbl": 1e-15, "tfn_tbl": 2e-11}[name]
        np.testing.assert_allclose(oracle.table("lw", name), mine, rtol=rtol, atol=1e-300)
    r = _lw_rt_np(s, o, tab)
    np.testing.assert_allclose(o["pwvcm"], r["pwvcm"], rtol=1e-13)
    cloudy_cols = (s["cldf"] > 0).any(axis=1)
    assert cloudy_cols.sum() >= 10 and (~cloudy_cols).sum() >= 10
    for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs"):
        scale = np.abs(r[k]).max(axis=1, keepdims=True)
        assert np.max(np.abs(o[k] - r[k]) / scale) < 2e-12, k
    bo = np.asarray(s["band_output"], bool)
    np.testing.assert_allclose(o["olrb"][bo], r["olrb"][bo], rtol=2e-12)
    np.testing.assert_allclose(o["dolrb_dTs"][bo], r["dolrb_dTs"][bo], rtol=2e-12)


# ---- an independent restatement of the LW gas optics for two structurally different bands ----------------------
# setcoef (LW/src/rrtmg_lw_setcoef.F90:204-575), the g-point reduction of band 1 (LW/src/rrtmg_lw_init.F90:114-145,
# cmbgb1 :356-437), taugb1 (one key species, N2 continuum minor; LW/src/rrtmg_lw_taumol.F90:200-296) and taugb3 (two
# key species with the three-way specparm interpolation, N2O minor with the abundance adjustment, Planck fractions
# interpolated in the binary parameter; :375-700), written from the Fortran in numpy on the ORIGINAL 16-g tables of
# the data blob, without reference to oracle/lw.c or oracle/lw_init.c.
def _lw_setcoef_np(s, c, tab):
    import math
    nlay = s["nlay"]
    preflog, tref, chi = tab["lw.ref.preflog"], tab["lw.ref.tref"], tab["lw.ref.chi_mls"]
    amd, amw, avogad, grav, stpfac = 28.9660, 18.0160, 6.02214199e+23, 9.8066, 296. / 1013.
    pz, L = s["plev"][c], []
    laytrop = 0
    for l in range(nlay):
        h2o, t, p = s["h2ovmr"][c, l], s["tlay"][c, l], s["play"][c, l]
        amm = (1. - h2o) * amd + h2o * amw
        coldry = (pz[l] - pz[l + 1]) * 1.e3 * avogad / (1.e2 * grav * amm * (1. + h2o))
        summol = s["co2vmr"][c, l] + s["o3vmr"][c, l] + s["n2ovmr"][c, l] + s["ch4vmr"][c, l] + s["o2vmr"][c, l]
        wbroad = coldry * (1. - summol)
        wv = h2o * coldry
        plog = math.log(p)
        jp = min(max(int(36. - 5 * (plog + 0.04)), 1), 58)
        fp = 5. * (preflog[jp - 1] - plog)
        jt = min(max(int(3. + (t - tref[jp - 1]) / 15.), 1), 4)
        ft = ((t - tref[jp - 1]) / 15.) - float(jt - 3)
        jt1 = min(max(int(3. + (t - tref[jp]) / 15.), 1), 4)
        ft1 = ((t - tref[jp]) / 15.) - float(jt1 - 3)
        water = wv / coldry
        scalefac = p * stpfac / t
        d = dict(jp=jp, jt=jt, jt1=jt1, coldry=coldry)
        d["forfac"] = scalefac / (1. + water)
        if plog > 4.56:
            laytrop += 1
            factor = (332. - t) / 36.
            d["indfor"] = min(2, max(1, int(factor)))
            d["forfrac"] = factor - float(d["indfor"])
            d["selffac"] = water * d["forfac"]
            factor = (t - 188.) / 7.2
            d["indself"] = min(9, max(1, int(factor) - 7))
            d["selffrac"] = factor - float(d["indself"] + 7)
        else:
            factor = (t - 188.) / 36.
            d["indfor"], d["forfrac"], d["selffac"], d["indself"], d["selffrac"] = 3, factor - 1., 0., 0, 0.
        d["scaleminorn2"] = (p / t) * (wbroad / (coldry + wv))
        factor = (t - 180.8) / 7.2
        d["indminor"] = min(18, max(1, int(factor)))
        d["minorfrac"] = factor - float(d["indminor"])
        d["rat_h2oco2"] = chi[0, jp - 1] / chi[1, jp - 1]
        d["rat_h2oco2_1"] = chi[0, jp] / chi[1, jp]
        d["colh2o"] = 1.e-20 * h2o * coldry
        d["colco2"] = 1.e-20 * s["co2vmr"][c, l] * coldry
        d["coln2o"] = 1.e-20 * s["n2ovmr"][c, l] * coldry
        if d["colco2"] == 0.: d["colco2"] = 1.e-32 * coldry
        if d["coln2o"] == 0.: d["coln2o"] = 1.e-32 * coldry
        d["colbrd"] = 1.e-20 * wbroad
        compfp = 1. - fp
        d["fac10"], d["fac00"], d["fac11"], d["fac01"] = compfp * ft, compfp * (1. - ft), fp * ft1, fp * (1. - ft1)
        d["selffac"] = d["colh2o"] * d["selffac"]
        d["forfac"] = d["colh2o"] * d["forfac"]
        L.append(d)
    return L, laytrop


def _lw_reduce_band1(tab):
    """rwgt of band 1 and the 16 -> 10 g-point combination (weighted for k, plain sums for the Planck fractions)."""
    wt, ngn = tab["lw.wvn.wt"], tab["lw.wvn.ngn"][:10]
    groups, i = [], 0
    for n in ngn:
        groups.append(list(range(i, i + n))); i += n
    rw = np.zeros(16)
    for g in groups:
        wsum = 0.
        for j in g: wsum = wsum + wt[j]
        for j in g: rw[j] = wt[j] / wsum

    def comb(a, weighted=True):                       # last axis = original g
        out = np.zeros(a.shape[:-1] + (10,))
        for k, g in enumerate(groups):
            for j in g:
                out[..., k] = out[..., k] + (a[..., j] * rw[j] if weighted else a[..., j])
        return out
    K = "lw.kg01."
    return dict(ka=comb(tab[K + "kao"]), kb=comb(tab[K + "kbo"]), selfref=comb(tab[K + "selfrefo"]),
                forref=comb(tab[K + "forrefo"]), ka_mn2=comb(tab[K + "kao_mn2"]), kb_mn2=comb(tab[K + "kbo_mn2"]),
                fracrefa=comb(tab[K + "fracrefao"], False), fracrefb=comb(tab[K + "fracrefbo"], False))


def _lw_taugb1_np(d, lower, p, R):
    lin = lambda t, i, f: t[i - 1] + f * (t[i] - t[i - 1])
    taufor = d["forfac"] * lin(R["forref"], d["indfor"], d["forfrac"])
    scalen2 = d["colbrd"] * d["scaleminorn2"]
    if lower:
        corradj = 1. - 0.15 * (250. - p) / 154.4 if p < 250. else 1.
        tauself = d["selffac"] * lin(R["selfref"], d["indself"], d["selffrac"])
        taun2 = scalen2 * lin(R["ka_mn2"], d["indminor"], d["minorfrac"])
        k, jp0, jp1 = R["ka"], d["jp"] - 1, d["jp"]                       # ka(jt, jp, ig), 0-based jp index
        major = (d["fac00"] * k[d["jt"] - 1, jp0] + d["fac10"] * k[d["jt"], jp0] +
                 d["fac01"] * k[d["jt1"] - 1, jp1] + d["fac11"] * k[d["jt1"], jp1])
        return corradj * (d["colh2o"] * major + tauself + taufor + taun2), R["fracrefa"]
    corradj = 1. - 0.15 * (p / 95.6)
    taun2 = scalen2 * lin(R["kb_mn2"], d["indminor"], d["minorfrac"])
    k, jp0, jp1 = R["kb"], d["jp"] - 13, d["jp"] - 12                      # kb(jt, 13:59, ig)
    major = (d["fac00"] * k[d["jt"] - 1, jp0] + d["fac10"] * k[d["jt"], jp0] +
             d["fac01"] * k[d["jt1"] - 1, jp1] + d["fac11"] * k[d["jt1"], jp1])
    return corradj * (d["colh2o"] * major + taufor + taun2), R["fracrefb"]


def _lw_taugb3_np(d, lower, tab):
    import math
    K, chi, oneminus = "lw.kg03.", tab["lw.ref.chi_mls"], 1. - 1.e-6
    lin = lambda t, i, f: t[i - 1] + f * (t[i] - t[i - 1])
    n = 8. if lower else 4.

    def binary(rat):
        speccomb = d["colh2o"] + rat * d["colco2"]
        specparm = min(d["colh2o"] / speccomb, oneminus)
        specmult = n * specparm
        return speccomb, specparm, 1 + int(specmult), math.fmod(specmult, 1.0)

    def major(k, jp0, jt, f0, f1, speccomb, specparm, js, fs):
        """k(js, jt, jp, ig): two temperatures jt, jt+1 at one reference pressure."""
        a = lambda dj, dt: k[js - 1 + dj, jt - 1 + dt, jp0]
        if lower and specparm < 0.125:
            p = fs - 1; p4 = p ** 4; fk0, fk1, fk2 = p4, 1 - p - 2.0 * p4, p + p4
            return speccomb * (fk0 * f0 * a(0, 0) + fk1 * f0 * a(1, 0) + fk2 * f0 * a(2, 0) +
                               fk0 * f1 * a(0, 1) + fk1 * f1 * a(1, 1) + fk2 * f1 * a(2, 1))
        if lower and specparm > 0.875:
            p = -fs; p4 = p ** 4; fk0, fk1, fk2 = p4, 1 - p - 2.0 * p4, p + p4
            return speccomb * (fk2 * f0 * a(-1, 0) + fk1 * f0 * a(0, 0) + fk0 * f0 * a(1, 0) +
                               fk2 * f1 * a(-1, 1) + fk1 * f1 * a(0, 1) + fk0 * f1 * a(1, 1))
        return speccomb * ((1. - fs) * f0 * a(0, 0) + fs * f0 * a(1, 0) + (1. - fs) * f1 * a(0, 1) + fs * f1 * a(1, 1))

    jp = d["jp"]
    k = tab[K + ("kao" if lower else "kbo")]
    off = 1 if lower else 13
    tau0 = major(k, jp - off, d["jt"], d["fac00"], d["fac10"], *binary(d["rat_h2oco2"]))
    tau1 = major(k, jp + 1 - off, d["jt1"], d["fac01"], d["fac11"], *binary(d["rat_h2oco2_1"]))
    refrat_m = chi[0, 2] / chi[1, 2] if lower else chi[0, 12] / chi[1, 12]
    refrat_planck = chi[0, 8] / chi[1, 8] if lower else chi[0, 12] / chi[1, 12]
    _, _, jmn2o, fmn2o = binary(refrat_m)
    _, _, jpl, fpl = binary(refrat_planck)
    chi_n2o = d["coln2o"] / d["coldry"]
    ratn2o = 1.e20 * chi_n2o / chi[3, jp]
    if ratn2o > 1.5:
        adjcoln2o = (0.5 + (ratn2o - 0.5) ** 0.65) * chi[3, jp] * d["coldry"] * 1.e-20
    else:
        adjcoln2o = d["coln2o"]
    km = tab[K + ("kao_mn2o" if lower else "kbo_mn2o")]
    im = d["indminor"]
    n2om1 = km[jmn2o - 1, im - 1] + fmn2o * (km[jmn2o, im - 1] - km[jmn2o - 1, im - 1])
    n2om2 = km[jmn2o - 1, im] + fmn2o * (km[jmn2o, im] - km[jmn2o - 1, im])
    absn2o = n2om1 + d["minorfrac"] * (n2om2 - n2om1)
    taufor = d["forfac"] * lin(tab[K + "forrefo"], d["indfor"], d["forfrac"])
    tauself = d["selffac"] * lin(tab[K + "selfrefo"], d["indself"], d["selffrac"]) if lower else 0.
    fr = tab[K + ("fracrefao" if lower else "fracrefbo")]
    pfrac = fr[:, jpl - 1] + fpl * (fr[:, jpl] - fr[:, jpl - 1])
    if lower:
        return tau0 + tau1 + tauself + taufor + adjcoln2o * absn2o, pfrac
    return tau0 + tau1 + taufor + adjcoln2o * absn2o, pfrac


def test_lw_gas_optics_bands_1_and_3_against_independent_numpy(oracle):
    from geosradiation_gridcomp_b200 import synthetic, tables
    tab = tables.load_tables()
    ncol = 24
    s = synthetic.make_columns(ncol, nlay=72, seed=91)
    s["n2ovmr"][:6] *= 2.5                                    # drives the N2O abundance adjustment (ratn2o > 1.5)
    s["h2ovmr"][6:9] *= 1e-3                                  # dry columns: specparm < 0.125 in the lower atmosphere
    s["co2vmr"][9:12] *= 1e-2                                 # CO2-poor columns: specparm > 0.875
    o = oracle.rrtmg_lw(s, taps=("jp", "jt", "jt1", "indfor", "indself", "indminor", "laytrop", "fac00", "fac01",
                                 "fac10", "fac11", "taug", "pfracs"))
    assert o["rc"] == 0
    R1 = _lw_reduce_band1(tab)
    np.testing.assert_allclose(oracle.table("lw", "absa", 1).reshape(65, 10, order="F"),
                               R1["ka"].reshape(65, 10, order="F"), rtol=1e-15)
    seen = set()
    for c in range(ncol):
        L, laytrop = _lw_setcoef_np(s, c, tab)
        assert laytrop == o["laytrop"][c]
        for l, d in enumerate(L):
            lower = l < laytrop
            for k in ("jp", "jt", "jt1", "indfor", "indminor"):
                assert d[k] == o[k][c, l], (k, c, l)
            if lower:
                assert d["indself"] == o["indself"][c, l]
            for k in ("fac00", "fac01", "fac10", "fac11"):
                assert abs(d[k] - o[k][c, l]) <= 1e-13, (k, c, l)
            aer = s["tauaer_lw"][c, l]
            t1, f1 = _lw_taugb1_np(d, lower, s["play"][c, l], R1)
            np.testing.assert_allclose(o["taug"][c, 0:10, l], t1 + aer[0], rtol=1e-11, err_msg=f"band 1 col {c} lay {l}")
            np.testing.assert_allclose(o["pfracs"][c, 0:10, l], f1, rtol=1e-15)
            t3, f3 = _lw_taugb3_np(d, lower, tab)
            np.testing.assert_allclose(o["taug"][c, 22:38, l], t3 + aer[2], rtol=1e-11, err_msg=f"band 3 col {c} lay {l}")
            np.testing.assert_allclose(o["pfracs"][c, 22:38, l], f3, rtol=1e-11)
            if lower:
                sp = min(d["colh2o"] / (d["colh2o"] + d["rat_h2oco2"] * d["colco2"]), 1. - 1.e-6)
                seen.add("lo" if sp < 0.125 else "hi" if sp > 0.875 else "mid")
            if 1.e20 * (d["coln2o"] / d["coldry"]) / tab["lw.ref.chi_mls"][3, d["jp"]] > 1.5:
                seen.add("adj")
    assert seen == {"lo", "mid", "hi", "adj"}, seen


# ---- an independent restatement of the McICA subcolumn generator -----------------------------------------------
# generate_stochastic_clouds (SH/cloud_subcol_gen.F90:93-330), the correlation lengths (:333-365), zcw_lookup
# (SH/cloud_condensate_inhomogeneity.F90, bilinear in the beta table) and clearCounts_threeBand (:610-760) in plain
# Python on top of kiss_python above, without reference to oracle/mcica.c.  The masks are integer work: bit-exact.
def _mcica_python(zmid, alat, doy, play, cldfrac, ciwp, clwp, nsub, xcw, so=(1, 2, 3, 4), cwp_tiny=1e-20, inhomo=True,
                  corr=(1.4315, 2.1219, 7., -25.584, 0.72192, 0.78996, 8.5, 40.404)):
    import math
    nlay, ncol = play.shape
    maximo = 2147483647 - 1

    def clength(am1, am2, am30, am4, lat):
        am3 = -4. * am30 / 365. * (doy - 272) if doy > 181 else 4. * am30 / 365. * (doy - 91)
        return (am1 + am2 * math.exp(-(lat * (180. / 3.14159265358979323846) - am3) ** 2 / am4 ** 2)) * 1.e3

    def zcw_lookup(cdf, sigma):
        r1 = cdf * (1000 - 1) + 1.
        i1 = max(1, min(int(r1), 999)); r1 = r1 - i1
        r2 = 40. * sigma - 3.
        i2 = max(1, min(int(r2), 139)); r2 = r2 - i2
        return ((1.0 - r1) * (1.0 - r2) * xcw[i1 - 1, i2 - 1] + (1.0 - r1) * r2 * xcw[i1 - 1, i2] +
                r1 * (1.0 - r2) * xcw[i1, i2 - 1] + r1 * r2 * xcw[i1, i2])

    mask = np.zeros((nlay, nsub, ncol), dtype=np.uint8)
    ci, cw = np.zeros((nlay, nsub, ncol)), np.zeros((nlay, nsub, ncol))
    for c in range(ncol):
        adl = clength(*corr[:4], alat[c])
        rdl = clength(*corr[4:], alat[c])
        alpha = [0.] + [math.exp(-abs(zmid[l, c] - zmid[l - 1, c]) / adl) for l in range(1, nlay)]
        rcorr = [0.] + [math.exp(-abs(zmid[l, c] - zmid[l - 1, c]) / rdl) for l in range(1, nlay)]
        sigma = [0.5 if f > 0.99 else 0.71 if f > 0.9 else 1.0 for f in cldfrac[:, c]]
        assert play[0, 0] > play[nlay - 1, 0]                         # surface at layer 1
        pseed = [play[k, c] * 100. for k in range(4)]
        seeds = [int((pseed[so[k] - 1] - int(pseed[so[k] - 1])) * maximo + 1) for k in range(4)]
        ndraw = nsub * nlay * (4 if inhomo else 2)
        u = list(kiss_python(seeds, ndraw))
        pos = 0
        for j in range(nsub):
            cdf1, cdf2 = u[pos:pos + 2 * nlay:2], u[pos + 1:pos + 2 * nlay:2]; pos += 2 * nlay
            for l in range(1, nlay):
                if cdf2[l] < alpha[l]: cdf1[l] = cdf1[l - 1]
            if inhomo:
                cdf2, cdf3 = u[pos:pos + 2 * nlay:2], u[pos + 1:pos + 2 * nlay:2]; pos += 2 * nlay
                for l in range(1, nlay):
                    if cdf2[l] < rcorr[l]: cdf3[l] = cdf3[l - 1]
            for l in range(nlay):
                if cdf1[l] >= 1. - cldfrac[l, c]:
                    z = zcw_lookup(cdf3[l], sigma[l]) if inhomo else 1.
                    i, w = ciwp[l, c] * z if inhomo else ciwp[l, c], clwp[l, c] * z if inhomo else clwp[l, c]
                    if i <= cwp_tiny: i = 0.
                    if w <= cwp_tiny: w = 0.
                    ci[l, j, c], cw[l, j, c] = i, w
                    mask[l, j, c] = 0 if (i == 0. and w == 0.) else 1
    return mask, ci, cw


def _clear_counts_python(mask, cloudLM, cloudMH):
    nlay, nsub, ncol = mask.shape
    out = np.zeros((4, ncol), dtype=np.int32)
    assert cloudLM < cloudMH
    for c in range(ncol):
        m = mask[:, :, c].astype(bool)
        out[0, c] = (~m.any(axis=0)).sum()
        out[3, c] = (~m[:cloudLM].any(axis=0)).sum()
        out[2, c] = (~m[cloudLM:cloudMH].any(axis=0)).sum()
        out[1, c] = (~m[cloudMH:].any(axis=0)).sum()
    return out


@pytest.mark.parametrize("inhomo", [True, False])
def test_mcica_generator_against_independent_python(oracle, inhomo):
    from geosradiation_gridcomp_b200 import tables
    xcw = tables.load_tables()["mcica.xcw_beta"]
    s = make_columns(40, 72, seed=5150)
    cols = np.flatnonzero((s["cldf"] > 0).any(axis=1))[:5]
    assert len(cols) == 5
    T = lambda k: np.ascontiguousarray(s[k][cols].T)                   # (nlay, ncol) partition layout
    zmid, play, cld, ciwp, clwp = T("zm"), T("play"), T("cldf"), T("ciwp"), T("clwp")
    cld[10:14, 0] = [0.95, 0.995, 1.0, 0.91]                           # all three sigma_qcw classes
    clwp[10:14, 0] = [30., 1e-21, 12., 4.]                             # a negligible water path un-clouds the cell
    ciwp[10:14, 0] = 0.
    alat, doy, nsub = s["alat"][cols], 200, 140
    oracle.set_mcica(1 if inhomo else 0)
    try:
        m, ci, cw = oracle.generate_stochastic_clouds(zmid, alat, doy, play, cld, ciwp, clwp, nsub, seed_order=(2, 1, 4, 3))
    finally:
        oracle.set_mcica(1)
    pm, pci, pcw = _mcica_python(zmid, alat, doy, play, cld, ciwp, clwp, nsub, xcw, so=(2, 1, 4, 3), inhomo=inhomo)
    np.testing.assert_array_equal(m, pm)
    assert 0.02 < pm.mean() < 0.5 and pm[11, :, 0].sum() == 0
    np.testing.assert_allclose(ci, pci, rtol=1e-14, atol=0)
    np.testing.assert_allclose(cw, pcw, rtol=1e-14, atol=0)
    LM, MH = int(s["cloudLM"]), int(s["cloudMH"])
    np.testing.assert_array_equal(oracle.clear_counts(m, LM, MH), _clear_counts_python(pm, LM, MH))


# ---- the clear-sky SW chain on the independent two-stream/adding restatement -------------------------------------
# spcvmc_sw's clear-sky orchestration (SW/src/rrtmg_sw_spcvmc.F90:391-492: surface albedo per band, aerosol mixing
# and delta scaling, direct-beam transmittances, flux accumulation with adjflux * ssi * mu0) and the driver's
# albedo / adjflux / cossza set-up (SW/src/rrtmg_sw_rad.F90:957, :1046, :1121-1127, :1230-1248, :1365), in numpy on
# _reftra_np / _vrtqdr_np above; fed the oracle's gas and Rayleigh optical depths and solar source (taps).
@pytest.mark.parametrize("isolvar", [-1, 0])
def test_sw_clear_sky_chain_against_independent_numpy(oracle, isolvar):
    from geosradiation_gridcomp_b200 import tables
    tab = tables.load_tables()
    ncol, nlay = 16, 72
    s = make_columns(ncol, nlay, seed=4242)
    s["coszen"][0] = 1e-12                                        # clamped to zepzen = 1e-10
    s["adjes"] = 1.0173
    o = oracle.rrtmg_sw(s, isolvar=isolvar, normFlx=0, taps=("taug", "pfracs", "ssi"))
    assert o["rc"] == 0
    ibm = tab["sw.wvn.ngb"].astype(int) - 16                      # 0-based band of each g-point
    adjflux = s["adjes"] * (s["scon"] / 1.36822e+03 if isolvar < 0 else 1.0)
    for c in range(ncol):
        mu = max(1.e-10, s["coszen"][c])
        albp = np.where((ibm <= 7) | (ibm == 13), s["aldir"][c], np.where(ibm >= 9, s["asdir"][c], (s["asdir"][c] + s["aldir"][c]) / 2.))
        albd = np.where((ibm <= 7) | (ibm == 13), s["aldif"][c], np.where(ibm >= 9, s["asdif"][c], (s["asdif"][c] + s["aldif"][c]) / 2.))
        top_down = lambda a: a[::-1]                               # jk = nlay+1-ikl
        taug, taur = top_down(o["taug"][c].T), top_down(o["pfracs"][c].T)          # (nlay, 112)
        taua, omga, asya = (top_down(s[k][c])[:, ibm] for k in ("tauaer_sw", "ssaaer", "asmaer"))
        ztauo = taur + taug + taua
        zomco = taur + taua * omga
        zgco = (asya * omga * taua) / zomco
        zomco = zomco / ztauo
        zf = zgco ** 2
        zwf = zomco * zf
        ztauo = (1. - zwf) * ztauo
        zomco = (zomco - zwf) / (1. - zwf)
        zgco = (zgco - zf) / (1. - zf)
        ref, refd, tra, trad = (np.vstack([a, b[None, :]]) for a, b in
                                zip(_reftra_np(ztauo, zomco, zgco, mu), (albp, albd, 0. * albp, 0. * albp)))
        dbt = np.exp(-ztauo / mu)
        tdbt = np.ones((nlay + 1, 112))
        for k in range(nlay):
            tdbt[k + 1] = dbt[k] * tdbt[k]
        fd, fu = _vrtqdr_np(ref, refd, tra, trad, dbt, tdbt)
        zinc = adjflux * o["ssi"][c] * mu
        dn, up = (zinc * fd).sum(axis=1)[::-1], (zinc * fu).sum(axis=1)[::-1]       # level 1 = surface
        scale = dn.max()
        assert np.max(np.abs(o["swdflxc"][c] - dn)) / scale < 5e-12, c
        assert np.max(np.abs(o["swuflxc"][c] - up)) / scale < 5e-12, c


# ---- an independent restatement of the LW cloud optics -----------------------------------------------------------
# cldprmc (LW/src/rrtmg_lw_cldprmc.F90:12-385): every ice parameterisation (iceflag 0-4) and the Hu & Stamnes liquid
# table, vectorised in numpy from the McICA water paths the oracle taps, without reference to oracle/lw.c.
def _lw_cldprmc_np(s, o, tab, iceflag):
    ngb = tab["lw.wvn.ngb"].astype(int)                         # 1-based band of each g-point
    ice1b = np.array([1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5])
    cldy = o["cldymc"].astype(bool)                             # [icol][ig][ilay]
    ciwp, clwp = o["ciwpmc"], o["clwpmc"]
    rei, rel = s["rei"][:, None, :], s["rel"][:, None, :]       # (ncol, 1, nlay)

    def table(name, factor, nmax):                              # linear interpolation with the end-interval rule
        t = tab["lw.cld." + name]
        idx = factor.astype(int)                                # int(): truncation, radii are positive
        assert idx.min() >= 0 and idx.max() <= nmax
        idx = np.where(idx == nmax, nmax - 1, np.where(idx == 0, 1, idx))
        fint = factor - idx
        lo, hi = t[idx - 1, ngb[None, :, None] - 1], t[idx, ngb[None, :, None] - 1]
        return lo + fint * (hi - lo)

    if iceflag == 0:
        a = tab["lw.cld.absice0"]
        abscoice = (a[0] + a[1] / rei) * np.ones((1, 140, 1))
    elif iceflag == 1:
        a = tab["lw.cld.absice1"]
        ib = ice1b[ngb - 1][None, :, None] - 1
        abscoice = a[0, ib] + a[1, ib] / rei
    elif iceflag == 2:
        abscoice = table("absice2", (rei - 2.) / 3. + 0 * ciwp, 43)
    elif iceflag == 3:
        abscoice = table("absice3", (rei - 2.) / 3. + 0 * ciwp, 46)
    else:
        abscoice = table("absice4", rei + 0 * ciwp, 200)
    tau = np.where(cldy & (ciwp > 0.), ciwp * abscoice, 0.)
    abscoliq = table("absliq1", rel - 1.5 + 0 * clwp, 58)
    return np.where(cldy & (clwp > 0.), tau + clwp * abscoliq, tau)


@pytest.mark.parametrize("iceflag", [0, 1, 2, 3, 4])
def test_lw_cloud_optics_against_independent_numpy(oracle, iceflag):
    from geosradiation_gridcomp_b200 import tables
    tab = tables.load_tables()
    s = make_columns(48, 72, seed=606)
    if iceflag == 2:
        s["rei"] = np.asfortranarray(np.minimum(s["rei"], 131.0))          # absice2 ends at 131 um
    if iceflag == 4:
        s["rei"] = np.asfortranarray(1.0 + (s["rei"] - 15.0) * 1.89)       # 1 .. 199.5: both end intervals of absice4
    o = oracle.rrtmg_lw(s, iceflg=iceflag, taps=("cldymc", "ciwpmc", "clwpmc", "taucmc"))
    assert o["rc"] == 0
    mine = _lw_cldprmc_np(s, o, tab, iceflag)
    assert (mine > 0).sum() > 5000
    np.testing.assert_allclose(o["taucmc"], mine, rtol=1e-14, atol=0)


# ---- the all-sky SW chain: cloud optics, cloudy-cell mixing, surface components, PAR optical thickness -------------
# cldprmc_sw (SW/src/rrtmg_sw_cldprmc.F90:62-418, iceflag 1-4 and the liquid table with its delta scaling), the
# cloudy pass of spcvmc_sw (SW/src/rrtmg_sw_spcvmc.F90:501-676), its PAR-weighted in-cloud optical thickness
# (:748-1108 without the SOLAR_RADVAL blocks) and the driver's output mapping and normalisation
# (SW/src/rrtmg_sw_rad.F90:1604-1660, :1769-1798), in numpy from the McICA cloud the oracle taps.
def _sw_cldprmc_np(s, o, tab, iceflag):
    ngb0 = tab["sw.wvn.ngb"].astype(int) - 16                    # 0-based band of each g-point
    icxa = tab["sw.wvn.icxa"].astype(int)
    cldy, ciwp, clwp = o["cldymc"].astype(bool), o["ciwpmc"], o["clwpmc"]      # [icol][ig][ilay]
    B = ngb0[None, :, None]
    rei, rel = s["rei"][:, None, :] + 0 * ciwp, s["rel"][:, None, :] + 0 * ciwp
    T = lambda n: tab["sw.cld." + n]
    lin = lambda t, idx, fint: t[idx - 1, B] + fint * (t[idx, B] - t[idx - 1, B])
    epsg, cldmin = 1.e-06, 1.e-20
    if iceflag == 1:
        ib = icxa[ngb0][None, :, None] - 1
        ext = T("abari")[ib] + T("bbari")[ib] / rei
        ssa = 1. - T("cbari")[ib] - T("dbari")[ib] * rei
        g = np.minimum(T("ebari")[ib] + T("fbari")[ib] * rei, 1. - epsg)
        forw = g * g
    else:
        factor = rei if iceflag == 4 else (rei - 2.) / 3.
        idx = factor.astype(int)
        if iceflag == 2: idx = np.where(idx == 43, 42, idx)
        if iceflag == 3: idx = np.where(idx == 46, 45, idx)
        fint = factor - idx
        sfx = str(iceflag)
        ext, ssa, g = lin(T("extice" + sfx), idx, fint), lin(T("ssaice" + sfx), idx, fint), lin(T("asyice" + sfx), idx, fint)
        forw = np.minimum(lin(T("fdlice3"), idx, fint) + 0.5 / ssa, g) if iceflag == 3 else g * g
    noice = ciwp == 0.
    ext, ssa, g, forw = (np.where(noice, 0., a) for a in (ext, ssa, g, forw))
    idx = (rel - 1.5).astype(int)
    idx = np.where(idx == 0, 1, np.where(idx == 58, 57, idx))
    fint = rel - 1.5 - idx
    extl, ssal, gl = lin(T("extliq1"), idx, fint), lin(T("ssaliq1"), idx, fint), lin(T("asyliq1"), idx, fint)
    ssal = np.where((fint < 0.) & (ssal > 1.), T("ssaliq1")[idx - 1, B], ssal)
    noliq = clwp == 0.
    extl, ssal, gl = (np.where(noliq, 0., a) for a in (extl, ssal, gl))
    forwl = gl * gl
    with np.errstate(invalid="ignore", divide="ignore"):
        tauliqorig, tauiceorig = clwp * extl, ciwp * ext
        ssaliq = ssal * (1. - forwl) / (1. - forwl * ssal)
        ssaice = ssa * (1. - forw) / (1. - forw * ssa)
        tauliq, tauice = (1. - forwl * ssal) * tauliqorig, (1. - forw * ssa) * tauiceorig
        scatliq, scatice = ssaliq * tauliq, ssaice * tauice
        tauc = tauliq + tauice
        tauc = np.where(tauc == 0., cldmin, tauc)
        scatice = np.where(scatice == 0., cldmin, scatice)
        ssac = (scatliq + scatice) / tauc
        if iceflag == 3:
            asmc = (1. / (scatliq + scatice)) * (scatliq * (gl - forwl) / (1. - forwl) + scatice * ((g - forw) / (1. - forw)))
        else:
            asmc = (scatliq * (gl - forwl) / (1. - forwl) + scatice * (g - forw) / (1. - forw)) / (scatliq + scatice)
    z = lambda a, fill: np.where(cldy, a, fill)
    return z(tauliqorig + tauiceorig, 0.), z(tauc, 0.), z(ssac, 1.), z(asmc, 0.)


@pytest.mark.parametrize("iceflag", [1, 2, 3, 4])
def test_sw_all_sky_chain_against_independent_numpy(oracle, iceflag):
    from geosradiation_gridcomp_b200 import tables
    tab = tables.load_tables()
    ncol, nlay = 24, 72
    s = make_columns(ncol, nlay, seed=977 + iceflag)
    if iceflag == 2:
        s["rei"] = np.asfortranarray(np.minimum(s["rei"], 131.0))
    o = oracle.rrtmg_sw(s, iceflg=iceflag, normFlx=1, do_drfband=True,
                        taps=("taug", "pfracs", "ssi", "cldymc", "ciwpmc", "clwpmc", "taucmc"))
    assert o["rc"] == 0
    taor, tauc, ssac, asmc = _sw_cldprmc_np(s, o, tab, iceflag)
    np.testing.assert_allclose(o["taucmc"], tauc, rtol=1e-13, atol=0)
    ibm = tab["sw.wvn.ngb"].astype(int) - 16
    LM, MH = int(s["cloudLM"]), int(s["cloudMH"])
    worst = 0.
    for c in range(ncol):
        mu = max(1.e-10, s["coszen"][c])
        nir = (ibm <= 7) | (ibm == 13)
        albp = np.where(nir, s["aldir"][c], np.where(ibm >= 9, s["asdir"][c], (s["asdir"][c] + s["aldir"][c]) / 2.))
        albd = np.where(nir, s["aldif"][c], np.where(ibm >= 9, s["asdif"][c], (s["asdif"][c] + s["aldif"][c]) / 2.))
        td = lambda a: a[::-1]
        taug, taur = td(o["taug"][c].T), td(o["pfracs"][c].T)
        taua, omga, asya = (td(s[k][c])[:, ibm] for k in ("tauaer_sw", "ssaaer", "asmaer"))
        ztauo = taur + taug + taua
        zomco = taur + taua * omga
        zgco = (asya * omga * taua) / zomco
        zomco = zomco / ztauo
        zf = zgco ** 2
        zwf = zomco * zf
        ztauo = (1. - zwf) * ztauo
        zomco = (zomco - zwf) / (1. - zwf)
        zgco = (zgco - zf) / (1. - zf)
        cld = td(o["cldymc"][c].T.astype(bool))
        ptau, pomg, pasy = td(tauc[c].T), td(ssac[c].T), td(asmc[c].T)
        g2 = ztauo * zomco * zgco + ptau * pomg * pasy
        o2 = ztauo * zomco + ptau * pomg
        t2 = ztauo + ptau
        g2 = g2 / o2
        o2 = o2 / t2
        ztauo, zomco, zgco = np.where(cld, t2, ztauo), np.where(cld, o2, zomco), np.where(cld, g2, zgco)
        ref, refd, tra, trad = (np.vstack([a, b[None, :]]) for a, b in
                                zip(_reftra_np(ztauo, zomco, zgco, mu), (albp, albd, 0. * albp, 0. * albp)))
        dbt = np.exp(-ztauo / mu)
        tdbt = np.ones((nlay + 1, 112))
        for k in range(nlay):
            tdbt[k + 1] = dbt[k] * tdbt[k]
        fd, fu = _vrtqdr_np(ref, refd, tra, trad, dbt, tdbt)
        zinc = s["adjes"] * o["ssi"][c] * mu
        dn, up = (zinc * fd).sum(axis=1)[::-1], (zinc * fu).sum(axis=1)[::-1]
        top = max(dn[-1], 1e-7)
        sel = lambda m, a: float((zinc * a)[m].sum())
        dirs, tots = tdbt[nlay], fd[nlay]
        nirr = sel((ibm <= 7) | (ibm == 13), dirs) + 0.5 * sel(ibm == 8, dirs)
        nirt = sel((ibm <= 7) | (ibm == 13), tots) + 0.5 * sel(ibm == 8, tots)
        parr = sel((ibm == 9) | (ibm == 10), dirs) + 0.5 * sel(ibm == 8, dirs)
        part = sel((ibm == 9) | (ibm == 10), tots) + 0.5 * sel(ibm == 8, tots)
        uvrr, uvrt = sel((ibm == 11) | (ibm == 12), dirs), sel((ibm == 11) | (ibm == 12), tots)
        mine = dict(swdflx=dn / top, swuflx=up / top, nirr=nirr / top, nirf=(nirt - nirr) / top, parr=parr / top,
                    parf=(part - parr) / top, uvrr=uvrr / top, uvrf=(uvrt - uvrr) / top,
                    fswband=np.array([sel(ibm == b, fd[nlay] - fu[nlay]) for b in range(14)]) / top,
                    drband=np.array([sel(ibm == b, dirs) for b in range(14)]) / top,
                    dfband=np.array([sel(ibm == b, tots) - sel(ibm == b, dirs) for b in range(14)]) / top)
        # PAR-weighted in-cloud optical thickness per super-layer (not normalised)
        wgt = np.where((ibm == 9) | (ibm == 10), 1.0, np.where(ibm == 8, 0.5, 0.0)) * (s["adjes"] * o["ssi"][c])
        t = taor[c]                                                 # (112, nlay), layer 1 = surface
        lo, mi, hi = t[:, :LM].sum(axis=1), t[:, LM:MH].sum(axis=1), t[:, MH:].sum(axis=1)
        tot = lo + mi + hi
        if cld.any():
            for key, v in (("l", lo), ("m", mi), ("h", hi), ("t", tot)):
                mine["cotd" + key + "p"] = float(wgt[v > 0.].sum())
                mine["cotn" + key + "p"] = float((wgt * v)[v > 0.].sum())
        for k, v in mine.items():
            got = o[k][c]
            err = np.max(np.abs(got - v)) / max(np.max(np.abs(v)), 1e-30)
            worst = max(worst, err)
            assert err < 2e-11, (k, c, err)
    assert (o["cotdtp"] > 0).sum() >= 8 and (o["cotdtp"] == 0).sum() >= 4     # cloudy and clear columns both seen


# ---- an independent restatement of the SW gas optics of bands 17 and 24 ---------------------------------------------------
# The driver's column amounts (SW/src/rrtmg_sw_rad.F90:1368-1383), setcoef_sw (SW/src/rrtmg_sw_setcoef.F90:44-140),
# the 16 -> 12 g-point reduction of band 17 (SW/src/rrtmg_sw_init.F90:125-150, cmbgb17 :570-650) and taumol17
# (two key species, self and foreign continuum, Rayleigh, solar source interpolated in the binary parameter at the
# layer where jp crosses layreffr; SW/src/rrtmg_sw_taumol.F90:352-528), in numpy on the blob's original 16-g tables.
def _sw_comb(tab, band):
    """rwgt of a band (SW/src/rrtmg_sw_init.F90:125-150) and its 16 -> ngc combination along one axis."""
    ngs = [0] + list(tab["sw.wvn.ngs"])
    wt, ngn = tab["sw.wvn.wt"], tab["sw.wvn.ngn"][ngs[band - 16]:ngs[band - 15]]
    groups, i = [], 0
    for n in ngn:
        groups.append(list(range(i, i + n))); i += n
    assert i == 16
    rw = np.zeros(16)
    for g in groups:
        wsum = 0.
        for j in g: wsum = wsum + wt[j]
        for j in g: rw[j] = wt[j] / wsum

    def comb(a, axis, weighted):
        a = np.moveaxis(a, axis, -1)
        out = np.zeros(a.shape[:-1] + (len(groups),))
        for k, g in enumerate(groups):
            for j in g:
                out[..., k] = out[..., k] + (a[..., j] * rw[j] if weighted else a[..., j])
        return out
    return comb, slice(ngs[band - 16], ngs[band - 15])


def _sw_setcoef_np(s, c, tab):
    """Column amounts of the driver and setcoef_sw, layer by layer; returns (list of dicts, laytrop)."""
    import math
    preflog, tref = tab["sw.ref.preflog"], tab["sw.ref.tref"]
    amd, amw, avogad, grav, stpfac = 28.9660, 18.0160, 6.02214199e+23, 9.8066, 296. / 1013.
    L, laytrop = [], 0
    for l in range(s["nlay"]):
        h2o, t, p = s["h2ovmr"][c, l], s["tlay"][c, l], s["play"][c, l]
        coldry = (s["plev"][c, l] - s["plev"][c, l + 1]) * 1.e3 * avogad / (1.e2 * grav * ((1. - h2o) * amd + h2o * amw) * (1. + h2o))
        d = {k: coldry * s[k + "vmr"][c, l] for k in ("h2o", "co2", "o3", "ch4", "o2")}
        plog = math.log(p)
        d["lower"] = plog > 4.56
        if plog >= 4.56: laytrop += 1
        jp = min(max(int(36. - 5 * (plog + 0.04)), 1), 58)
        fp = 5. * (preflog[jp - 1] - plog)
        jt = min(max(int(3. + (t - tref[jp - 1]) / 15.), 1), 4)
        ft = ((t - tref[jp - 1]) / 15.) - float(jt - 3)
        jt1 = min(max(int(3. + (t - tref[jp]) / 15.), 1), 4)
        ft1 = ((t - tref[jp]) / 15.) - float(jt1 - 3)
        water = d["h2o"] / coldry
        d["forfac"] = p * stpfac / t / (1. + water)
        if d["lower"]:
            factor = (332. - t) / 36.
            d["indfor"] = min(2, max(1, int(factor))); d["forfrac"] = factor - float(d["indfor"])
            d["selffac"] = water * d["forfac"]
            factor = (t - 188.) / 7.2
            d["indself"] = min(9, max(1, int(factor) - 7)); d["selffrac"] = factor - float(d["indself"] + 7)
        else:
            d["indfor"], d["forfrac"] = 3, (t - 188.) / 36. - 1.
        for k in ("h2o", "co2", "o3", "ch4", "o2"):
            d[k] = 1.e-20 * d[k]
        d["colmol"] = 1.e-20 * coldry + d["h2o"]
        for k in ("co2", "ch4", "o2"):
            if d[k] == 0.: d[k] = 1.e-32 * coldry
        compfp = 1. - fp
        d.update(jp=jp, jt=jt, jt1=jt1, fac10=compfp * ft, fac00=compfp * (1. - ft), fac11=fp * ft1, fac01=fp * (1. - ft1))
        L.append(d)
    return L, laytrop


def _sw_binary(d, other, strrat, n):
    import math
    speccomb = d["h2o"] + strrat * d[other]
    specmult = n * min(d["h2o"] / speccomb, 1. - 1.e-06)
    return speccomb, 1 + int(specmult), math.fmod(specmult, 1.)


def _sw_major2(k, off, d, speccomb, js, fs):
    """speccomb * the eight-point interpolation in (binary parameter, temperature, pressure); k(js, jt, jp, ig)."""
    a = lambda dj, jtt, jpp: k[js - 1 + dj, jtt - 1, jpp - off]
    jp, jt, jt1 = d["jp"], d["jt"], d["jt1"]
    return speccomb * ((1. - fs) * d["fac00"] * a(0, jt, jp) + fs * d["fac00"] * a(1, jt, jp) +
                       (1. - fs) * d["fac10"] * a(0, jt + 1, jp) + fs * d["fac10"] * a(1, jt + 1, jp) +
                       (1. - fs) * d["fac01"] * a(0, jt1, jp + 1) + fs * d["fac01"] * a(1, jt1, jp + 1) +
                       (1. - fs) * d["fac11"] * a(0, jt1 + 1, jp + 1) + fs * d["fac11"] * a(1, jt1 + 1, jp + 1))


def _sw_source(src, js, fs, isolvar, scon):
    f = lambda t: t[js - 1] + fs * (t[js] - t[js - 1])
    if isolvar < 0:
        return f(src["sfluxref"])
    # isolvar = 0: the three NRLSSI2 terms share one scaling, scon over the mean-cycle integrals
    # (SW/src/rrtmg_sw_rad.F90:1050-1055, SW/src/NRLSSI2.F90:47-49)
    svar = scon / (0.996047 + -0.511590 + 1360.37)
    return svar * f(src["facbrght"]) + svar * f(src["snsptdrk"]) + svar * f(src["irradnce"])


_lin = lambda t, i, f: t[i - 1] + f * (t[i] - t[i - 1])


def _sw_band17_np(s, c, tab, isolvar):
    K = "sw.kg17."
    comb, _ = _sw_comb(tab, 17)
    ka, kb = comb(tab[K + "kao"], 3, True), comb(tab[K + "kbo"], 3, True)           # (js, jt, jp, g)
    selfref, forref = comb(tab[K + "selfrefo"], 1, True), comb(tab[K + "forrefo"], 1, True)
    src = {n: comb(tab[K + n + "o"], 0, False) for n in ("sfluxref", "irradnce", "facbrght", "snsptdrk")}   # (js, g)
    rayl, strrat, layreffr = float(tab[K + "rayl"][0]), 0.364641, 30
    L, laytrop = _sw_setcoef_np(s, c, tab)
    nlay = len(L)
    taug, taur = np.zeros((nlay, 12)), np.zeros((nlay, 12))
    for l, d in enumerate(L):
        speccomb, js, fs = _sw_binary(d, "co2", strrat, 8. if d["lower"] else 4.)
        major = _sw_major2(ka if d["lower"] else kb, 1 if d["lower"] else 13, d, speccomb, js, fs)
        cont = d["forfac"] * _lin(forref, d["indfor"], d["forfrac"])
        if d["lower"]:
            cont = d["selffac"] * _lin(selfref, d["indself"], d["selffrac"]) + cont
        taug[l] = major + d["h2o"] * cont
        taur[l] = d["colmol"] * rayl
    laysolfr, ssi = nlay, None
    for lay in range(laytrop + 1, nlay + 1):                     # 1-based layers above the tropopause
        if L[lay - 2]["jp"] < layreffr and L[lay - 1]["jp"] >= layreffr: laysolfr = lay
        if lay == laysolfr:
            _, js, fs = _sw_binary(L[lay - 1], "co2", strrat, 4.)
            ssi = _sw_source(src, js, fs, isolvar, s["scon"])
            break
    return taug, taur, ssi, laytrop


def _sw_band24_np(s, c, tab, isolvar):
    """taumol24 (SW/src/rrtmg_sw_taumol.F90:1364-1504) and cmbgb24 (SW/src/rrtmg_sw_init.F90:1214-1322): H2O/O2 below
    the tropopause with O3 and the water continuum, O2 and O3 above; the Rayleigh coefficient depends on the g-point
    and, below, on the binary parameter; the source layer search runs over the LOWER layers."""
    K = "sw.kg24."
    comb, _ = _sw_comb(tab, 24)
    ka, kb = comb(tab[K + "kao"], 3, True), comb(tab[K + "kbo"], 2, True)           # (js, jt, jp, g), (jt, jp, g)
    selfref, forref = comb(tab[K + "selfrefo"], 1, True), comb(tab[K + "forrefo"], 1, True)
    src = {n: comb(tab[K + n + "o"], 0, False) for n in ("sfluxref", "irradnce", "facbrght", "snsptdrk")}
    rayla, raylb = comb(tab[K + "raylao"], 0, True), comb(tab[K + "raylbo"], 0, True)   # (js, g), (g)
    abso3a, abso3b = comb(tab[K + "abso3ao"], 0, True), comb(tab[K + "abso3bo"], 0, True)
    strrat, layreffr = 0.124692, 1
    L, laytrop = _sw_setcoef_np(s, c, tab)
    nlay = len(L)
    taug, taur = np.zeros((nlay, 8)), np.zeros((nlay, 8))
    for l, d in enumerate(L):
        if d["lower"]:
            speccomb, js, fs = _sw_binary(d, "o2", strrat, 8.)
            cont = d["selffac"] * _lin(selfref, d["indself"], d["selffrac"]) + d["forfac"] * _lin(forref, d["indfor"], d["forfrac"])
            taug[l] = _sw_major2(ka, 1, d, speccomb, js, fs) + d["o3"] * abso3a + d["h2o"] * cont
            taur[l] = d["colmol"] * _lin(rayla, js, fs)
        else:
            jp, jt, jt1 = d["jp"], d["jt"], d["jt1"]
            taug[l] = d["o2"] * (d["fac00"] * kb[jt - 1, jp - 13] + d["fac10"] * kb[jt, jp - 13] +
                                 d["fac01"] * kb[jt1 - 1, jp - 12] + d["fac11"] * kb[jt1, jp - 12]) + d["o3"] * abso3b
            taur[l] = d["colmol"] * raylb
    laysolfr, ssi = laytrop, None
    for lay in range(1, laytrop + 1):
        if L[lay - 1]["jp"] < layreffr and L[lay]["jp"] >= layreffr: laysolfr = min(lay + 1, laytrop)
        if lay == laysolfr:
            _, js, fs = _sw_binary(L[lay - 1], "o2", strrat, 8.)
            ssi = _sw_source(src, js, fs, isolvar, s["scon"])
            break
    return taug, taur, ssi, laytrop


@pytest.mark.parametrize("isolvar", [-1, 0])
def test_sw_gas_optics_bands_17_and_24_against_independent_numpy(oracle, isolvar):
    from geosradiation_gridcomp_b200 import tables
    tab = tables.load_tables()
    ncol = 16
    s = make_columns(ncol, 72, seed=1717)
    s["co2vmr"][:3] *= 30.0                                       # move the binary parameter across its range
    s["h2ovmr"][3:6] *= 1e-2
    o = oracle.rrtmg_sw(s, isolvar=isolvar, taps=("taug", "pfracs", "ssi", "laytrop"))
    assert o["rc"] == 0
    for c in range(ncol):
        for fn, g in ((_sw_band17_np, slice(6, 18)), (_sw_band24_np, slice(66, 74))):
            taug, taur, ssi, laytrop = fn(s, c, tab, isolvar)
            assert laytrop == o["laytrop"][c]
            np.testing.assert_allclose(o["taug"][c, g, :].T, taug, rtol=1e-11, err_msg=f"taug col {c}")
            np.testing.assert_allclose(o["pfracs"][c, g, :].T, taur, rtol=1e-13, err_msg=f"taur col {c}")
            np.testing.assert_allclose(o["ssi"][c, g], ssi, rtol=1e-13, err_msg=f"ssi col {c}")


def test_negative_input_traps_follow_the_reference_order(oracle):
    """LW/src/rrtmg_lw_rad.F90:209-318 (`error stop` per array, in this order) and SW/src/rrtmg_sw_rad.F90:365-383
    (`_ASSERT` per array): the first offending array in the reference's own order decides, code -(100 + position)."""
    lw_order = ["play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
                "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "emis", "cldf", "ciwp", "clwp", "rei", "rel", "tauaer_lw"]
    sw_order = ["play", "plev", "tlay", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "o2vmr", "asdir", "aldir", "asdif",
                "aldif", "cldf", "ciwp", "clwp", "rei", "rel", "tauaer_sw", "ssaaer"]
    s = make_columns(8, 72, seed=5)
    for run, order in ((oracle.rrtmg_lw, lw_order), (oracle.rrtmg_sw, sw_order)):
        for i, name in enumerate(order):
            bad = dict(s)
            bad[name] = s[name].copy(order="F")
            bad[name].flat[bad[name].size // 2] = -1e-30
            assert run(bad)["rc"] == -(101 + i), name
            if i + 1 < len(order):                     # two offenders: the earlier array of the reference's list wins
                later = order[-1]
                bad[later] = s[later].copy(order="F")
                bad[later].flat[0] = -1.0
                assert run(bad)["rc"] == -(101 + i), (name, later)


# ---- the data blob against the reference's own data statements -------------------------------------------------------
_REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(_REF), reason="the reference tree is only present in the build container")
def test_committed_table_blob_is_a_fresh_extraction_of_the_reference(tmp_path):
    """Every number the oracle and the CUDA library compute from is DATA of the reference (k-distributions, Planck
    and cloud tables, reference atmospheres, NRLSSI2 cycle, McICA condensate tables): the committed blob must equal,
    byte for byte, what tools/extract_tables.py reads out of the reference sources today."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "tables.bin"
    env = dict(os.environ, RRTMG_TABLES_OUT=str(out))
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "extract_tables.py")], env=env,
                          stdout=subprocess.DEVNULL)
    committed = open(os.path.join(root, "geosradiation_gridcomp_b200", "data", "rrtmg_tables.bin"), "rb").read()
    assert out.read_bytes() == committed


@pytest.mark.skipif(not os.path.isdir(_REF), reason="the reference tree is only present in the build container")
def test_blob_values_against_a_second_minimal_parser():
    """A few constructors read straight out of the Fortran with one regular expression each (no shared code with
    tools/extract_tables.py), from the first and the last table of a module and from both RRTMG trees."""
    import re
    from geosradiation_gridcomp_b200 import tables
    tab = tables.load_tables()

    def constructor(path, lhs):
        src = open(os.path.join(_REF, path), errors="ignore").read()
        m = re.search(re.escape(lhs) + r"\s*=\s*\(/(.*?)/\)", src, re.S)
        assert m, lhs
        body = re.sub(r"!.*", "", m.group(1)).replace("&", " ")
        return np.array([float(x.lower().replace("d", "e")) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eEdD][-+]?\d+)?", body)])

    lw, sw = "GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/src/", "GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/"
    np.testing.assert_array_equal(constructor(lw + "rrtmg_lw_k_g_01.F90", "kao(:, 1, 1)"), tab["lw.kg01.kao"][:, 0, 0])
    np.testing.assert_array_equal(constructor(lw + "rrtmg_lw_k_g_01.F90", "kao(:, 2, 1)"), tab["lw.kg01.kao"][:, 1, 0])
    np.testing.assert_array_equal(constructor(sw + "rrtmg_sw_k_g_17.F90", "kbo(:, 5,13, 1)"), tab["sw.kg17.kbo"][:, 4, 0, 0])
    np.testing.assert_array_equal(constructor(sw + "rrtmg_sw_k_g_29.F90", "sfluxrefo(:)"), tab["sw.kg29.sfluxrefo"])
    np.testing.assert_array_equal(constructor(sw + "rrtmg_sw_k_g_29.F90", "irradnceo(:)"), tab["sw.kg29.irradnceo"])


def test_kiss_range_matches_the_one_number_the_reference_records():
    """SH/cloud_subcol_gen.F90:578-604 keeps the output of an `ifort` run of its own scaling at the two ends of the
    int32 range: 8.9406967E-08 and 0.9999999 in the production real*4.  The same expression in IEEE single reproduces
    both prints digit for digit; in the promoted real*8 of this project's contract the ends are 9.3746e-08 and
    0.99999991, still inside (0, 1), so a draw can never reach either cloud-fraction bound exactly."""
    f = np.float32
    lo = f(-2147483648) * f(2.328306e-10) + f(0.5)
    hi = f(2147483647) * f(2.328306e-10) + f(0.5)
    assert "%.7E" % lo == "8.9406967E-08" and "%.7f" % hi == "0.9999999"
    lo64, hi64 = -2147483648 * 2.328306e-10 + 0.5, 2147483647 * 2.328306e-10 + 0.5
    assert 0.0 < lo64 < 1e-7 and 0.9999999 < hi64 < 1.0
    draws = kiss_python((123456789, 362436069, 521288629, 916191069), 20000)
    assert lo64 <= draws.min() and draws.max() <= hi64


def test_mcica_generator_with_custom_correlation_lengths(oracle):
    """initialize_cloud_subcol_gen (SH/cloud_subcol_gen.F90:108-129): non-default decorrelation lengths."""
    from geosradiation_gridcomp_b200 import tables
    xcw = tables.load_tables()["mcica.xcw_beta"]
    s = make_columns(40, 72, seed=5151)
    cols = np.flatnonzero((s["cldf"] > 0).any(axis=1))[:4]
    T = lambda k: np.ascontiguousarray(s[k][cols].T)
    corr = (0.6, 3.0, 5.0, -30.0, 0.3, 1.1, 10.0, 35.0)
    args = (T("zm"), s["alat"][cols], 45, T("play"), T("cldf"), T("ciwp"), T("clwp"), 112)
    oracle.set_mcica(1, corr)
    try:
        m, ci, cw = oracle.generate_stochastic_clouds(*args)
    finally:
        oracle.set_mcica(1)
    pm, pci, pcw = _mcica_python(*args, xcw, corr=corr)
    np.testing.assert_array_equal(m, pm)
    np.testing.assert_allclose(cw, pcw, rtol=1e-14, atol=0)
    m0, _, _ = oracle.generate_stochastic_clouds(*args)
    assert (m0 != m).any()                      # the lengths do change the overlap
