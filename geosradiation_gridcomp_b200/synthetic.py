"""Deterministic synthetic column states for RRTMG LW/SW.

Arrays follow the reference driver interfaces (LW/src/rrtmg_lw_rad.F90:113-201,
SW/src/rrtmg_sw_rad.F90:130-357): Fortran order, column index fastest, layer 1 at the
surface, pressures in hPa, gases as mole fraction w.r.t. dry air, water paths in g/m2,
radii in microns, heights in m, latitude in radians.  Everything is fp64 and a pure function
of (seed, global column index), so any slab [col0, col0+ncol) of a larger grid can be
generated independently (used for the column-slab sharding across GPUs).
"""
import numpy as np

NBNDLW, NBNDSW = 16, 14
RGAS, GRAV = 287.04, 9.80665  # MAPL_RGAS, MAPL_GRAV as used at IRR:3352


def _eta(nlay):
    """Monotone interface coordinate eta(0)=1 (surface) .. eta(nlay)=0 (top), dense near the
    surface and in the stratosphere (hybrid-sigma-like)."""
    x = np.linspace(0.0, 1.0, nlay + 1)
    return (1.0 - x) ** 1.6 * (1.0 - 0.35 * x)


_STRIDE = 256  # doubles reserved per (stream, column); nlay <= 256


def _draw(seed, col0, ncol, stream, nper):
    """(ncol, nper) uniforms in [0,1) whose row c depends only on (seed, stream, col0 + c):
    Philox is counter based, one counter step yields four 64-bit words, so column c owns
    counters [c*_STRIDE/4, (c+1)*_STRIDE/4)."""
    assert nper <= _STRIDE
    bg = np.random.Philox(key=[np.uint64(seed), np.uint64(stream)])
    bg.advance(int(col0) * (_STRIDE // 4))
    return np.random.Generator(bg).random((ncol, _STRIDE))[:, :nper]


def make_columns(ncol, nlay=72, seed=20260118, col0=0, lit=True):
    """Return a dict of boundary arrays for columns [col0, col0+ncol)."""
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    U = lambda stream, nper=1: _draw(seed, col0, ncol, stream, nper)
    ptop = 0.01 if nlay <= 100 else 0.001
    eta = _eta(nlay)

    psfc = 520.0 + 520.0 * U(1)[:, 0]                                  # hPa
    plev = ptop + (psfc[:, None] - ptop) * eta[None, :]                # (ncol, nlay+1)
    play = 0.5 * (plev[:, :-1] + plev[:, 1:])
    play = play + 0.01 * U(2, nlay) * np.minimum(1.0, (plev[:, :-1] - plev[:, 1:]) * 0.2)  # <=1 Pa jitter
    play = np.minimum(np.maximum(play, plev[:, 1:] * (1 + 1e-12)), plev[:, :-1] * (1 - 1e-12))

    # temperature: troposphere lapse, floor, stratopause warming, per-layer noise
    def tprof(p):
        t = 288.0 * (p / 1013.0) ** 0.19
        t = np.maximum(t, 205.0)
        t = t + 65.0 * np.exp(-(np.log(np.maximum(p, 1e-4) / 1.0)) ** 2 / 6.0)
        return t
    noise = (U(3, nlay) + U(4, nlay) + U(5, nlay) - 1.5) * 6.0         # ~N(0,3K)
    tlay = np.clip(tprof(play) + noise, 180.0, 320.0)
    tlev = np.empty((ncol, nlay + 1))
    dp = plev[:, :-1] - plev[:, 1:]
    tlev[:, 1:-1] = (tlay[:, :-1] * dp[:, 1:] + tlay[:, 1:] * dp[:, :-1]) / (dp[:, :-1] + dp[:, 1:])
    tlev[:, -1] = tlev[:, -2]
    tlev[:, 0] = tlay[:, 0] + 4.0 * (U(6)[:, 0] - 0.25)
    tsfc = np.clip(tlev[:, 0] + (-3.0 + 8.0 * U(7)[:, 0]), 180.0, 335.0)

    # gases
    h2o = (np.minimum(0.03, 0.02 * (play / 1013.0) ** 3.5) + 3e-6) * (0.3 * 10.0 ** U(8, nlay))
    o3 = 8e-6 * np.exp(-(np.log(play / 10.0)) ** 2 / 3.0) + 3e-8
    ones = np.ones((ncol, nlay))
    gases = dict(h2ovmr=h2o, o3vmr=o3, co2vmr=4.2e-4 * ones, ch4vmr=1.9e-6 * ones,
                 n2ovmr=3.3e-7 * ones, o2vmr=0.209 * ones, cfc11vmr=2.2e-10 * ones,
                 cfc12vmr=5.0e-10 * ones, cfc22vmr=2.5e-10 * ones, ccl4vmr=7.5e-11 * ones)

    # clouds: 40% clear columns, 1-3 decks
    cldf = np.zeros((ncol, nlay)); ciwp = np.zeros((ncol, nlay)); clwp = np.zeros((ncol, nlay))
    ucol = U(9, 16)
    cloudy_col = ucol[:, 0] >= 0.4
    ndeck = 1 + (ucol[:, 1] * 3).astype(int)
    ub = U(10, nlay); uw = U(11, nlay); ui = U(12, nlay)
    sig = play / psfc[:, None]
    for d in range(3):
        centre = 0.25 + 0.7 * ucol[:, 2 + 2 * d]                      # sigma of deck centre
        half = 0.02 + 0.06 * ucol[:, 3 + 2 * d]
        indeck = (np.abs(sig - centre[:, None]) < half[:, None]) & cloudy_col[:, None] & (ndeck[:, None] > d)
        # Beta(0.5,0.5) via arcsine law, snapped so exact 0 and 1 occur
        frac = np.sin(0.5 * np.pi * ub) ** 2
        frac = np.where(frac > 0.97, 1.0, np.where(frac < 0.03, 0.0, frac))
        cldf = np.where(indeck, np.maximum(cldf, frac), cldf)
    warm = play > 440.0
    clwp = np.where((cldf > 0) & (warm | (uw < 0.3)), 1.0 * 200.0 ** uw, 0.0)
    ciwp = np.where((cldf > 0) & (~warm | (ui < 0.2)), 0.5 * 160.0 ** ui, 0.0)
    rel = 4.0 + 21.0 * U(13, nlay)
    rei = 15.0 + 105.0 * U(14, nlay)

    # geometry (IRR:3348-3355)
    zm = np.zeros((ncol, nlay))
    for k in range(1, nlay):
        zm[:, k] = zm[:, k - 1] + RGAS * tlev[:, k] / GRAV * (play[:, k - 1] - play[:, k]) / plev[:, k]
    alat = (U(15)[:, 0] - 0.5) * np.pi

    emis = np.repeat(0.90 + 0.10 * U(16)[:, 0][:, None], NBNDLW, axis=1)

    # aerosols: LW absorption optical depth (ncol,nlay,16); SW tau/ssa/asm (ncol,nlay,14)
    low = (play > 700.0)[:, :, None]
    ua = U(17, nlay)[:, :, None]
    tauaer_lw = np.where(low, 0.002 * 0.1 * 100.0 ** ua, 0.0) * np.linspace(1.0, 0.4, NBNDLW)[None, None, :]
    tauaer_sw = np.where(low, 0.006 * 0.1 * 100.0 ** ua, 0.0) * np.linspace(0.5, 1.5, NBNDSW)[None, None, :]
    ssaaer = (0.85 + 0.14 * U(18, nlay))[:, :, None] * np.ones((1, 1, NBNDSW))
    asmaer = (0.55 + 0.20 * U(19, nlay))[:, :, None] * np.ones((1, 1, NBNDSW))

    us = U(20, 8)
    coszen = 0.02 + 0.98 * us[:, 0] if lit else np.maximum(0.0, 2.0 * us[:, 0] - 1.0)
    out = dict(ncol=ncol, nlay=nlay,
               play=f(play), plev=f(plev), tlay=f(tlay), tlev=f(tlev), tsfc=f(tsfc), emis=f(emis),
               cldf=f(cldf), ciwp=f(ciwp), clwp=f(clwp), rei=f(rei), rel=f(rel),
               tauaer_lw=f(tauaer_lw), zm=f(zm), alat=f(alat),
               tauaer_sw=f(tauaer_sw), ssaaer=f(ssaaer), asmaer=f(asmaer),
               coszen=f(coszen),
               asdir=f(0.03 + 0.37 * us[:, 1]), asdif=f(0.03 + 0.37 * us[:, 2]),
               aldir=f(0.10 + 0.50 * us[:, 3]), aldif=f(0.10 + 0.50 * us[:, 4]))
    for k, v in gases.items():
        out[k] = f(v)
    # super-layer interfaces from the reference pressure profile (first layers above 700/400 hPa)
    pref = ptop + (1013.0 - ptop) * 0.5 * (eta[:-1] + eta[1:])
    out["cloudLM"] = int(np.argmax(pref < 700.0))      # layers 1..cloudLM are "low"
    out["cloudMH"] = int(np.argmax(pref < 400.0))
    out["dyofyr"] = 172
    out["scon"] = 1361.0
    out["adjes"] = 1.0
    out["band_output"] = np.array([1 if b in (6, 9, 10, 11) else 0 for b in range(1, 17)], dtype=np.int32)
    return out


# MAPL constants the GEOS drivers use in the Run-phase glue (MAPL is not in the reference tree; these
# are the values of MAPL_Constants: MAPL_AIRMW, MAPL_H2OMW, MAPL_O3MW, MAPL_RUNIV / MAPL_AIRMW, MAPL_GRAV)
MAPL = dict(airmw=28.965, h2omw=18.015, o3mw=47.9982, rgas=8314.47 / 28.965, grav=9.80665, undef=1.0e15)


def make_native_state(ncol, lm=72, seed=20260118, col0=0):
    """GEOS-native state of the same synthetic columns, as the two drivers hold it before their RRTMG glue
    (GEOS_IrradGridComp.F90:3237-3371, GEOS_SolarGridComp.F90:6113-6223): (ncol,LM) arrays with level 1 at
    the MODEL TOP, pressures in Pa, specific humidity and ozone mass mixing ratio, cloud water contents
    in kg/kg, extinction / scattering (/ asymmetry-weighted) aerosol optical depths.  A few negative
    mixing ratios and cloud fractions are planted so the drivers' clean-up of negatives is exercised."""
    s = make_columns(ncol, lm, seed=seed, col0=col0)
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    flip = lambda a: a[:, ::-1]
    ple = flip(s["plev"]) * 100.0                       # (ncol, LM+1) top-down, Pa
    dp = ple[:, 1:] - ple[:, :-1]
    wq = MAPL["airmw"] / MAPL["h2omw"]
    vmr = flip(s["h2ovmr"])
    q = (vmr / wq) / (1.0 + vmr / wq)
    o3 = flip(s["o3vmr"]) * (MAPL["o3mw"] / MAPL["airmw"])
    xx = 1.02 * 100 * dp
    u = _draw(seed, col0, ncol, 40, lm)
    neg = u < 0.002                                      # planted negatives
    n = dict(ncol=ncol, lm=lm, doy=int(s["dyofyr"]),
             lcldlm=lm - int(s["cloudLM"]) + 1, lcldmh=lm - int(s["cloudMH"]) + 1,
             ple=f(ple), pl=f(flip(s["play"]) * 100.0), t=f(flip(s["tlay"])),
             q=f(np.where(neg, -1e-9, q)), o3=f(np.where(_draw(seed, col0, ncol, 41, lm) < 0.002, -1e-10, o3)),
             ch4=f(flip(s["ch4vmr"])), n2o=f(flip(s["n2ovmr"])), co2=f(flip(s["co2vmr"])),
             cfc11=f(flip(s["cfc11vmr"])), cfc12=f(flip(s["cfc12vmr"])), hcfc22=f(flip(s["cfc22vmr"])),
             fcld=f(np.where(_draw(seed, col0, ncol, 42, lm) < 0.002, -1e-3, flip(s["cldf"]))),
             qliq=f(flip(s["clwp"]) / xx), qice=f(flip(s["ciwp"]) / xx),
             # radii beyond the RRTMG limits on both sides, so the drivers' clamps act
             rliq=f(flip(s["rel"]) * 2.6 - 8.0), rice=f(flip(s["rei"]) * 1.3 - 16.0),
             ts=f(s["tsfc"]), t2m=f(s["tlev"][:, 0]), emis=f(s["emis"][:, 0]), lats=f(s["alat"]),
             co2_fixed=4.2e-4, o2=0.209, ccl4=7.5e-11,
             zt=f(s["coszen"]), albvr=f(s["asdir"]), albvf=f(s["asdif"]), albnr=f(s["aldir"]), albnf=f(s["aldif"]),
             sc=float(s["scon"]), dist=float(s["adjes"]), band_output=s["band_output"])
    # aerosol system output: extinction, un-normalised scattering (tau*ssa) and asymmetry (tau*ssa*g)
    sca_lw = 0.45 * s["tauaer_lw"]
    n["taua_lw"] = f((s["tauaer_lw"] + sca_lw)[:, ::-1, :])
    n["ssaa_lw"] = f(sca_lw[:, ::-1, :])
    n["taua_sw"] = f(s["tauaer_sw"][:, ::-1, :])
    n["ssaa_sw"] = f((s["tauaer_sw"] * s["ssaaer"])[:, ::-1, :])
    n["asya_sw"] = f((s["tauaer_sw"] * s["ssaaer"] * s["asmaer"])[:, ::-1, :])
    n.update(MAPL)
    return n
