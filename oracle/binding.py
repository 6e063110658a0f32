"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under geosradiation_gridcomp_b200/ imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
BLOB = os.path.join(HERE, "..", "geosradiation_gridcomp_b200", "data", "rrtmg_tables.bin")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


class Taps(C.Structure):
    _fields_ = [(n, _ip) for n in ("jp", "jt", "jt1", "indfor", "indself", "indminor", "laytrop")] + \
               [(n, _dp) for n in ("fac00", "fac01", "fac10", "fac11")] + \
               [("cldymc", _up)] + \
               [(n, _dp) for n in ("ciwpmc", "clwpmc", "taug", "pfracs", "taucmc", "pwvcm", "ssi", "radval")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_init.argtypes = [C.c_char_p]
        rc = _lib.oracle_init(os.path.abspath(BLOB).encode())
        if rc:
            raise RuntimeError(f"oracle_init failed: {rc}")
        _lib.oracle_lw_table.restype = _dp
        _lib.oracle_sw_table.restype = _dp
        _lib.oracle_lw_table.argtypes = [C.c_char_p, C.c_int, _ip]
        _lib.oracle_sw_table.argtypes = [C.c_char_p, C.c_int, _ip]
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.f_contiguous, "expect fp64 Fortran-order arrays"
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32
    return a.ctypes.data_as(_ip)


def set_mcica(ih=1, corr=None):
    c = None if corr is None else np.ascontiguousarray(corr, dtype=np.float64).ctypes.data_as(_dp)
    rc = lib().oracle_set_mcica(int(ih), c)
    if rc:
        raise RuntimeError(f"oracle_set_mcica: {rc}")


def num_threads():
    return lib().oracle_num_threads()


def table(kind, name, band=0):
    n = C.c_int(0)
    fn = lib().oracle_lw_table if kind == "lw" else lib().oracle_sw_table
    p = fn(name.encode(), band, C.byref(n))
    if not p or n.value == 0:
        return None
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


_keep = []


def _opt(v, n):
    """optional real argument (absent -> NULL)"""
    if v is None:
        return None
    a = np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64)
    assert a.size == n
    _keep[:] = _keep[-8:] + [a]
    return a.ctypes.data_as(_dp)


def _make_taps(want, ncol, nlay, ngpt):
    """Allocate the requested tap arrays; returns (struct, dict of numpy arrays)."""
    if not want:
        return None, {}
    t = Taps()
    out = {}
    for name, ctype in Taps._fields_:
        if name not in want:
            continue
        if name in ("jp", "jt", "jt1", "indfor", "indself", "indminor"):
            a = np.zeros((ncol, nlay), dtype=np.int32, order="F")
        elif name == "laytrop":
            a = np.zeros(ncol, dtype=np.int32)
        elif name in ("fac00", "fac01", "fac10", "fac11"):
            a = np.zeros((ncol, nlay), dtype=np.float64, order="F")
        elif name == "cldymc":
            a = np.zeros((ncol, ngpt, nlay), dtype=np.uint8)         # [icol][ig][ilay]
        elif name in ("pwvcm",):
            a = np.zeros(ncol, dtype=np.float64)
        elif name == "ssi":
            a = np.zeros((ncol, ngpt), dtype=np.float64)
        elif name == "radval":   # the SOLAR_RADVAL dummies of rrtmg_sw, (ncol,120) column fastest
            a = np.zeros((ncol, 120), dtype=np.float64, order="F")
        else:
            a = np.zeros((ncol, ngpt, nlay), dtype=np.float64)       # [icol][ig][ilay]
        out[name] = a
        setattr(t, name, a.ctypes.data_as(ctype))
    return t, out


def rrtmg_lw(s, psize=4, dudTs=True, iceflg=3, liqflg=1, taps=()):
    """Run the LW oracle on a synthetic-state dict `s`; returns dict of outputs (+ taps)."""
    ncol, nlay = s["ncol"], s["nlay"]
    o = {k: np.zeros((ncol, nlay + 1), order="F") for k in
         ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")}
    o["olrb"] = np.zeros((16, ncol), order="F")
    o["dolrb_dTs"] = np.zeros((16, ncol), order="F")
    o["clearCounts"] = np.zeros((ncol, 4), dtype=np.int32, order="F")
    t, tout = _make_taps(set(taps), ncol, nlay, 140)
    bo = np.ascontiguousarray(s["band_output"], dtype=np.int32)
    rc = lib().oracle_rrtmg_lw(
        ncol, nlay, psize, int(dudTs), _d(s["play"]), _d(s["plev"]), _d(s["tlay"]), _d(s["tlev"]),
        _d(s["tsfc"]), _d(s["emis"]), _d(s["h2ovmr"]), _d(s["o3vmr"]), _d(s["co2vmr"]),
        _d(s["ch4vmr"]), _d(s["n2ovmr"]), _d(s["o2vmr"]), _d(s["cfc11vmr"]), _d(s["cfc12vmr"]),
        _d(s["cfc22vmr"]), _d(s["ccl4vmr"]), _d(s["cldf"]), _d(s["ciwp"]), _d(s["clwp"]),
        _d(s["rei"]), _d(s["rel"]), iceflg, liqflg, _d(s["tauaer_lw"]), _d(s["zm"]), _d(s["alat"]),
        int(s["dyofyr"]), int(s["cloudLM"]), int(s["cloudMH"]),
        o["clearCounts"].ctypes.data_as(_ip), _d(o["uflx"]), _d(o["dflx"]), _d(o["uflxc"]),
        _d(o["dflxc"]), _d(o["duflx_dTs"]), _d(o["duflxc_dTs"]), _i(bo), _d(o["olrb"]),
        _d(o["dolrb_dTs"]), C.byref(t) if t is not None else None)
    o["rc"] = rc
    o.update(tout)
    return o


def rrtmg_sw(s, rpart=0, isolvar=0, iceflg=3, liqflg=1, iaer=10, normFlx=1, do_drfband=False,
             taps=(), bndscl=None, indsolvar=None, solcycfrac=None, radval=False):
    """radval: the SOLAR_RADVAL build (SW/src/rrtmg_sw_rad.F90:85-122); the result gains "radval" (ncol,120)."""
    if radval:
        taps = tuple(taps) + ("radval",)
    ncol, nlay = s["ncol"], s["nlay"]
    o = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
    for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "cotdtp", "cotdhp", "cotdmp", "cotdlp",
              "cotntp", "cotnhp", "cotnmp", "cotnlp"):
        o[k] = np.zeros(ncol)
    o["fswband"] = np.zeros((ncol, 14), order="F")
    o["drband"] = np.zeros((ncol, 14), order="F")
    o["dfband"] = np.zeros((ncol, 14), order="F")
    o["clearCounts"] = np.zeros((ncol, 4), dtype=np.int32, order="F")
    t, tout = _make_taps(set(taps), ncol, nlay, 112)
    L = lib()
    L.oracle_rrtmg_sw.argtypes = None
    rc = L.oracle_rrtmg_sw(
        C.c_int(rpart), C.c_int(ncol), C.c_int(nlay), C.c_double(s["scon"]), C.c_double(s["adjes"]),
        _d(s["coszen"]), C.c_int(isolvar), _d(s["play"]), _d(s["plev"]), _d(s["tlay"]),
        _d(s["h2ovmr"]), _d(s["o3vmr"]), _d(s["co2vmr"]), _d(s["ch4vmr"]), _d(s["o2vmr"]),
        C.c_int(iceflg), C.c_int(liqflg), _d(s["cldf"]), _d(s["ciwp"]), _d(s["clwp"]), _d(s["rei"]),
        _d(s["rel"]), C.c_int(int(s["dyofyr"])), _d(s["zm"]), _d(s["alat"]), C.c_int(iaer),
        _d(s["tauaer_sw"]), _d(s["ssaaer"]), _d(s["asmaer"]), _d(s["asdir"]), _d(s["asdif"]),
        _d(s["aldir"]), _d(s["aldif"]), C.c_int(int(s["cloudLM"])), C.c_int(int(s["cloudMH"])),
        C.c_int(normFlx), o["clearCounts"].ctypes.data_as(_ip), _d(o["swuflx"]), _d(o["swdflx"]),
        _d(o["swuflxc"]), _d(o["swdflxc"]), _d(o["nirr"]), _d(o["nirf"]), _d(o["parr"]),
        _d(o["parf"]), _d(o["uvrr"]), _d(o["uvrf"]), _d(o["fswband"]), _d(o["cotdtp"]),
        _d(o["cotdhp"]), _d(o["cotdmp"]), _d(o["cotdlp"]), _d(o["cotntp"]), _d(o["cotnhp"]),
        _d(o["cotnmp"]), _d(o["cotnlp"]), C.c_int(int(do_drfband)), _d(o["drband"]), _d(o["dfband"]),
        _opt(bndscl, 14), _opt(indsolvar, 2), _opt(solcycfrac, 1), C.byref(t) if t is not None else None)
    o["rc"] = rc
    o.update(tout)
    return o


def rng_kiss(seeds, n):
    """n draws of the reference KISS generator (SH/cloud_subcol_gen.F90:546-607) from 4 int32 seeds;
    returns (draws, final seeds)."""
    L = lib()
    s = [C.c_int(int(np.int32(v))) for v in seeds]
    out = np.zeros(n)
    r = C.c_double(0.)
    for i in range(n):
        L.oracle_rng_kiss(C.byref(s[0]), C.byref(s[1]), C.byref(s[2]), C.byref(s[3]), C.byref(r))
        out[i] = r.value
    return out, [v.value for v in s]


def generate_stochastic_clouds(zmid, alat, doy, play, cldfrac, ciwp, clwp, nsubcol, seed_order=(1, 2, 3, 4),
                               cwp_tiny=1e-20):
    """Stand-alone McICA generator on (nlay,ncol) partition-layout arrays; returns
    (cldy_stoch uint8, ciwp_stoch, clwp_stoch), each (nlay,nsubcol,ncol) Fortran order."""
    nlay, ncol = play.shape
    cl = np.zeros((nlay, nsubcol, ncol), dtype=np.uint8, order="F")
    ci = np.zeros((nlay, nsubcol, ncol), order="F")
    cw = np.zeros((nlay, nsubcol, ncol), order="F")
    so = np.ascontiguousarray(seed_order, dtype=np.int32)
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    zmid, play, cldfrac, ciwp, clwp = map(f, (zmid, play, cldfrac, ciwp, clwp))
    alat = np.ascontiguousarray(alat, dtype=np.float64)
    rc = lib().oracle_generate_stochastic_clouds(
        C.c_int(ncol), C.c_int(ncol), C.c_int(nsubcol), C.c_int(nlay), _d(zmid), alat.ctypes.data_as(_dp),
        C.c_int(doy), _d(play), _d(cldfrac), _d(ciwp), _d(clwp), C.c_double(cwp_tiny),
        cl.ctypes.data_as(_up), _d(ci), _d(cw), so.ctypes.data_as(_ip))
    if rc:
        raise RuntimeError(f"generate_stochastic_clouds: {rc}")
    return cl, ci, cw


def clear_counts(cldy_stoch, cloudLM, cloudMH):
    nlay, nsub, ncol = cldy_stoch.shape
    out = np.zeros((4, ncol), dtype=np.int32, order="F")
    rc = lib().oracle_clearCounts_threeBand(C.c_int(ncol), C.c_int(ncol), C.c_int(nsub), C.c_int(nlay),
                                            C.c_int(cloudLM), C.c_int(cloudMH),
                                            np.asfortranarray(cldy_stoch).ctypes.data_as(_up), out.ctypes.data_as(_ip))
    if rc:
        raise RuntimeError(f"clearCounts_threeBand: {rc}")
    return out


# ---- Run-phase glue (oracle/glue.c) ------------------------------------------------------------
_IRR_STATE = ["ple", "pl", "t", "q", "o3", "ch4", "n2o", "co2", "cfc11", "cfc12", "hcfc22", "fcld", "qliq", "qice",
              "rliq", "rice", "ts", "t2m", "emis", "lats", "taua", "ssaa"]
_LW_INPUTS = ["play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
              "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel", "tauaer", "zm", "alat"]
_SOL_STATE = ["ple", "pl", "t", "q", "o3", "ch4", "cl", "qliq", "qice", "rliq", "rice", "ts", "taua", "ssaa", "asya"]
_SW_INPUTS = ["play", "plev", "tlay", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "o2vmr", "cld", "ciwp", "clwp", "rei",
              "rel", "zm", "tauaer", "ssaaer", "asmaer"]


class IrradState(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "lm", "iceflg", "liqflg", "lcldmh", "lcldlm")] +
                [(n, C.c_double) for n in ("co2_fixed", "o2", "ccl4", "airmw", "h2omw", "o3mw", "rgas", "grav")] +
                [(n, _dp) for n in _IRR_STATE])


class LwInputs(C.Structure):
    _fields_ = [("cloudLM", C.c_int), ("cloudMH", C.c_int)] + [(n, _dp) for n in _LW_INPUTS]


class IrradFluxes(C.Structure):
    _fields_ = [(n, _dp) for n in ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc", "sfcem", "cldtt", "cldhi",
                                   "cldmd", "cldlo")]


class SolarState(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "lm", "iceflg", "liqflg", "lcldmh", "lcldlm")] +
                [(n, C.c_double) for n in ("co2", "o2", "airmw", "h2omw", "o3mw", "rgas", "grav")] +
                [(n, _dp) for n in _SOL_STATE])


class SwInputs(C.Structure):
    _fields_ = [("cloudLM", C.c_int), ("cloudMH", C.c_int)] + [(n, _dp) for n in _SW_INPUTS]


class SolarFluxes(C.Structure):
    _fields_ = [(n, _dp) for n in ("fsw", "fsc", "fswu", "fscu", "cldts", "cldhs", "cldms", "cldls", "cottp", "cothp",
                                   "cotmp", "cotlp")]


def irrad_prepare(n, iceflg=3, liqflg=1):
    """GEOS_IrradGridComp.F90:3237-3371 on the native state `n` (synthetic.make_native_state): returns a
    synthetic-state dict in the layout of make_columns, ready for rrtmg_lw()."""
    ncol, lm = n["ncol"], n["lm"]
    s = IrradState()
    s.ncol, s.lm, s.iceflg, s.liqflg, s.lcldmh, s.lcldlm = ncol, lm, iceflg, liqflg, n["lcldmh"], n["lcldlm"]
    for k in ("co2_fixed", "o2", "ccl4", "airmw", "h2omw", "o3mw", "rgas", "grav"):
        setattr(s, k, float(n[k]))
    for k in _IRR_STATE:
        v = n.get({"taua": "taua_lw", "ssaa": "ssaa_lw"}.get(k, k))
        setattr(s, k, None if v is None else _d(v))
    o = LwInputs()
    out = {"ncol": ncol, "nlay": lm}
    for k in _LW_INPUTS:
        shape = {"plev": (ncol, lm + 1), "tlev": (ncol, lm + 1), "tsfc": (ncol,), "alat": (ncol,), "emis": (ncol, 16),
                 "tauaer": (ncol, lm, 16)}.get(k, (ncol, lm))
        out[k] = np.zeros(shape, order="F")
        setattr(o, k, _d(out[k]))
    rc = lib().oracle_irrad_prepare(C.byref(s), C.byref(o))
    if rc:
        raise RuntimeError(f"oracle_irrad_prepare: {rc}")
    out["tauaer_lw"] = out["tauaer"]
    out["cloudLM"], out["cloudMH"] = o.cloudLM, o.cloudMH
    out["dyofyr"] = n["doy"]
    out["band_output"] = n["band_output"]
    return out


def irrad_finish(n, o):
    """GEOS_IrradGridComp.F90:3486-3533 on the outputs `o` of rrtmg_lw()."""
    ncol, lm = n["ncol"], n["lm"]
    f = IrradFluxes()
    out = {k: np.zeros((ncol, lm + 1), order="F") for k in ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc")}
    out.update({k: np.zeros(ncol) for k in ("sfcem", "cldtt", "cldhi", "cldmd", "cldlo")})
    for k, v in out.items():
        setattr(f, k, _d(v))
    L = lib()
    L.oracle_irrad_finish.argtypes = None
    rc = L.oracle_irrad_finish(C.c_int(ncol), C.c_int(lm), _d(n["emis"]), o["clearCounts"].ctypes.data_as(_ip),
                               _d(o["uflx"]), _d(o["dflx"]), _d(o["uflxc"]), _d(o["dflxc"]), _d(o["duflx_dTs"]),
                               _d(o["duflxc_dTs"]), C.byref(f))
    if rc:
        raise RuntimeError(f"oracle_irrad_finish: {rc}")
    out["olrb"], out["dolrb_dts"] = o["olrb"], o["dolrb_dTs"]
    return out


def solar_prepare(n, iceflg=3, liqflg=1):
    """GEOS_SolarGridComp.F90:6113-6223: returns a synthetic-state dict ready for rrtmg_sw()."""
    ncol, lm = n["ncol"], n["lm"]
    s = SolarState()
    s.ncol, s.lm, s.iceflg, s.liqflg, s.lcldmh, s.lcldlm = ncol, lm, iceflg, liqflg, n["lcldmh"], n["lcldlm"]
    s.co2 = float(n["co2_fixed"])
    for k in ("o2", "airmw", "h2omw", "o3mw", "rgas", "grav"):
        setattr(s, k, float(n[k]))
    for k in _SOL_STATE:
        v = n.get({"taua": "taua_sw", "ssaa": "ssaa_sw", "asya": "asya_sw", "cl": "fcld"}.get(k, k))
        setattr(s, k, None if v is None else _d(v))
    o = SwInputs()
    out = {"ncol": ncol, "nlay": lm}
    for k in _SW_INPUTS:
        shape = {"plev": (ncol, lm + 1), "tauaer": (ncol, lm, 14), "ssaaer": (ncol, lm, 14),
                 "asmaer": (ncol, lm, 14)}.get(k, (ncol, lm))
        out[k] = np.zeros(shape, order="F")
        setattr(o, k, _d(out[k]))
    rc = lib().oracle_solar_prepare(C.byref(s), C.byref(o))
    if rc:
        raise RuntimeError(f"oracle_solar_prepare: {rc}")
    out["tauaer_sw"] = out["tauaer"]
    out["cldf"] = out["cld"]
    out["cloudLM"], out["cloudMH"] = o.cloudLM, o.cloudMH
    out["dyofyr"], out["scon"], out["adjes"] = n["doy"], n["sc"], n["dist"]
    out["coszen"], out["alat"] = n["zt"], n["lats"]
    out["asdir"], out["asdif"], out["aldir"], out["aldif"] = n["albvr"], n["albvf"], n["albnr"], n["albnf"]
    return out


def solar_finish(n, o):
    """GEOS_SolarGridComp.F90:6395-6447 on the outputs `o` of rrtmg_sw()."""
    ncol, lm = n["ncol"], n["lm"]
    f = SolarFluxes()
    out = {k: np.zeros((ncol, lm + 1), order="F") for k in ("fsw", "fsc", "fswu", "fscu")}
    out.update({k: np.zeros(ncol) for k in ("cldts", "cldhs", "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp")})
    for k, v in out.items():
        setattr(f, k, _d(v))
    cotd = (_dp * 4)(*[_d(o[k]) for k in ("cotdtp", "cotdhp", "cotdmp", "cotdlp")])
    cotn = (_dp * 4)(*[_d(o[k]) for k in ("cotntp", "cotnhp", "cotnmp", "cotnlp")])
    L = lib()
    L.oracle_solar_finish.argtypes = None
    rc = L.oracle_solar_finish(C.c_int(ncol), C.c_int(lm), C.c_double(n["undef"]), o["clearCounts"].ctypes.data_as(_ip),
                               _d(o["swuflx"]), _d(o["swdflx"]), _d(o["swuflxc"]), _d(o["swdflxc"]), cotd, cotn,
                               C.byref(f))
    if rc:
        raise RuntimeError(f"oracle_solar_finish: {rc}")
    for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband"):
        out[k] = o[k]
    return out


_UPD_3D = ("flx", "flc", "flxu", "flcu", "flxd", "flcd")
_UPD_2D = ("olr", "olc", "sfcem", "lws", "lcs", "flns", "flnsc")


class IrradExports(C.Structure):
    _fields_ = [(n, _dp) for n in _UPD_3D + _UPD_2D]


def irrad_update(f, ts_int, tsinst):
    """GEOS_IrradGridComp.F90 Update (:3861, :3929-3990) on the refresh outputs `f` of irrad_finish()."""
    ncol, lm1 = f["flxu"].shape
    e = IrradExports()
    out = {k: np.zeros((ncol, lm1), order="F") for k in _UPD_3D}
    out.update({k: np.zeros(ncol) for k in _UPD_2D})
    for k, v in out.items():
        setattr(e, k, _d(v))
    L = lib()
    L.oracle_irrad_update.argtypes = None
    rc = L.oracle_irrad_update(C.c_int(ncol), C.c_int(lm1 - 1), _d(f["flxu"]), _d(f["flxd"]), _d(f["flcu"]), _d(f["flcd"]),
                               _d(f["dfdts"]), _d(f["dfdtsc"]), _d(f["sfcem"]), _d(ts_int), _d(tsinst), C.byref(e))
    if rc:
        raise RuntimeError(f"oracle_irrad_update: {rc}")
    return out
