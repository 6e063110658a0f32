"""Host-side mirror of the reference driver interfaces, over the C ABI (include/rrtmgx.h).

Function names, argument names, argument meaning, array layouts and error behaviour follow
  rrtmg_lw      LW/src/rrtmg_lw_rad.F90:15-23,113-201
  rrtmg_sw      SW/src/rrtmg_sw_rad.F90:68-124,130-357
  rrtmg_lw_ini  LW/src/rrtmg_lw_init.F90:22      rrtmg_sw_ini  SW/src/rrtmg_sw_init.F90:49
  set_inhomogeneity            GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90:45
  initialize_cloud_subcol_gen  GEOS_RadiationShared/cloud_subcol_gen.F90:108
(LW/ = GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/, SW/ = GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/).

Arrays are Fortran-ordered fp64 with the column index fastest, layer 1 at the surface.  Array
arguments may be numpy arrays (host pointers; the library stages them through the GPU) or
torch CUDA tensors / raw device addresses (``device=True``; no copies).  The compute path is
the CUDA library only: importing this module without ``librrtmgx.so`` or calling it without
a GPU raises - there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librrtmgx.so")
BLOB_PATH = os.path.join(HERE, "data", "rrtmg_tables.bin")

NBNDLW, NGPTLW, NBNDSW, NGPTSW = 16, 140, 14, 112
DEVICE_PTRS, NO_SYNC, SKIP_CHECKS, KEEP_STATUS, REUSE_CLOUDS, F32_ARRAYS, LIT_ONLY = 1, 2, 4, 8, 16, 32, 64

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


class RrtmgxError(RuntimeError):
    """Raised where the reference would `error stop` (LW) or set RC /= 0 (SW)."""

    def __init__(self, status, message):
        super().__init__(f"rrtmgx status {status}: {message}")
        self.status = status


class Config(C.Structure):
    _fields_ = [("table_blob", C.c_char_p), ("device", C.c_int), ("inhomogeneity", C.c_int), ("corr", _dp)]


_LW_IN = ["play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr",
          "o2vmr", "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel",
          "tauaer", "zm", "alat"]
_LW_OUT = ["uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs", "olrb", "dolrb_dTs"]


class LwArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "nlay", "psize", "dudTs", "iceflglw", "liqflglw", "dyofyr",
                                        "cloudLM", "cloudMH", "flags")] +
                [("stream", _vp)] + [(n, _vp) for n in _LW_IN] + [("band_output", _vp), ("clearCounts", _vp)] +
                [(n, _vp) for n in _LW_OUT])


class SwNoAerosol(C.Structure):
    _fields_ = [(n, _vp) for n in ("swuflx", "swdflx", "swuflxc", "swdflxc", "fswband")]


class LwVariants(C.Structure):
    _fields_ = [("nvar", C.c_int), ("gas", _vp), ("uflx", _vp), ("dflx", _vp), ("duflx_dTs", _vp)]


GAS = {"H2O": 1, "O3": 2, "CO2": 3, "CH4": 4, "N2O": 5, "CFC11": 6, "CFC12": 7, "HCFC22": 8}   # RRTMGX_GAS_*

_SW_IN = ["coszen", "play", "plev", "tlay", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "o2vmr", "cld", "ciwp",
          "clwp", "rei", "rel", "zm", "alat", "tauaer", "ssaaer", "asmaer", "asdir", "asdif", "aldir", "aldif"]
_SW_OUT = ["swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband",
           "cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp", "drband", "dfband"]


class SwArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "nlay", "rpart", "isolvar", "iceflgsw", "liqflgsw", "dyofyr",
                                        "cloudLM", "cloudMH", "iaer", "normFlx", "do_drfband", "flags")] +
                [("stream", _vp), ("scon", C.c_double), ("adjes", C.c_double), ("bndscl", _vp),
                 ("indsolvar", _vp), ("solcycfrac", _vp)] +
                [(n, _vp) for n in _SW_IN] + [("clearCounts", _vp)] + [(n, _vp) for n in _SW_OUT] + [("radval", _vp)])


# The SOLAR_RADVAL dummies of rrtmg_sw in the order of its argument list (SW/src/rrtmg_sw_rad.F90:85-122; include/rrtmgx.h
# RRTMGX_RADVAL_FAMILIES): column q of the (ncol,120) `radval` array is RADVAL_NAMES[q].
NRADVAL = 120
RADVAL_FAMILIES = ("cds", "cotl", "cdsl", "coti", "cdsi", "ssal", "sdsl", "ssai", "sdsi", "asml", "adsl", "asmi", "adsi",
                   "forl", "fori")
RADVAL_NAMES = tuple(f + dn + lev + "p" for f in RADVAL_FAMILIES for dn in "dn" for lev in "thml")


# ---- fused Run-phase glue (include/rrtmgx.h: RrtmgxIrradArgs, RrtmgxSolarArgs) ---------------------
_IRR_IN = ["ple", "pl", "t", "q", "o3", "ch4", "n2o", "co2", "cfc11", "cfc12", "hcfc22", "fcld", "qliq", "qice",
           "rliq", "rice", "ts", "t2m", "emis", "lats", "taua", "ssaa"]
_IRR_OUT = ["flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc", "sfcem", "cldtt", "cldhi", "cldmd", "cldlo", "olrb",
            "dolrb_dts"]


class IrradArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "lm", "iceflg", "liqflg", "doy", "lcldmh", "lcldlm", "flags")] +
                [("stream", _vp)] +
                [(n, C.c_double) for n in ("co2_fixed", "o2", "ccl4", "airmw", "h2omw", "o3mw", "rgas", "grav")] +
                [(n, _vp) for n in _IRR_IN] + [("band_output", _vp)] + [(n, _vp) for n in _IRR_OUT])


_UPD_IN = ["flxu_int", "flxd_int", "flcu_int", "flcd_int", "dfdts", "dfdtsc", "sfcem_int", "ts_int", "tsinst"]
_UPD_OUT = ["flx", "flc", "flxu", "flcu", "flxd", "flcd", "olr", "olc", "sfcem", "lws", "lcs", "flns", "flnsc"]


class IrradUpdateArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "lm", "flags")] + [("stream", _vp)] +
                [(n, _vp) for n in _UPD_IN + _UPD_OUT])


_SOL_IN = ["ple", "pl", "t", "q", "o3", "ch4", "cl", "qliq", "qice", "rliq", "rice", "ts", "zt", "lats", "albvr",
           "albvf", "albnr", "albnf", "taua", "ssaa", "asya"]
_SOL_OUT = ["fsw", "fsc", "fswu", "fscu", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband", "cldts", "cldhs",
            "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp"]


class SolarArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("ncol", "lm", "iceflg", "liqflg", "doy", "isolvar", "lcldmh", "lcldlm",
                                        "flags")] +
                [("stream", _vp)] +
                [(n, C.c_double) for n in ("sc", "dist", "co2", "o2", "airmw", "h2omw", "o3mw", "rgas", "grav",
                                           "undef")] +
                [("solcycfrac", _vp)] + [(n, _vp) for n in _SOL_IN] + [(n, _vp) for n in _SOL_OUT])


_TAP_I = ["jp", "jt", "jt1", "indfor", "indself", "indminor", "laytrop"]
_TAP_D = ["fac00", "fac01", "fac10", "fac11"]


class Taps(C.Structure):
    _fields_ = ([(n, _vp) for n in _TAP_I] + [(n, _vp) for n in _TAP_D] +
                [("cldymc", _vp), ("taucmc", _vp), ("pwvcm", _vp), ("taug", _vp), ("pfracs", _vp), ("ssi", _vp)])


_lib = None
_initialised = False


def lib():
    """Load librrtmgx.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                              "(geosradiation_gridcomp_b200/csrc/build.sh); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rrtmgx_init.argtypes = [C.POINTER(Config)]
        L.rrtmgx_set_mcica.argtypes = [C.c_int, _dp]
        L.rrtmgx_strerror.restype = C.c_char_p
        L.rrtmgx_strerror.argtypes = [C.c_int]
        L.rrtmgx_launch_count.restype = C.c_longlong
        L.rrtmgx_profile.argtypes = [C.c_int]
        L.rrtmgx_profile.restype = None
        L.rrtmgx_profile_report.argtypes = [C.c_char_p, C.c_size_t]
        L.rrtmgx_profile_report.restype = C.c_size_t
        L.rrtmgx_lw_run.argtypes = [C.POINTER(LwArgs)]
        L.rrtmgx_sw_run.argtypes = [C.POINTER(SwArgs)]
        L.rrtmgx_set_taps.argtypes = [C.POINTER(Taps), C.POINTER(Taps)]
        L.rrtmgx_table.restype = _dp
        L.rrtmgx_table.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _ip]
        L.rrtmgx_sw_run_with_clean.argtypes = [C.POINTER(SwArgs), C.POINTER(SwNoAerosol)]
        L.rrtmgx_lw_run_variants.argtypes = [C.POINTER(LwArgs), C.POINTER(LwVariants)]
        L.rrtmgx_irrad_refresh.argtypes = [C.POINTER(IrradArgs)]
        L.rrtmgx_irrad_prepare.argtypes = [C.POINTER(IrradArgs), C.POINTER(LwArgs)]
        L.rrtmgx_irrad_update.argtypes = [C.POINTER(IrradUpdateArgs)]
        L.rrtmgx_solar_refresh.argtypes = [C.POINTER(SolarArgs)]
        L.rrtmgx_solar_prepare.argtypes = [C.POINTER(SolarArgs), C.POINTER(SwArgs)]
        L.rrtmgx_debug_divide.argtypes = [C.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp]
        L.rrtmgx_heating_rate.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int, _vp]
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise RrtmgxError(status, lib().rrtmgx_strerror(status).decode())


def init(device=-1, inhomogeneity=-1, corr=None, table_blob=None):
    """rrtmg_lw_ini + rrtmg_sw_ini; idempotent.  `inhomogeneity` (0..2; -1 = the default, beta) and `corr` are
    applied by the FIRST call only: on an initialised library the call is a pure no-op, like the reference's _ini
    routines, which never touch the McICA module state (set_inhomogeneity / initialize_cloud_subcol_gen own it)."""
    global _initialised
    c = Config()
    c.table_blob = (table_blob or BLOB_PATH).encode()
    c.device = device
    c.inhomogeneity = inhomogeneity
    keep = None
    if corr is not None:
        keep = np.ascontiguousarray(corr, dtype=np.float64)
        assert keep.size == 8
        c.corr = keep.ctypes.data_as(_dp)
    _check(lib().rrtmgx_init(C.byref(c)))
    _initialised = True


def rrtmg_lw_ini():
    """LW/src/rrtmg_lw_init.F90:22 (GEOS calls it on every refresh; idempotent here)."""
    init()


def rrtmg_sw_ini():
    """SW/src/rrtmg_sw_init.F90:49."""
    init()


_mcica = {"ih": 1, "corr": None}


def set_inhomogeneity(ih):
    """GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90:45 (0 homogeneous, 1 beta, 2 gamma)."""
    if not _initialised:
        init()
    _mcica["ih"] = int(ih)
    _apply_mcica()


def initialize_cloud_subcol_gen(adl_am1, adl_am2, adl_am30, adl_am4, rdl_am1, rdl_am2, rdl_am30, rdl_am4):
    """GEOS_RadiationShared/cloud_subcol_gen.F90:108-129."""
    if not _initialised:
        init()
    _mcica["corr"] = [adl_am1, adl_am2, adl_am30, adl_am4, rdl_am1, rdl_am2, rdl_am30, rdl_am4]
    _apply_mcica()


def _apply_mcica():
    corr = _mcica["corr"]
    p = None
    if corr is not None:
        arr = np.ascontiguousarray(corr, dtype=np.float64)
        p = arr.ctypes.data_as(_dp)
    _check(lib().rrtmgx_set_mcica(_mcica["ih"], p))


def finalize():
    global _initialised
    if _lib is not None:
        _lib.rrtmgx_finalize()
    _initialised = False


def knobs():
    """{'chunk', 'host_chunk', 'stages', 'ih'} in force (rrtmgx_get_knobs)."""
    k = (C.c_longlong * 4)()
    _check(lib().rrtmgx_get_knobs(k))
    return {"chunk": int(k[0]), "host_chunk": int(k[1]), "stages": int(k[2]), "ih": int(k[3])}


def launch_count():
    return int(lib().rrtmgx_launch_count())


def profile(enable):
    """Switch per-kernel CUDA-event timing on (clears totals) or off; profiled steps are serialised."""
    lib().rrtmgx_profile(1 if enable else 0)


def profile_report():
    """{kernel name: (launches, total device ms)} accumulated since profile(True)."""
    n = lib().rrtmgx_profile_report(None, 0)
    buf = C.create_string_buffer(int(n) + 16)
    lib().rrtmgx_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out


def table(kind, name, band=0):
    n = C.c_int32(0)
    p = lib().rrtmgx_table(kind.encode(), name.encode(), band, C.byref(n))
    if not p or n.value == 0:
        return None
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def _addr(a, device, dtype=np.float64, keep=None):
    """Raw address of a numpy array (host) or torch CUDA tensor / int (device)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):   # torch tensor
        if device != a.is_cuda:
            raise ValueError("mixing host and device arrays in one call")
        return a.data_ptr()
    a = np.asarray(a)
    if device:
        raise ValueError("device=True needs torch CUDA tensors or raw addresses")
    if a.dtype != dtype or not (a.flags.f_contiguous or a.ndim <= 1):
        raise ValueError("arrays must be Fortran-ordered with the reference's element type")
    if keep is not None:
        keep.append(a)
    return a.ctypes.data


def _new_taps(names, ncol, nlay, ngpt):
    t = Taps()
    out = {}
    for n in names:
        if n in _TAP_I[:-1]:
            a = np.zeros((ncol, nlay), dtype=np.int32, order="F")
        elif n == "laytrop":
            a = np.zeros(ncol, dtype=np.int32)
        elif n in _TAP_D:
            a = np.zeros((ncol, nlay), order="F")
        elif n == "cldymc":
            a = np.zeros((ncol, ngpt, nlay), dtype=np.uint8, order="F")   # [ilay][ig][icol] in memory
        elif n == "pwvcm":
            a = np.zeros(ncol)
        elif n == "ssi":
            a = np.zeros((ncol, ngpt), order="F")
        else:
            a = np.zeros((ncol, ngpt, nlay), order="F")
        out[n] = a
        setattr(t, n, a.ctypes.data)
    return t, out


def rrtmg_lw(ncol, nlay, psize, dudTs, play, plev, tlay, tlev, tsfc, emis, h2ovmr, o3vmr, co2vmr, ch4vmr,
             n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, cldf, ciwp, clwp, rei, rel, iceflglw,
             liqflglw, tauaer, zm, alat, dyofyr, cloudLM, cloudMH, clearCounts, uflx, dflx, uflxc, dflxc,
             duflx_dTs, duflxc_dTs, band_output, olrb, dolrb_dTs, *, device=False, stream=None, sync=True,
             skip_checks=False, reuse_clouds=False, f32=False, taps=(), rats=None):
    """Drop-in for `rrtmg_lw` (LW/src/rrtmg_lw_rad.F90:15-23): same argument order and meaning;
    outputs are written in place.  Raises RrtmgxError where the reference stops.
    Returns a dict of requested intermediate taps (tests only).
    rats = (names, uflxrat, dflxrat, duflx_dt_rat): the removed-gas loop of the LW driver in the same call
    (IRR:3405-3468), arrays (ncol,nlay+1,len(names))."""
    if not _initialised:
        init()
    keep = []
    a = LwArgs()
    a.ncol, a.nlay, a.psize, a.dudTs = int(ncol), int(nlay), int(psize), int(bool(dudTs))
    a.iceflglw, a.liqflglw, a.dyofyr = int(iceflglw), int(liqflglw), int(dyofyr)
    a.cloudLM, a.cloudMH = int(cloudLM), int(cloudMH)
    a.flags = ((DEVICE_PTRS if device else 0) | (0 if sync else NO_SYNC) | (SKIP_CHECKS if skip_checks else 0) |
               (REUSE_CLOUDS if reuse_clouds else 0) | (F32_ARRAYS if f32 else 0))
    rk = np.float32 if f32 else np.float64   # element kind of the caller's real arrays
    a.stream = stream
    loc = locals()
    for n in _LW_IN + _LW_OUT:
        setattr(a, n, _addr(loc[n], device, dtype=rk, keep=keep))
    a.clearCounts = _addr(clearCounts, device, dtype=np.int32, keep=keep)
    bo = np.ascontiguousarray(band_output, dtype=np.int32)   # logical(16), always host
    a.band_output = bo.ctypes.data
    t, tout = (None, {})
    if taps:
        t, tout = _new_taps(taps, ncol, nlay, NGPTLW)
        lib().rrtmgx_set_taps(C.byref(t), None)
    try:
        if rats is None:
            _check(lib().rrtmgx_lw_run(C.byref(a)))
        else:
            names, ur, dr, dur = rats
            v = LwVariants()
            gas = np.array([GAS[n] for n in names], dtype=np.int32)
            v.nvar, v.gas = len(names), gas.ctypes.data
            v.uflx, v.dflx, v.duflx_dTs = (_addr(x, device, dtype=rk, keep=keep) for x in (ur, dr, dur))
            _check(lib().rrtmgx_lw_run_variants(C.byref(a), C.byref(v)))
    finally:
        if taps:
            lib().rrtmgx_set_taps(None, None)
    return tout


def rrtmg_sw(rpart, ncol, nlay, scon, adjes, coszen, isolvar, play, plev, tlay, h2ovmr, o3vmr, co2vmr, ch4vmr,
             o2vmr, iceflgsw, liqflgsw, cld, ciwp, clwp, rei, rel, dyofyr, zm, alat, iaer, tauaer, ssaaer, asmaer,
             asdir, asdif, aldir, aldif, cloudLM, cloudMH, normFlx, clearCounts, swuflx, swdflx, swuflxc, swdflxc,
             nirr, nirf, parr, parf, uvrr, uvrf, fswband, cotdtp, cotdhp, cotdmp, cotdlp, cotntp, cotnhp, cotnmp,
             cotnlp, do_drfband=False, drband=None, dfband=None, bndscl=None, indsolvar=None, solcycfrac=None, *,
             device=False, stream=None, sync=True, skip_checks=False, reuse_clouds=False, f32=False, taps=(),
             clean=None, radval=None):
    """Drop-in for `rrtmg_sw` (SW/src/rrtmg_sw_rad.F90:68-124) without the MAPL handle (used by
    the reference only for timers and asserts).  Outputs are written in place.
    radval: an (ncol,120) array (column fastest) selects the SOLAR_RADVAL build and receives its extra dummies
    (:85-122) in RADVAL_NAMES order."""
    if not _initialised:
        init()
    keep = []
    a = SwArgs()
    a.ncol, a.nlay, a.rpart, a.isolvar = int(ncol), int(nlay), int(rpart), int(isolvar)
    a.iceflgsw, a.liqflgsw, a.dyofyr = int(iceflgsw), int(liqflgsw), int(dyofyr)
    a.cloudLM, a.cloudMH, a.iaer = int(cloudLM), int(cloudMH), int(iaer)
    a.normFlx, a.do_drfband = int(bool(normFlx)), int(bool(do_drfband))
    a.flags = ((DEVICE_PTRS if device else 0) | (0 if sync else NO_SYNC) | (SKIP_CHECKS if skip_checks else 0) |
               (REUSE_CLOUDS if reuse_clouds else 0) | (F32_ARRAYS if f32 else 0))
    rk = np.float32 if f32 else np.float64   # element kind of the caller's real arrays
    a.stream = stream
    a.scon, a.adjes = float(scon), float(adjes)
    for n, v, cnt in (("bndscl", bndscl, 14), ("indsolvar", indsolvar, 2), ("solcycfrac", solcycfrac, 1)):
        if v is not None:
            arr = np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64)
            assert arr.size == cnt
            keep.append(arr)
            setattr(a, n, arr.ctypes.data)
    loc = locals()
    for n in _SW_IN + _SW_OUT:
        setattr(a, n, _addr(loc[n], device, dtype=rk, keep=keep))
    a.clearCounts = _addr(clearCounts, device, dtype=np.int32, keep=keep)
    if radval is not None:
        a.radval = _addr(radval, device, dtype=rk, keep=keep)
    t, tout = (None, {})
    if taps:
        t, tout = _new_taps(taps, ncol, nlay, NGPTSW)
        lib().rrtmgx_set_taps(None, C.byref(t))
    try:
        if clean is None:
            _check(lib().rrtmgx_sw_run(C.byref(a)))
        else:   # dict of the no-aerosol outputs: the two SORADCORE passes in one call (SOL:3249-3287)
            na = SwNoAerosol()
            for n in ("swuflx", "swdflx", "swuflxc", "swdflxc", "fswband"):
                setattr(na, n, _addr(clean[n], device, dtype=rk, keep=keep))
            _check(lib().rrtmgx_sw_run_with_clean(C.byref(a), C.byref(na)))
    finally:
        if taps:
            lib().rrtmgx_set_taps(None, None)
    return tout


def lw_status():
    return lib().rrtmgx_lw_status()


def sw_status():
    return lib().rrtmgx_sw_status()


def heating_rate(fnet_up_minus_down, plev, grav=9.80665, cp=1004.68506, device=False, stream=None, out=None):
    """GEOS_RadiationGridComp.F90:798-819 (RADLW / RADSW) in K/day; arrays (ncol,nlay+1) -> (ncol,nlay)."""
    if not _initialised:
        init()
    if device:
        ncol, nlev = fnet_up_minus_down.shape[-1], fnet_up_minus_down.shape[0]   # torch [lev][col]
    else:
        ncol, nlev = fnet_up_minus_down.shape
        if out is None:
            out = np.zeros((ncol, nlev - 1), order="F")
    _check(lib().rrtmgx_heating_rate(ncol, nlev - 1, _addr(fnet_up_minus_down, device), _addr(plev, device),
                                     _addr(out, device), grav, cp, DEVICE_PTRS if device else 0, stream))
    return out


def debug_divide(a, b):
    """Test hook (include/rrtmgx.h): the band kernels' branch-free a/b and 1/b beside the IEEE ones."""
    if not _initialised:
        init()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = [np.empty_like(a) for _ in range(4)]
    _check(lib().rrtmgx_debug_divide(a.size, _addr(a, False), _addr(b, False), *[_addr(o, False) for o in out]))
    return tuple(out)


def debug_kiss(seeds, ndraw=0, jump_table=None, values=None):
    """Test hook (include/rrtmgx.h rrtmgx_debug_kiss): the device KISS generator on its own.  seeds: (nstream, 4) int32.
    Returns a dict: kiss / ran8 / ran4 (nstream, ndraw); jumped / replayed (nstream, 2*nsub, 4) for
    jump_table=(nsub, nlay, inhomo); val8 / val4 for the integers `values`."""
    if not _initialised:
        init()
    seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 4)
    ns = seeds.shape[0]
    out = {"kiss": np.zeros((ns, ndraw), dtype=np.int32), "ran8": np.zeros((ns, ndraw)),
           "ran4": np.zeros((ns, ndraw), dtype=np.float32)}
    nsub, nlay, inhomo = jump_table if jump_table else (0, 0, 0)
    out["jumped"] = np.zeros((ns, 2 * nsub, 4), dtype=np.uint32)
    out["replayed"] = np.zeros((ns, 2 * nsub, 4), dtype=np.uint32)
    v = np.ascontiguousarray(values if values is not None else [], dtype=np.int32)
    out["val8"], out["val4"] = np.zeros(v.size), np.zeros(v.size, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L = lib()
    L.rrtmgx_debug_kiss.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.rrtmgx_debug_kiss(ns, p(seeds), int(ndraw), p(out["kiss"]), p(out["ran8"]), p(out["ran4"]), int(nsub),
                               int(nlay), int(inhomo), p(out["jumped"]), p(out["replayed"]), int(v.size), p(v),
                               p(out["val8"]), p(out["val4"])))
    return out


# ---- fused Run-phase glue: GEOS-native state in, GEOS-native fluxes out ---------------------------------
def _irrad_args(n, iceflg, liqflg, device, keep, f32=False):
    a = IrradArgs()
    a.ncol, a.lm, a.iceflg, a.liqflg, a.doy = int(n["ncol"]), int(n["lm"]), int(iceflg), int(liqflg), int(n["doy"])
    a.lcldmh, a.lcldlm = int(n["lcldmh"]), int(n["lcldlm"])
    a.flags = (DEVICE_PTRS if device else 0) | (F32_ARRAYS if f32 else 0)
    for k in ("co2_fixed", "o2", "ccl4", "airmw", "h2omw", "o3mw", "rgas", "grav"):
        setattr(a, k, float(n[k]))
    for k in _IRR_IN:
        v = n.get({"taua": "taua_lw", "ssaa": "ssaa_lw"}.get(k, k))
        setattr(a, k, _addr(v, device, dtype=np.float32 if f32 else np.float64, keep=keep))
    bo = np.ascontiguousarray(n["band_output"], dtype=np.int32)
    keep.append(bo)
    a.band_output = bo.ctypes.data
    return a


def irrad_prepare(n, iceflg=3, liqflg=1):
    """The first half of the LW driver glue (GEOS_IrradGridComp.F90:3237-3371) on the device: returns the
    rrtmg_lw input arrays (layer 1 at the surface, hPa, vmr, g/m2) of the native state `n`
    (synthetic.make_native_state) plus cloudLM / cloudMH."""
    if not _initialised:
        init()
    keep = []
    ncol, lm = n["ncol"], n["lm"]
    a = _irrad_args(n, iceflg, liqflg, False, keep)
    o = {}
    lw = LwArgs()
    for k in _LW_IN:
        shape = {"plev": (ncol, lm + 1), "tlev": (ncol, lm + 1), "tsfc": (ncol,), "alat": (ncol,), "emis": (ncol, 16),
                 "tauaer": (ncol, lm, 16)}.get(k, (ncol, lm))
        o[k] = np.zeros(shape, order="F")
        setattr(lw, k, o[k].ctypes.data)
    _check(lib().rrtmgx_irrad_prepare(C.byref(a), C.byref(lw)))
    o["cloudLM"], o["cloudMH"] = lw.cloudLM, lw.cloudMH
    return o


def irrad_refresh(n, iceflg=3, liqflg=1, device=False, out=None, f32=False):
    """One LW refresh from the GEOS-native state: what LW_Driver does between :3237 and :3547 with RRTMG
    (flip / units / TLEV / ZM -> rrtmg_lw -> unflip / sign / SFCEM / cloud fractions), fused on the device.
    Returns the native outputs: FLXU_INT ... DFDTSC (ncol,0:LM) top-down, upward negative; SFCEM_INT;
    CLDTTLW, CLDHILW, CLDMDLW, CLDLOLW; OLRB / DOLRB_DTS (16,ncol)."""
    if not _initialised:
        init()
    keep = []
    ncol, lm = n["ncol"], n["lm"]
    a = _irrad_args(n, iceflg, liqflg, device, keep, f32)
    rk = np.float32 if f32 else np.float64
    if out is None:
        if device:
            import torch
            z = lambda *sh: torch.zeros(tuple(reversed(sh)), dtype=torch.float64, device="cuda")
        else:
            z = lambda *sh: np.zeros(sh, dtype=rk, order="F")
        out = {k: z(ncol, lm + 1) for k in _IRR_OUT[:6]}
        out.update({k: z(ncol) for k in _IRR_OUT[6:11]})
        out["olrb"], out["dolrb_dts"] = z(16, ncol), z(16, ncol)
    for k in _IRR_OUT:
        setattr(a, k, _addr(out[k], device, dtype=rk, keep=keep))
    _check(lib().rrtmgx_irrad_refresh(C.byref(a)))
    return out


def irrad_update(f, ts_int, tsinst, device=False, want=None):
    """The between-refresh linear update of the LW exports (GEOS_IrradGridComp.F90 Update :3861, :3929-3990)
    from the outputs `f` of irrad_refresh(); `want` selects exports (default: all)."""
    if not _initialised:
        init()
    keep = []
    if device:
        ncol, lm1 = f["flxu"].shape[-1], f["flxu"].shape[0]
        import torch
        z = lambda *sh: torch.zeros(tuple(reversed(sh)), dtype=torch.float64, device="cuda")
    else:
        ncol, lm1 = f["flxu"].shape
        z = lambda *sh: np.zeros(sh, order="F")
    a = IrradUpdateArgs()
    a.ncol, a.lm, a.flags = int(ncol), int(lm1 - 1), DEVICE_PTRS if device else 0
    src = {"flxu_int": f["flxu"], "flxd_int": f["flxd"], "flcu_int": f["flcu"], "flcd_int": f["flcd"],
           "dfdts": f["dfdts"], "dfdtsc": f["dfdtsc"], "sfcem_int": f["sfcem"], "ts_int": ts_int, "tsinst": tsinst}
    for k in _UPD_IN:
        setattr(a, k, _addr(src[k], device, keep=keep))
    out = {}
    for k in (want or _UPD_OUT):
        out[k] = z(ncol, lm1) if k in _UPD_OUT[:6] else z(ncol)
        setattr(a, k, _addr(out[k], device, keep=keep))
    _check(lib().rrtmgx_irrad_update(C.byref(a)))
    return out


def _solar_args(n, iceflg, liqflg, isolvar, device, keep, f32=False):
    a = SolarArgs()
    a.ncol, a.lm, a.iceflg, a.liqflg, a.doy = int(n["ncol"]), int(n["lm"]), int(iceflg), int(liqflg), int(n["doy"])
    a.isolvar, a.lcldmh, a.lcldlm = int(isolvar), int(n["lcldmh"]), int(n["lcldlm"])
    a.flags = (DEVICE_PTRS if device else 0) | (F32_ARRAYS if f32 else 0)
    a.co2 = float(n["co2_fixed"])
    for k in ("sc", "dist", "o2", "airmw", "h2omw", "o3mw", "rgas", "grav", "undef"):
        setattr(a, k, float(n[k]))
    for k in _SOL_IN:
        v = n.get({"taua": "taua_sw", "ssaa": "ssaa_sw", "asya": "asya_sw", "cl": "fcld"}.get(k, k))
        setattr(a, k, _addr(v, device, dtype=np.float32 if f32 else np.float64, keep=keep))
    return a


def solar_prepare(n, iceflg=3, liqflg=1, isolvar=0):
    """The first half of the SW driver glue (GEOS_SolarGridComp.F90:6113-6223) on the device: the rrtmg_sw
    input arrays of the native state `n` plus cloudLM / cloudMH."""
    if not _initialised:
        init()
    keep = []
    ncol, lm = n["ncol"], n["lm"]
    a = _solar_args(n, iceflg, liqflg, isolvar, False, keep)
    o = {}
    sw = SwArgs()
    for k in _SW_IN:
        shape = {"plev": (ncol, lm + 1), "coszen": (ncol,), "alat": (ncol,), "asdir": (ncol,), "asdif": (ncol,),
                 "aldir": (ncol,), "aldif": (ncol,), "tauaer": (ncol, lm, 14), "ssaaer": (ncol, lm, 14),
                 "asmaer": (ncol, lm, 14)}.get(k, (ncol, lm))
        o[k] = np.zeros(shape, order="F")
        setattr(sw, k, o[k].ctypes.data)
    _check(lib().rrtmgx_solar_prepare(C.byref(a), C.byref(sw)))
    o["cloudLM"], o["cloudMH"] = sw.cloudLM, sw.cloudMH
    return o


def solar_refresh(n, iceflg=3, liqflg=1, isolvar=0, device=False, out=None, f32=False, lit_only=False):
    """One SW refresh from the GEOS-native state (SORADCORE :6113-6447 around rrtmg_sw), fused on the device.
    Returns FSW, FSC, FSWU, FSCU (ncol,LM+1) top-down, the surface diagnostics, CLDTS..CLDLS and COTTP..COTLP
    (MAPL_UNDEF where no cloud).  lit_only: only the columns with ZTH > 0 are run (the driver's PackIt / UnPackIt,
    GEOS_SolarGridComp.F90:3686-3687, :7753-7799); night columns get 0 fluxes and MAPL_UNDEF diagnostics."""
    if not _initialised:
        init()
    keep = []
    ncol, lm = n["ncol"], n["lm"]
    a = _solar_args(n, iceflg, liqflg, isolvar, device, keep, f32)
    if lit_only:
        a.flags |= LIT_ONLY
    rk = np.float32 if f32 else np.float64
    if out is None:
        if device:
            import torch
            z = lambda *sh: torch.zeros(tuple(reversed(sh)), dtype=torch.float64, device="cuda")
        else:
            z = lambda *sh: np.zeros(sh, dtype=rk, order="F")
        out = {k: z(ncol, lm + 1) for k in _SOL_OUT[:4]}
        out.update({k: z(ncol) for k in _SOL_OUT[4:10] + _SOL_OUT[11:]})
        out["fswband"] = z(ncol, 14)
    for k in _SOL_OUT:
        setattr(a, k, _addr(out[k], device, dtype=rk, keep=keep))
    _check(lib().rrtmgx_solar_refresh(C.byref(a)))
    return out


# ---- convenience wrappers over the synthetic-state dicts of synthetic.make_columns ---------------
def alloc_lw_outputs(ncol, nlay):
    o = {k: np.zeros((ncol, nlay + 1), order="F") for k in
         ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")}
    o["olrb"] = np.zeros((16, ncol), order="F")
    o["dolrb_dTs"] = np.zeros((16, ncol), order="F")
    o["clearCounts"] = np.zeros((ncol, 4), dtype=np.int32, order="F")
    return o


def run_lw(s, psize=4, dudTs=True, iceflg=3, liqflg=1, taps=(), out=None, **kw):
    """rrtmg_lw on a synthetic.make_columns state; returns the output dict (+ taps)."""
    o = out if out is not None else alloc_lw_outputs(s["ncol"], s["nlay"])
    t = rrtmg_lw(s["ncol"], s["nlay"], psize, dudTs, s["play"], s["plev"], s["tlay"], s["tlev"], s["tsfc"],
                 s["emis"], s["h2ovmr"], s["o3vmr"], s["co2vmr"], s["ch4vmr"], s["n2ovmr"], s["o2vmr"],
                 s["cfc11vmr"], s["cfc12vmr"], s["cfc22vmr"], s["ccl4vmr"], s["cldf"], s["ciwp"], s["clwp"],
                 s["rei"], s["rel"], iceflg, liqflg, s["tauaer_lw"], s["zm"], s["alat"], s["dyofyr"],
                 s["cloudLM"], s["cloudMH"], o["clearCounts"], o["uflx"], o["dflx"], o["uflxc"], o["dflxc"],
                 o["duflx_dTs"], o["duflxc_dTs"], s["band_output"], o["olrb"], o["dolrb_dTs"], taps=taps, **kw)
    o.update(t)
    return o


def alloc_sw_outputs(ncol, nlay):
    o = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
    for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp",
              "cotnhp", "cotnmp", "cotnlp"):
        o[k] = np.zeros(ncol)
    for k in ("fswband", "drband", "dfband"):
        o[k] = np.zeros((ncol, 14), order="F")
    o["clearCounts"] = np.zeros((ncol, 4), dtype=np.int32, order="F")
    return o


def run_sw(s, rpart=0, isolvar=0, iceflg=3, liqflg=1, iaer=10, normFlx=1, do_drfband=False, taps=(), out=None,
           bndscl=None, indsolvar=None, solcycfrac=None, radval=False, **kw):
    """rrtmg_sw on a synthetic.make_columns state; returns the output dict (+ taps).
    radval=True: the SOLAR_RADVAL build, the result gains "radval" (ncol,120)."""
    o = out if out is not None else alloc_sw_outputs(s["ncol"], s["nlay"])
    if radval:
        o.setdefault("radval", np.zeros((s["ncol"], NRADVAL), order="F"))
        kw["radval"] = o["radval"]
    t = rrtmg_sw(rpart, s["ncol"], s["nlay"], s["scon"], s["adjes"], s["coszen"], isolvar, s["play"], s["plev"],
                 s["tlay"], s["h2ovmr"], s["o3vmr"], s["co2vmr"], s["ch4vmr"], s["o2vmr"], iceflg, liqflg,
                 s["cldf"], s["ciwp"], s["clwp"], s["rei"], s["rel"], s["dyofyr"], s["zm"], s["alat"], iaer,
                 s["tauaer_sw"], s["ssaaer"], s["asmaer"], s["asdir"], s["asdif"], s["aldir"], s["aldif"],
                 s["cloudLM"], s["cloudMH"], normFlx, o["clearCounts"], o["swuflx"], o["swdflx"], o["swuflxc"],
                 o["swdflxc"], o["nirr"], o["nirf"], o["parr"], o["parf"], o["uvrr"], o["uvrf"], o["fswband"],
                 o["cotdtp"], o["cotdhp"], o["cotdmp"], o["cotdlp"], o["cotntp"], o["cotnhp"], o["cotnmp"],
                 o["cotnlp"], do_drfband, o["drband"], o["dfband"], bndscl, indsolvar, solcycfrac, taps=taps, **kw)
    o.update(t)
    return o
