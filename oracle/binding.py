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
               [(n, _dp) for n in ("ciwpmc", "clwpmc", "taug", "pfracs", "taucmc", "pwvcm", "ssi")]


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
             taps=(), bndscl=None, indsolvar=None, solcycfrac=None):
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
