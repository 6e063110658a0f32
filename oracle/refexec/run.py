"""The reference's RRTMG LW + SW + McICA, executed from its OWN SOURCE TEXT through oracle/refexec/f90py.py.

Test infrastructure (like everything under oracle/): needs /root/reference (or $REFERENCE_ROOT), so it runs in the
build container only; what travels is its output, tests/golden/rrtmg_refexec_golden_L72.npz, written by
tests/golden/make_golden_from_refexec.py.  The entry points take the synthetic state of
geosradiation_gridcomp_b200.synthetic.make_columns and return the same dictionaries as oracle/binding.py, so that
`oracle.binding.rrtmg_lw(s)` and `refexec.run.rrtmg_lw(s)` compare key by key.

Call sequence = oracle/ref_recipe/ref_capi.F90 (the wrappers a compiled oracle/_ref would use):
  init:  rrtmg_lw_ini, rrtmg_sw_ini (LW/src/rrtmg_lw_init.F90, SW/src/rrtmg_sw_init.F90), unset_inhomogeneity,
         set_inhomogeneity(ih) (SH/cloud_condensate_inhomogeneity.F90)
  LW:    rrtmg_lw  (LW/src/rrtmg_lw_rad.F90:15)
  SW:    rrtmg_sw  (SW/src/rrtmg_sw_rad.F90:68; MAPL timers are no-ops; SOLAR_RADVAL undefined, or defined in a second
         translation of the same files: rrtmg_sw(..., radval=True))
"""
import glob
import os

import numpy as np

from . import f90py
from .f90py import FA

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
LW = os.path.join(REF, "GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model")
SW = os.path.join(REF, "GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model")
SH = os.path.join(REF, "GEOS_RadiationShared")

_ns = None
_ih = None
_ns_rv = None   # the same sources translated with SOLAR_RADVAL defined (GEOSsolar_GridComp/CMakeLists.txt:18-20)
_ih_rv = None


def available():
    return os.path.isdir(os.path.join(REF, "GEOSirrad_GridComp"))


def sources():
    """Every reference file that is translated (read where it lies, nothing copied)."""
    fs = [os.path.join(SH, "cloud_condensate_inhomogeneity.F90"), os.path.join(SH, "cloud_subcol_gen.F90")]
    for root in (LW, SW):
        fs += sorted(glob.glob(os.path.join(root, "modules", "*.F90"))) + sorted(glob.glob(os.path.join(root, "src", "*.F90")))
    return fs


def namespace(ih=1, radval=False):
    """Translate once, run the reference's init routines, select the condensate inhomogeneity option.
    radval: the translation with the SOLAR_RADVAL compile-time flag (a second, independent set of module state)."""
    global _ns, _ih, _ns_rv, _ih_rv
    if radval:
        if _ns_rv is None:
            _ns_rv = f90py.load(sources(), defines=("SOLAR_RADVAL",))
            _ns_rv["P_rrtmg_lw_init__rrtmg_lw_ini"]()
            _ns_rv["P_rrtmg_sw_init__rrtmg_sw_ini"]()
        if _ih_rv != ih:
            _ns_rv["P_cloud_condensate_inhomogeneity__unset_inhomogeneity"]()
            if ih > 0:
                _ns_rv["P_cloud_condensate_inhomogeneity__set_inhomogeneity"](int(ih))
            _ih_rv = ih
        return _ns_rv
    if _ns is None:
        _ns = f90py.load(sources())
        _ns["P_rrtmg_lw_init__rrtmg_lw_ini"]()
        _ns["P_rrtmg_sw_init__rrtmg_sw_ini"]()
    if _ih != ih:
        _ns["P_cloud_condensate_inhomogeneity__unset_inhomogeneity"]()
        if ih > 0:
            _ns["P_cloud_condensate_inhomogeneity__set_inhomogeneity"](int(ih))
        _ih = ih
    return _ns


# the module variables initialize_cloud_subcol_gen sets, in its argument order (SH/cloud_subcol_gen.F90:108-129)
CORR_NAMES = ("aam1", "aam2", "aam30", "aam4", "ram1", "ram2", "ram30", "ram4")


def initialize_cloud_subcol_gen(am, ih=1):
    """SH/cloud_subcol_gen.F90 initialize_cloud_subcol_gen: the eight correlation-length parameters."""
    namespace(ih)["P_cloud_subcol_gen__initialize_cloud_subcol_gen"](*[float(x) for x in am])


def _f(a):
    return FA(np.array(a, dtype=np.float64, order="F"))


def _z(*shape):
    return FA(np.zeros(shape, dtype=np.float64, order="F"))


class _Tap:
    """Wraps procedures of the generated namespace for the duration of one driver call and records, per partition, the
    intermediate arrays the oracle's `taps` expose (oracle/binding.py Taps).  The generated code calls every procedure
    with keyword arguments, so a wrapper sees the dummies by the reference's own names."""

    def __init__(self, ns, hooks):
        self.ns, self.hooks, self.saved, self.parts = ns, hooks, {}, []

    def __enter__(self):
        for name, hook in self.hooks.items():
            self.saved[name] = orig = self.ns[name]
            self.ns[name] = (lambda orig, hook: lambda **kw: hook(orig, kw))(orig, hook)
        return self

    def __exit__(self, *exc):
        self.ns.update(self.saved)


def _cat(parts, key, axis=-1):
    return np.concatenate([p[key] for p in parts], axis=axis)


def rrtmg_lw_taps(s, psize=4, iceflg=3, liqflg=1, ih=1):
    """rrtmg_lw plus the intermediates: McICA sub-columns and cloud optical depths (after cldprmc), the setcoef module
    arrays and taumol's output (as rtrnmc receives them).  Layouts as oracle/binding.py: (ncol, nlay) for the
    per-layer arrays, (ncol, ngpt, nlay) for the per-g-point ones."""
    ns = namespace(ih)
    cl, rt = [], []

    def after_cldprmc(orig, kw):
        r = orig(**kw)
        n = kw["ncol"]
        cl.append({k: np.array(kw[k].a[..., :n]) for k in ("cldymc", "ciwpmc", "clwpmc", "taucmc")})
        return r

    def before_rtrnmc(orig, kw):
        M, n = ns["M_rrtmg_lw_setcoef"], kw["ncol"]
        d = {k: np.array(getattr(M, k).a[..., :n]) for k in ("jp", "jt", "jt1", "indself", "indfor", "indminor", "fac00",
                                                              "fac01", "fac10", "fac11", "laytrop", "pwvcm")}
        d.update(taug=np.array(kw["taug"].a[..., :n]), pfracs=np.array(kw["pfracs"].a[..., :n]))
        rt.append(d)
        return orig(**kw)

    with _Tap(ns, {"P_rrtmg_lw_cldprmc__cldprmc": after_cldprmc, "P_rrtmg_lw_rtrnmc__rtrnmc": before_rtrnmc}):
        out = rrtmg_lw(s, psize=psize, iceflg=iceflg, liqflg=liqflg, ih=ih)
    for k in ("cldymc", "ciwpmc", "clwpmc", "taucmc"):
        out[k] = np.ascontiguousarray(_cat(cl, k).T)
    out["cldymc"] = out["cldymc"].astype(np.uint8)
    for k in ("jp", "jt", "jt1", "indself", "indfor", "indminor"):
        out[k] = np.asfortranarray(_cat(rt, k).T.astype(np.int32))
    for k in ("fac00", "fac01", "fac10", "fac11"):
        out[k] = np.asfortranarray(_cat(rt, k).T)
    out["laytrop"] = _cat(rt, "laytrop").astype(np.int32)
    out["pwvcm"] = _cat(rt, "pwvcm")
    for k in ("taug", "pfracs"):
        out[k] = np.ascontiguousarray(_cat(rt, k).T)
    return out


def rrtmg_sw_taps(s, ih=1, **opts):
    """rrtmg_sw plus the intermediates of setcoef_sw, cldprmc_sw and taumol_sw, put back into the caller's column
    order: the driver runs the cloud-free columns first, then the cloudy ones (SW/src/rrtmg_sw_rad.F90:1138-1204)."""
    ns = namespace(ih)
    ncol, nlay = int(s["ncol"]), int(s["nlay"])
    sc, tm, cp = [], [], []

    def after_setcoef(orig, kw):
        r = orig(**kw)
        n = kw["ncol"]
        sc.append({k: np.array(kw[k].a[..., :n]) for k in ("jp", "jt", "jt1", "indself", "indfor", "fac00", "fac01", "fac10",
                                                           "fac11", "laytrop")})
        cp.append(None)
        return r

    def after_cldprmc(orig, kw):
        r = orig(**kw)
        n = kw["ncol"]
        cp.append({k: np.array(kw[k].a[..., :n]) for k in ("cldymc", "taucmc")})
        return r

    def after_taumol(orig, kw):
        r = orig(**kw)
        n = kw["ncol"]
        tm.append({"taug": np.array(kw["taug"].a[..., :n]), "taur": np.array(kw["taur"].a[..., :n]),
                   "sfluxzen": np.array(kw["sfluxzen"].a[..., :n]), "ssi": np.array(kw["ssi"].a[..., :n])})
        return r

    with _Tap(ns, {"P_rrtmg_sw_setcoef__setcoef_sw": after_setcoef, "P_rrtmg_sw_cldprmc__cldprmc_sw": after_cldprmc,
                   "P_rrtmg_sw_taumol__taumol_sw": after_taumol}):
        out = rrtmg_sw(s, ih=ih, **opts)
    cloudy = (np.asarray(s["cldf"]) > 0).any(axis=1)
    order = np.concatenate([np.nonzero(~cloudy)[0], np.nonzero(cloudy)[0]])   # driver order -> caller's column
    inv = np.empty(ncol, dtype=np.int64)
    inv[order] = np.arange(ncol)
    for k in ("jp", "jt", "jt1", "indself", "indfor"):
        out[k] = np.asfortranarray(_cat(sc, k)[..., inv].T.astype(np.int32))
    for k in ("fac00", "fac01", "fac10", "fac11"):
        out[k] = np.asfortranarray(_cat(sc, k)[..., inv].T)
    out["laytrop"] = _cat(sc, "laytrop")[inv].astype(np.int32)
    out["taug"] = np.ascontiguousarray(_cat(tm, "taug")[..., inv].T)
    out["taur"] = np.ascontiguousarray(_cat(tm, "taur")[..., inv].T)
    out["sfluxzen"] = np.ascontiguousarray(_cat(tm, "sfluxzen")[..., inv].T)
    out["ssi"] = np.ascontiguousarray(_cat(tm, "ssi")[..., inv].T)   # what the sweeps use when isolvar >= 0
    # cloud optics exist for the cloudy partitions only (cldprmc_sw is not called for cloud-free ones)
    ngpt = out["taug"].shape[1]
    taucmc = np.zeros((ncol, ngpt, nlay))
    cldymc = np.zeros((ncol, ngpt, nlay), dtype=np.uint8)
    done = int((~cloudy).sum())
    for part in [c for c in cp if c is not None]:
        n = part["taucmc"].shape[-1]
        cols = order[done:done + n]
        taucmc[cols] = part["taucmc"].T
        cldymc[cols] = part["cldymc"].T
        done += n
    out.update(taucmc=taucmc, cldymc=cldymc)
    return out


def rrtmg_lw(s, psize=4, dudTs=True, iceflg=3, liqflg=1, ih=1):
    ns = namespace(ih)
    ncol, nlay = int(s["ncol"]), int(s["nlay"])
    names = ("play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
             "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel")
    ins = [_f(s[k]) for k in names]
    cc = FA(np.zeros((ncol, 4), dtype=np.int64, order="F"))
    fl = {k: _z(ncol, nlay + 1) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")}
    bo = FA(np.asarray(s["band_output"]).astype(bool))
    olrb, dolrb = _z(16, ncol), _z(16, ncol)
    ns["P_rrtmg_lw_rad__rrtmg_lw"](
        ncol, nlay, int(psize), bool(dudTs), *ins, int(iceflg), int(liqflg), _f(s["tauaer_lw"]), _f(s["zm"]), _f(s["alat"]),
        int(s["dyofyr"]), int(s["cloudLM"]), int(s["cloudMH"]), cc,
        *[fl[k] for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")], bo, olrb, dolrb)
    out = {k: v.a for k, v in fl.items()}
    out.update(clearCounts=cc.a.astype(np.int32), olrb=olrb.a, dolrb_dTs=dolrb.a)
    return out


# The SOLAR_RADVAL dummies of rrtmg_sw in the order of its argument list (SW/src/rrtmg_sw_rad.F90:85-122): fifteen
# families, each <family>{d,n}{t,h,m,l}p = denominator / numerator sums of the Tot|High|Mid|Low super-layers.
RADVAL_FAMILIES = ("cds", "cotl", "cdsl", "coti", "cdsi", "ssal", "sdsl", "ssai", "sdsi", "asml", "adsl", "asmi", "adsi",
                   "forl", "fori")
RADVAL_NAMES = tuple(f + dn + lev + "p" for f in RADVAL_FAMILIES for dn in "dn" for lev in "thml")


def rrtmg_sw(s, rpart=0, isolvar=0, iceflg=3, liqflg=1, iaer=10, normFlx=1, do_drfband=False, bndscl=None,
             indsolvar=None, solcycfrac=None, ih=1, radval=False):
    """radval: call the SOLAR_RADVAL build of rrtmg_sw; the result gains "radval" (ncol, 120), columns in
    RADVAL_NAMES order."""
    ns = namespace(ih, radval)
    rv = [_z(int(s["ncol"])) for _ in RADVAL_NAMES] if radval else []
    ncol, nlay = int(s["ncol"]), int(s["nlay"])
    prof = {k: _z(ncol, nlay + 1) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
    sfc = {k: _z(ncol) for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf")}
    cot = {k: _z(ncol) for k in ("cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp")}
    fswband, drb, dfb = _z(ncol, 14), _z(ncol, 14), _z(ncol, 14)
    cc = FA(np.zeros((ncol, 4), dtype=np.int64, order="F"))
    kw = {}
    if bndscl is not None:
        kw["bndscl"] = _f(bndscl)
    if indsolvar is not None:
        kw["indsolvar"] = _f(indsolvar)
    if solcycfrac is not None:
        kw["solcycfrac"] = float(solcycfrac)
    r = ns["P_rrtmg_sw_rad__rrtmg_sw"](
        None, int(rpart), ncol, nlay, float(s["scon"]), float(s["adjes"]), _f(s["coszen"]), int(isolvar),
        _f(s["play"]), _f(s["plev"]), _f(s["tlay"]), _f(s["h2ovmr"]), _f(s["o3vmr"]), _f(s["co2vmr"]), _f(s["ch4vmr"]),
        _f(s["o2vmr"]), int(iceflg), int(liqflg), _f(s["cldf"]), _f(s["ciwp"]), _f(s["clwp"]), _f(s["rei"]), _f(s["rel"]),
        int(s["dyofyr"]), _f(s["zm"]), _f(s["alat"]), int(iaer), _f(s["tauaer_sw"]), _f(s["ssaaer"]), _f(s["asmaer"]),
        _f(s["asdir"]), _f(s["asdif"]), _f(s["aldir"]), _f(s["aldif"]), int(s["cloudLM"]), int(s["cloudMH"]), int(normFlx),
        cc, *[prof[k] for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")],
        *[sfc[k] for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf")], fswband,
        *[cot[k] for k in ("cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp")], *rv,
        bool(do_drfband), drb, dfb, rc=0, **kw)
    out = {k: v.a for k, v in {**prof, **sfc, **cot}.items()}
    if radval:
        out["radval"] = np.stack([v.a for v in rv], axis=1)
    out.update(clearCounts=cc.a.astype(np.int32), fswband=fswband.a, drband=drb.a, dfband=dfb.a, ret=r)
    return out
