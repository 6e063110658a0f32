"""The Run-phase glue of the two GEOS drivers around RRTMG, executed from the REFERENCE'S OWN LINES.

The glue is not a routine of its own in the reference: it is a stretch of statements inside LW_Driver
(GEOS_IrradGridComp.F90) and SORADCORE (GEOS_SolarGridComp.F90), between MAPL calls.  Here those line ranges are READ
from the files where they lie under /root/reference at run time, wrapped into a subroutine whose only own text is
the dummy-argument list and the declarations the statements need (the reference declares the same names in its
5 000-line routines), and executed through oracle/refexec/f90py.py.  Nothing of the reference is stored in this file.

Test infrastructure: tests/test_refexec_pin_cpu.py holds oracle/glue.c (and through it csrc/glue.cuh) to what these
lines compute.  Build container only (needs the reference tree).

  irrad_prepare  IRR:3238-3371  flip, TLEV, water paths, radius limits, unit conversions, absorption aerosol, ZM, negatives
  irrad_finish   IRR:3487-3533  clear counts -> cloud fractions, unflip, sign convention, SFCEM
  solar_prepare  SOL:6116-6219  aerosol normalisation, DPR, water paths, radius limits, TLEV, flips, conversions, ZL, aerosols
  solar_finish   SOL:6395-6454  unflip, cloud fractions, COT ratios with MAPL_UNDEF, FSW / FSC / FSWU / FSCU
  irrad_update   IRR:3604, 3606, 3861, 3932-3992  the between-refresh linear update of the LW exports (USE_RRTMG branch)
  heating_rates  RAD:801-802, 811, 813-814  DTDT, RADLW, RADSW of the parent component (GEOS_RadiationGridComp.F90)
"""
import os

import numpy as np

from . import f90py
from .f90py import FA

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
IRR = os.path.join(REF, "GEOSirrad_GridComp/GEOS_IrradGridComp.F90")
SOL = os.path.join(REF, "GEOSsolar_GridComp/GEOS_SolarGridComp.F90")
RAD = os.path.join(REF, "GEOS_RadiationGridComp.F90")

_ns = None


def _lines(path, first, last, starts, ends):
    """Lines first..last of a reference file; the first and the last one must be the statements this module was
    written against (a different checkout of the reference is reported, not silently mis-cut)."""
    with open(path, errors="replace") as f:
        ls = f.readlines()[first - 1:last]
    if not (ls and ls[0].strip().startswith(starts) and ls[-1].strip().startswith(ends)):
        raise RuntimeError(f"{path}:{first}-{last} is not the expected fragment ({starts!r} ... {ends!r})")
    return "".join(ls)


_IRR_PREP_HEAD = """
module refglue_irr_prep
contains
subroutine irr_prep(IM, JM, LM, LCLDMH, LCLDLM, LIQFLGLW, ICEFLGLW, KLIQUID, KICE, MAPL_AIRMW, MAPL_H2OMW, MAPL_O3MW, &
      MAPL_RGAS, MAPL_GRAV, CO2_FIXED, O2, CCL4, TS, EMIS, LATS, T2M, PLE, PL, T, Q, O3, CH4, N2O, CO2_3d, CFC11, CFC12, &
      HCFC22, FCLD, CWC, REFF, TAUA, SSAA, TSFC, EMISS, ALAT, CLIQWP, CICEWP, RELIQ, REICE, PLE_R, TLEV_R, PL_R, T_R, Q_R, &
      O3_R, CH4_R, N2O_R, CO2_R, O2_R, CCL4_R, CFC11_R, CFC12_R, CFC22_R, FCLD_R, TAUAER, ZM_R)
   integer, intent(in) :: IM, JM, LM, LIQFLGLW, ICEFLGLW, KLIQUID, KICE
   integer, intent(inout) :: LCLDMH, LCLDLM
   real, intent(in) :: MAPL_AIRMW, MAPL_H2OMW, MAPL_O3MW, MAPL_RGAS, MAPL_GRAV, CO2_FIXED, O2, CCL4
   real, intent(in) :: TS(IM,JM), EMIS(IM,JM), LATS(IM,JM), T2M(IM,JM), PLE(IM,JM,0:LM)
   real, intent(in), dimension(IM,JM,LM) :: PL, T, Q, O3, CH4, N2O, CFC11, CFC12, HCFC22, FCLD
   real, pointer :: CO2_3d(:,:,:)
   real, intent(in) :: CWC(IM,JM,LM,4), REFF(IM,JM,LM,4), TAUA(IM,JM,LM,16), SSAA(IM,JM,LM,16)
   real, intent(out) :: TSFC(IM*JM), EMISS(IM*JM,16), ALAT(IM*JM)
   real, intent(out), dimension(IM*JM,LM) :: CLIQWP, CICEWP, RELIQ, REICE, PL_R, T_R, Q_R, O3_R, CH4_R, N2O_R, CO2_R, O2_R, &
      CCL4_R, CFC11_R, CFC12_R, CFC22_R, FCLD_R, ZM_R
   real, intent(out) :: PLE_R(IM*JM,0:LM), TLEV_R(IM*JM,0:LM), TAUAER(IM*JM,LM,16)
   integer :: I, J, K, IJ, LV
   real :: xx, DP(LM), TLEV(LM+1)
"""
_IRR_FIN_HEAD = """
module refglue_irr_fin
contains
subroutine irr_fin(IM, JM, LM, NGPTLW, nRATS, CLEARCOUNTS, UFLX, DFLX, UFLXC, DFLXC, DUFLX_DTS, DUFLXC_DTS, EMIS, &
      CLDTTLW, CLDHILW, CLDMDLW, CLDLOLW, FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC, SFCEM_INT)
   integer, intent(in) :: IM, JM, LM, NGPTLW, nRATS, CLEARCOUNTS(IM*JM,4)
   real, intent(in), dimension(IM*JM,LM+1) :: UFLX, DFLX, UFLXC, DFLXC, DUFLX_DTS, DUFLXC_DTS
   real, intent(in) :: EMIS(IM,JM)
   real, pointer, dimension(:,:) :: CLDTTLW, CLDHILW, CLDMDLW, CLDLOLW
   real, intent(out), dimension(IM,JM,0:LM) :: FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC
   real, intent(out) :: SFCEM_INT(IM,JM)
   integer :: I, J, K, IJ, LV
   real :: UFLXRAT(1,1,1), DFLXRAT(1,1,1), DUFLX_DT_RAT(1,1,1), EMISS(1,1)
   real :: FLXU_INT_RAT(1,1,0:1,1), FLXD_INT_RAT(1,1,0:1,1), DFDTS_RAT(1,1,0:1,1), SFCEM_INT_RAT(1,1,1)
"""
_SOL_PREP_HEAD = """
module refglue_sol_prep
contains
subroutine sol_prep(NCOL, LM, num_aero_vars, ICEFLGSW, LIQFLGSW, MAPL_AIRMW, MAPL_H2OMW, MAPL_O3MW, MAPL_RGAS, MAPL_GRAV, &
      CO2, O2, DIST, PLE, PL, T, Q, O3, CH4, CL, QQ3, RR3, TS, TAUA, SSAA, ASYA, CICEWP, CLIQWP, REICE, RELIQ, PLE_R, TLEV_R, &
      PL_R, T_R, Q_R, O3_R, CH4_R, CO2_R, O2_R, FCLD_R, ZL_R, TAUAER, SSAAER, ASMAER, ADJES)
   integer, intent(in) :: NCOL, LM, num_aero_vars, ICEFLGSW, LIQFLGSW
   real, intent(in) :: MAPL_AIRMW, MAPL_H2OMW, MAPL_O3MW, MAPL_RGAS, MAPL_GRAV, CO2, O2, DIST
   real, intent(in) :: PLE(NCOL,LM+1), TS(NCOL), QQ3(NCOL,LM,4), RR3(NCOL,LM,4)
   real, intent(in), dimension(NCOL,LM) :: PL, T, Q, O3, CH4, CL
   real, intent(inout), dimension(NCOL,LM,14) :: TAUA, SSAA, ASYA
   real, intent(out), dimension(NCOL,LM) :: CICEWP, CLIQWP, REICE, RELIQ, PL_R, T_R, Q_R, O3_R, CH4_R, CO2_R, O2_R, FCLD_R, ZL_R
   real, intent(out) :: PLE_R(NCOL,LM+1), TLEV_R(NCOL,LM+1), TAUAER(NCOL,LM,14), SSAAER(NCOL,LM,14), ASMAER(NCOL,LM,14)
   real, intent(out) :: ADJES
   integer :: k
   real :: DPR(NCOL,LM), TLEV(NCOL,LM+1)
"""
_SOL_FIN_HEAD = """
module refglue_sol_fin
contains
subroutine sol_fin(NCOL, LM, NGPTSW, include_aerosols, MAPL_UNDEF, CLEARCOUNTS, SWUFLX, SWDFLX, SWUFLXC, SWDFLXC, &
      COTNTP, COTDTP, COTNHP, COTDHP, COTNMP, COTDMP, COTNLP, COTDLP, CLDTS, CLDHS, CLDMS, CLDLS, COTTP, COTHP, COTMP, COTLP, &
      FSW, FSC, FSWU, FSCU)
   integer, intent(in) :: NCOL, LM, NGPTSW, CLEARCOUNTS(NCOL,4)
   logical, intent(in) :: include_aerosols
   real, intent(in) :: MAPL_UNDEF
   real, intent(in), dimension(NCOL,LM+1) :: SWUFLX, SWDFLX, SWUFLXC, SWDFLXC
   real, intent(in), dimension(NCOL) :: COTNTP, COTDTP, COTNHP, COTDHP, COTNMP, COTDMP, COTNLP, COTDLP
   real, intent(out), dimension(NCOL) :: CLDTS, CLDHS, CLDMS, CLDLS, COTTP, COTHP, COTMP, COTLP
   real, intent(out), dimension(NCOL,LM+1) :: FSW, FSC, FSWU, FSCU
   real, dimension(NCOL,LM+1) :: SWUFLXR, SWDFLXR, SWUFLXCR, SWDFLXCR
"""


_IRR_UPD_HEAD = """
module refglue_irr_upd
contains
subroutine irr_upd(IM, JM, LM, MAPL_UNDEF, TSINST, TS_INT, FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC, SFCEM_INT, &
      CLDTT, FLX, FLC, FLXU, FLCU, FLXD, FLCD, OLR, OLC, SFCEM, LWS, LCS, FLNS, FLNSC, DSFDTS, OLCC5, LCSC5, &
      FLXA, FLA, FLXAU, FLAU, FLXAD, FLAD, OLRA, OLA, LWSA, LAS, FLNSNA, FLNSA)
   integer, intent(in) :: IM, JM, LM
   real, intent(in) :: MAPL_UNDEF, TSINST(IM,JM), TS_INT(IM,JM), SFCEM_INT(IM,JM), CLDTT(IM,JM)
   real, intent(in), dimension(IM,JM,0:LM) :: FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC
   real, pointer, dimension(:,:,:) :: FLX, FLC, FLXU, FLCU, FLXD, FLCD, FLXA, FLA, FLXAU, FLAU, FLXAD, FLAD
   real, pointer, dimension(:,:) :: OLR, OLC, SFCEM, LWS, LCS, FLNS, FLNSC, DSFDTS, OLCC5, LCSC5, OLRA, OLA, LWSA, LAS, &
      FLNSNA, FLNSA
   integer :: K
   real :: DELT(IM,JM), FLX_INT(IM,JM,0:LM), FLC_INT(IM,JM,0:LM)
"""


_RAD_HR_HEAD = """
module refglue_rad_hr
contains
subroutine rad_hr(IM, JM, LM, MAPL_GRAV, MAPL_CP, PLE, FLW, FSW, DTDT, RADLW, RADSW)
   integer, intent(in) :: IM, JM, LM
   real, intent(in) :: MAPL_GRAV, MAPL_CP
   real, intent(in), dimension(IM,JM,0:LM) :: PLE, FLW, FSW
   real, pointer, dimension(:,:,:) :: DTDT, RADLW, RADSW
   real :: DMI(IM,JM,LM)
"""


def available():
    return os.path.isfile(IRR) and os.path.isfile(SOL) and os.path.isfile(RAD)


def source_text():
    """The four synthetic modules: own header + the reference's lines + own `end` statements."""
    parts = [
        _IRR_PREP_HEAD + _lines(IRR, 3238, 3371, "LCLDMH = LM - LCLDMH + 1", "WHERE (FCLD_R < 0.) FCLD_R = 0.")
        + "end subroutine irr_prep\nend module refglue_irr_prep\n",
        _IRR_FIN_HEAD + _lines(IRR, 3487, 3533, "IJ = 0", "enddo ! JM") + "end subroutine irr_fin\nend module refglue_irr_fin\n",
        _SOL_PREP_HEAD + _lines(SOL, 6116, 6219, "if (num_aero_vars > 0) then", "ASMAER(:,1:LM,:) = ASYA(:,LM:1:-1,:)")
        + "end subroutine sol_prep\nend module refglue_sol_prep\n",
        _SOL_FIN_HEAD + _lines(SOL, 6395, 6454, "SWUFLXR (:,1:LM+1) = SWUFLX (:,LM+1:1:-1)", "FSCU = SWUFLXCR")
        + "end subroutine sol_fin\nend module refglue_sol_fin\n",
        # the between-refresh update of the LW exports: net fluxes (:3604, :3606), DELT (:3861), the USE_RRTMG branch
        _IRR_UPD_HEAD + _lines(IRR, 3604, 3604, "FLX_INT  = FLXD_INT  + FLXU_INT", "FLX_INT") + _lines(IRR, 3606, 3606, "FLC_INT  = FLCD_INT  + FLCU_INT", "FLC_INT")
        + _lines(IRR, 3861, 3861, "DELT = TSINST - TS_INT", "DELT") + _lines(IRR, 3932, 3992, "do K = 0, LM", "if(associated(FLNSA )) FLNSA  = MAPL_UNDEF")
        + "end subroutine irr_upd\nend module refglue_irr_upd\n",
        # the parent component's heating rates: total tendency (:801-802), DMI (:811), RADLW / RADSW (:813-814)
        _RAD_HR_HEAD + _lines(RAD, 801, 802, "if( associated (DTDT    ) ) DTDT     = (", "(FSW(:,:,0:LM-1) - FSW(:,:,1:LM)) ) * (MAPL_GRAV/MAPL_CP)")
        + _lines(RAD, 811, 811, "DMI = MAPL_GRAV/(MAPL_CP*(PLE(:,:,1:LM)-PLE(:,:,0:LM-1)))", "DMI")
        + _lines(RAD, 813, 814, "if( associated (RADLW   ) ) RADLW", "if( associated (RADSW   ) ) RADSW")
        + "end subroutine rad_hr\nend module refglue_rad_hr\n",
    ]
    return parts


def namespace():
    global _ns
    if _ns is None:
        prog = f90py.Program()
        for i, text in enumerate(source_text()):
            prog.add_source(text, f"<refglue {i}>", ())
        src = f90py.Translator(prog).generate()
        ns = dict(f90py.RUNTIME)
        exec(compile(src, "<f90py refglue>", "exec"), ns)
        ns["_init_all"]()
        ns["__source__"] = src
        _ns = ns
    return _ns


def _quiet(fn):
    """WHERE constructs evaluate their right-hand sides on the whole array before masking (numpy semantics of the
    translation): the divisions by zero under a false mask are not errors."""
    def wrapped(*a, **k):
        with np.errstate(divide="ignore", invalid="ignore"):
            return fn(*a, **k)
    wrapped.__doc__ = fn.__doc__
    return wrapped


def _f(a, shape=None, lb=None):
    a = np.array(a, dtype=np.float64, order="F")
    if shape is not None:
        a = a.reshape(shape, order="F")
    return FA(a, lb)


def _z(shape, lb=None, dtype=np.float64):
    return FA(np.zeros(shape, dtype=dtype, order="F"), lb)


@_quiet
def irrad_prepare(n, iceflg=3, liqflg=1):
    """IRR:3238-3371 on a native state of geosradiation_gridcomp_b200.synthetic.make_native_state (IM = ncol, JM = 1).
    Returns the rrtmg_lw argument arrays under oracle.binding.irrad_prepare's names."""
    ns = namespace()
    nc, lm = int(n["ncol"]), int(n["lm"])
    g3 = lambda k: _f(n[k], (nc, 1, lm))
    g2 = lambda k: _f(n[k], (nc, 1))
    cwc = np.zeros((nc, 1, lm, 4), order="F")
    reff = np.zeros((nc, 1, lm, 4), order="F")
    KICE, KLIQUID = 1, 2
    cwc[:, 0, :, KICE - 1], cwc[:, 0, :, KLIQUID - 1] = n["qice"], n["qliq"]
    reff[:, 0, :, KICE - 1], reff[:, 0, :, KLIQUID - 1] = n["rice"], n["rliq"]
    has_aer = n.get("taua_lw") is not None
    taua = np.asarray(n["taua_lw"]).reshape(nc, 1, lm, 16, order="F") if has_aer else np.zeros((nc, 1, lm, 16))
    ssaa = np.asarray(n["ssaa_lw"]).reshape(nc, 1, lm, 16, order="F") if has_aer else np.zeros((nc, 1, lm, 16))
    co2 = n.get("co2")
    out = {k: _z((nc, lm)) for k in ("clwp", "ciwp", "rel", "rei", "play", "tlay", "h2ovmr", "o3vmr", "ch4vmr", "n2ovmr",
                                     "co2vmr", "o2vmr", "ccl4vmr", "cfc11vmr", "cfc12vmr", "cfc22vmr", "cldf", "zm")}
    out.update(tsfc=_z((nc,)), emis=_z((nc, 16)), alat=_z((nc,)), plev=_z((nc, lm + 1), (1, 0)), tlev=_z((nc, lm + 1), (1, 0)),
               tauaer_lw=_z((nc, lm, 16)))
    r = ns["P_refglue_irr_prep__irr_prep"](
        nc, 1, lm, int(n["lcldmh"]), int(n["lcldlm"]), int(liqflg), int(iceflg), KLIQUID, KICE, float(n["airmw"]),
        float(n["h2omw"]), float(n["o3mw"]), float(n["rgas"]), float(n["grav"]), float(n["co2_fixed"]), float(n["o2"]),
        float(n["ccl4"]), g2("ts"), g2("emis"), g2("lats"), g2("t2m"), FA(np.array(n["ple"], order="F").reshape(nc, 1, lm + 1, order="F"), (1, 1, 0)),
        g3("pl"), g3("t"), g3("q"), g3("o3"), g3("ch4"), g3("n2o"), (g3("co2") if co2 is not None else None), g3("cfc11"),
        g3("cfc12"), g3("hcfc22"), g3("fcld"), FA(cwc), FA(reff), FA(np.asfortranarray(taua)), FA(np.asfortranarray(ssaa)),
        out["tsfc"], out["emis"], out["alat"], out["clwp"], out["ciwp"], out["rel"], out["rei"], out["plev"], out["tlev"],
        out["play"], out["tlay"], out["h2ovmr"], out["o3vmr"], out["ch4vmr"], out["n2ovmr"], out["co2vmr"], out["o2vmr"],
        out["ccl4vmr"], out["cfc11vmr"], out["cfc12vmr"], out["cfc22vmr"], out["cldf"], out["tauaer_lw"], out["zm"])
    res = {k: v.a for k, v in out.items()}
    # the scalars come back first in the generated function's result (LCLDMH, LCLDLM reversed in place, IRR:3238-3239)
    scal = [x for x in (r if isinstance(r, tuple) else (r,)) if isinstance(x, (int, np.integer))]
    res["cloudMH"], res["cloudLM"] = int(scal[0]), int(scal[1])
    return res


@_quiet
def irrad_finish(n, o):
    """IRR:3487-3533 on the outputs `o` of an rrtmg_lw call (oracle.binding.rrtmg_lw's dictionary)."""
    ns = namespace()
    nc, lm = int(n["ncol"]), int(n["lm"])
    fl = {k: _z((nc, 1, lm + 1), (1, 1, 0)) for k in ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc")}
    cf = {k: _z((nc, 1)) for k in ("cldtt", "cldhi", "cldmd", "cldlo")}
    sfc = _z((nc, 1))
    ns["P_refglue_irr_fin__irr_fin"](
        nc, 1, lm, 140, 0, FA(np.asfortranarray(o["clearCounts"], dtype=np.int64)),
        *[_f(o[k]) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")], _f(n["emis"], (nc, 1)),
        cf["cldtt"], cf["cldhi"], cf["cldmd"], cf["cldlo"], fl["flxu"], fl["flxd"], fl["flcu"], fl["flcd"], fl["dfdts"],
        fl["dfdtsc"], sfc)
    res = {k: v.a[:, 0, :] for k, v in fl.items()}
    res.update({k: v.a[:, 0] for k, v in cf.items()})
    res["sfcem"] = sfc.a[:, 0]
    return res


@_quiet
def solar_prepare(n, iceflg=3, liqflg=1):
    """SOL:6116-6219 on a native state (the Solar driver works on packed 1-D columns already)."""
    ns = namespace()
    nc, lm = int(n["ncol"]), int(n["lm"])
    qq3 = np.zeros((nc, lm, 4), order="F")
    rr3 = np.zeros((nc, lm, 4), order="F")
    qq3[:, :, 0], qq3[:, :, 1], rr3[:, :, 0], rr3[:, :, 1] = n["qice"], n["qliq"], n["rice"], n["rliq"]
    has_aer = n.get("taua_sw") is not None
    aer = [(_f(n[k]) if has_aer else _z((nc, lm, 14))) for k in ("taua_sw", "ssaa_sw", "asya_sw")]
    out = {k: _z((nc, lm)) for k in ("ciwp", "clwp", "rei", "rel", "play", "tlay", "h2ovmr", "o3vmr", "ch4vmr", "co2vmr",
                                     "o2vmr", "cld", "zm")}
    out.update(plev=_z((nc, lm + 1)), tlev=_z((nc, lm + 1)), tauaer=_z((nc, lm, 14)), ssaaer=_z((nc, lm, 14)),
               asmaer=_z((nc, lm, 14)))
    r = ns["P_refglue_sol_prep__sol_prep"](
        nc, lm, 14 if has_aer else 0, int(iceflg), int(liqflg), float(n["airmw"]), float(n["h2omw"]), float(n["o3mw"]),
        float(n["rgas"]), float(n["grav"]), float(n["co2_fixed"]), float(n["o2"]), float(n["dist"]), _f(n["ple"]), _f(n["pl"]),
        _f(n["t"]), _f(n["q"]), _f(n["o3"]), _f(n["ch4"]), _f(n["fcld"]), FA(qq3), FA(rr3), _f(n["ts"]), *aer,
        out["ciwp"], out["clwp"], out["rei"], out["rel"], out["plev"], out["tlev"], out["play"], out["tlay"], out["h2ovmr"],
        out["o3vmr"], out["ch4vmr"], out["co2vmr"], out["o2vmr"], out["cld"], out["zm"], out["tauaer"], out["ssaaer"],
        out["asmaer"], 0.0)
    res = {k: v.a for k, v in out.items()}
    res["adjes"] = float([x for x in (r if isinstance(r, tuple) else (r,)) if isinstance(x, float)][-1])
    return res


@_quiet
def solar_finish(n, o):
    """SOL:6395-6454 on the outputs `o` of an rrtmg_sw call (oracle.binding.rrtmg_sw's dictionary)."""
    ns = namespace()
    nc, lm = int(n["ncol"]), int(n["lm"])
    v1 = {k: _z((nc,)) for k in ("cldts", "cldhs", "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp")}
    v2 = {k: _z((nc, lm + 1)) for k in ("fsw", "fsc", "fswu", "fscu")}
    ns["P_refglue_sol_fin__sol_fin"](
        nc, lm, 112, True, float(n["undef"]), FA(np.asfortranarray(o["clearCounts"], dtype=np.int64)),
        *[_f(o[k]) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")],
        *[_f(o[k]) for k in ("cotntp", "cotdtp", "cotnhp", "cotdhp", "cotnmp", "cotdmp", "cotnlp", "cotdlp")],
        *[v1[k] for k in ("cldts", "cldhs", "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp")],
        *[v2[k] for k in ("fsw", "fsc", "fswu", "fscu")])
    res = {k: v.a for k, v in {**v1, **v2}.items()}
    return res


@_quiet
def irrad_update(f, ts_int, tsinst, undef=1e15, cldtt=None):
    """IRR:3604, 3606, 3861, 3932-3992 on the refresh outputs `f` (irrad_finish's dictionary): every export of the
    USE_RRTMG branch that is not a no-aerosol diagnostic (those are MAPL_UNDEF fills and stay unassociated)."""
    ns = namespace()
    nc, lm1 = f["flxu"].shape
    lm = lm1 - 1
    g3 = lambda k: FA(np.array(f[k], dtype=np.float64, order="F").reshape(nc, 1, lm1, order="F"), (1, 1, 0))
    o3 = {k: _z((nc, 1, lm1), (1, 1, 0)) for k in ("flx", "flc", "flxu", "flcu", "flxd", "flcd")}
    o2 = {k: _z((nc, 1)) for k in ("olr", "olc", "sfcem", "lws", "lcs", "flns", "flnsc", "dsfdts", "olcc5", "lcsc5")}
    ct = _f(cldtt if cldtt is not None else np.zeros(nc), (nc, 1))
    ns["P_refglue_irr_upd__irr_upd"](
        nc, 1, lm, float(undef), _f(tsinst, (nc, 1)), _f(ts_int, (nc, 1)), g3("flxu"), g3("flxd"), g3("flcu"), g3("flcd"),
        g3("dfdts"), g3("dfdtsc"), _f(f["sfcem"], (nc, 1)), ct,
        *[o3[k] for k in ("flx", "flc", "flxu", "flcu", "flxd", "flcd")],
        *[o2[k] for k in ("olr", "olc", "sfcem", "lws", "lcs", "flns", "flnsc", "dsfdts", "olcc5", "lcsc5")],
        *([None] * 12))
    res = {k: v.a[:, 0, :] for k, v in o3.items()}
    res.update({k: v.a[:, 0] for k, v in o2.items()})
    return res


def heating_rates(ple, flw, fsw, grav, cp):
    """RAD:801-802, 811, 813-814 on native arrays (ncol, 0:LM), level 0 at the model top: PLE [Pa], the net LW and SW
    fluxes FLW, FSW [W/m2, downward positive] as the parent component holds them.  Returns DTDT [K Pa / s], RADLW and
    RADSW [K/s], each (ncol, LM)."""
    ns = namespace()
    nc, lm1 = np.shape(ple)
    lm = lm1 - 1
    g3 = lambda a: FA(np.array(a, dtype=np.float64, order="F").reshape(nc, 1, lm1, order="F"), (1, 1, 0))
    out = [_z((nc, 1, lm)) for _ in range(3)]
    ns["P_refglue_rad_hr__rad_hr"](nc, 1, lm, float(grav), float(cp), g3(ple), g3(flw), g3(fsw), *out)
    return {k: v.a[:, 0, :] for k, v in zip(("dtdt", "radlw", "radsw"), out)}


# ---- a whole refresh of a driver from the reference's text: glue lines -> RRTMG sources -> glue lines -----------------
def irrad_refresh(n, iceflg=3, liqflg=1):
    """LW_Driver's RRTMG branch end to end (IRR:3238-3371, rrtmg_lw of LW/src through refexec.run, IRR:3487-3533):
    returns (prepared rrtmg_lw arguments, rrtmg_lw outputs, native exports)."""
    from . import run
    s = irrad_prepare(n, iceflg, liqflg)
    s.update(ncol=int(n["ncol"]), nlay=int(n["lm"]), dyofyr=int(n["doy"]), band_output=n["band_output"])
    o = run.rrtmg_lw(s, iceflg=iceflg, liqflg=liqflg)
    f = irrad_finish(n, o)
    f["olrb"], f["dolrb_dts"] = o["olrb"], o["dolrb_dTs"]
    return s, o, f


def solar_refresh(n, iceflg=3, liqflg=1, isolvar=0):
    """SORADCORE's RRTMG branch end to end (SOL:6116-6219, rrtmg_sw of SW/src through refexec.run with the scalars of
    SOL:6234-6240, 6345-6347: iaer = 10, normFlx = 1, the albedo and super-layer mapping; SOL:6395-6454)."""
    from . import run
    s = solar_prepare(n, iceflg, liqflg)
    lm = int(n["lm"])
    s.update(ncol=int(n["ncol"]), nlay=lm, dyofyr=int(n["doy"]), scon=float(n["sc"]), coszen=n["zt"], alat=n["lats"],
             cldf=s["cld"], tauaer_sw=s["tauaer"], asdir=n["albvr"], asdif=n["albvf"], aldir=n["albnr"], aldif=n["albnf"],
             cloudLM=lm - int(n["lcldlm"]) + 1, cloudMH=lm - int(n["lcldmh"]) + 1)
    o = run.rrtmg_sw(s, isolvar=isolvar, iceflg=iceflg, liqflg=liqflg, iaer=10, normFlx=1)
    f = solar_finish(n, o)
    for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband"):
        f[k] = o[k]
    return s, o, f
