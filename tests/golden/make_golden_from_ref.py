#!/usr/bin/env python
"""Golden vectors from the REFERENCE itself: builds oracle/_ref (oracle/build_ref.sh: the unmodified Fortran RRTMG LW +
SW + McICA from /root/reference with MAPL/ESMF stand-ins) and runs it on the seeded synthetic columns of
geosradiation_gridcomp_b200.synthetic, writing tests/golden/rrtmg_ref_golden_L72.npz.  tests/test_ref_pin_cpu.py then
holds the C restatement (oracle/*.c) to these numbers, which is what turns "parity unpinned" into a pin.

Exit status 3 when no Fortran compiler exists (this image, and every GPU box seen so far), 4 without a reference
tree; nothing is written then.  Usage:  python tests/golden/make_golden_from_ref.py [ncol] [nlay]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libgeosref.so")
OUT = os.path.join(ROOT, "tests", "golden", "rrtmg_ref_golden_L72.npz")


def build():
    if os.path.exists(LIB):
        return 0
    return subprocess.call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def run(ncol=48, nlay=72, seed=20260118, ih=1):
    from geosradiation_gridcomp_b200.synthetic import make_columns
    L = C.CDLL(LIB)
    L.ref_real_bytes.restype = C.c_int
    rb = L.ref_real_bytes()
    rk = np.float64 if rb == 8 else np.float32
    creal = C.c_double if rb == 8 else C.c_float
    L.ref_init(C.c_int(ih))
    s = make_columns(ncol, nlay, seed=seed)
    f = lambda k: np.asfortranarray(s[k], dtype=rk)
    z = lambda *sh: np.zeros(sh, dtype=rk, order="F")
    i = C.c_int
    out = {"ncol": ncol, "nlay": nlay, "seed": seed, "ih": ih, "real_bytes": rb}
    # ---- LW (LW/src/rrtmg_lw_rad.F90:15-23)
    lw_in = [f(k) for k in ("play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr",
                            "o2vmr", "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel")]
    taua, zm, alat = f("tauaer_lw"), f("zm"), f("alat")
    cc = np.zeros((ncol, 4), dtype=np.int32, order="F")
    fl = {k: z(ncol, nlay + 1) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")}
    bo = np.ascontiguousarray(s["band_output"], dtype=np.int32)
    olrb, dolrb = z(16, ncol), z(16, ncol)
    L.ref_rrtmg_lw(i(ncol), i(nlay), i(4), i(1), *[P(a) for a in lw_in], i(3), i(1), P(taua), P(zm), P(alat),
                   i(int(s["dyofyr"])), i(int(s["cloudLM"])), i(int(s["cloudMH"])), P(cc),
                   *[P(fl[k]) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")], P(bo), P(olrb), P(dolrb))
    out.update({"lw_" + k: v for k, v in fl.items()})
    out.update(lw_clearCounts=cc.copy(), lw_olrb=olrb, lw_dolrb_dTs=dolrb)
    # ---- SW (SW/src/rrtmg_sw_rad.F90:68-124), default options of GEOS: isolvar 0 style scalars from the synthetic state
    sw_prof = {k: z(ncol, nlay + 1) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
    sw_sfc = {k: z(ncol) for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf")}
    fswband, cot, drb, dfb = z(ncol, 14), z(ncol, 8), z(ncol, 14), z(ncol, 14)
    cc2 = np.zeros((ncol, 4), dtype=np.int32, order="F")
    bnd, ind = np.ones(14, dtype=rk), np.ones(2, dtype=rk)
    L.ref_rrtmg_sw.restype = C.c_int
    rc = L.ref_rrtmg_sw(
        i(2), i(ncol), i(nlay), creal(float(s["scon"])), creal(float(s["adjes"])), P(f("coszen")), i(0),
        P(f("play")), P(f("plev")), P(f("tlay")), P(f("h2ovmr")), P(f("o3vmr")), P(f("co2vmr")), P(f("ch4vmr")), P(f("o2vmr")),
        i(3), i(1), P(f("cldf")), P(f("ciwp")), P(f("clwp")), P(f("rei")), P(f("rel")), i(int(s["dyofyr"])), P(zm), P(alat),
        i(10), P(f("tauaer_sw")), P(f("ssaaer")), P(f("asmaer")), P(f("asdir")), P(f("asdif")), P(f("aldir")), P(f("aldif")),
        i(int(s["cloudLM"])), i(int(s["cloudMH"])), i(1), P(cc2),
        *[P(sw_prof[k]) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")],
        *[P(sw_sfc[k]) for k in ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf")], P(fswband), P(cot), i(1), P(drb), P(dfb),
        i(0), P(bnd), P(ind), creal(0.0))
    out.update({"sw_" + k: v for k, v in sw_prof.items()})
    out.update({"sw_" + k: v for k, v in sw_sfc.items()})
    out.update(sw_rc=rc, sw_clearCounts=cc2, sw_fswband=fswband, sw_cot=cot, sw_drband=drb, sw_dfband=dfb)
    return out


if __name__ == "__main__":
    rc = build()
    if rc:
        sys.exit(rc)
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 72
    res = run(ncol, nlay)
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, {k: getattr(v, "shape", v) for k, v in res.items()})
