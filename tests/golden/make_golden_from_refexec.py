#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN SOURCE TEXT, executed here without a Fortran compiler.

oracle/refexec/f90py.py translates the Fortran of /root/reference (RRTMG LW, RRTMG SW, McICA, condensate
inhomogeneity, NRLSSI2 - every file of SURVEY.md section 8a, read where it lies) into Python in memory, with the
reference's fp64 contract (default real promoted to 8 bytes, expressions evaluated as written, 32-bit integer
wrap-around for the KISS generator); oracle/refexec/run.py calls the reference's drivers `rrtmg_lw` / `rrtmg_sw` the
way oracle/ref_recipe/ref_capi.F90 would.  This script runs the cases of tests/golden/refexec_cases.py and writes
tests/golden/rrtmg_refexec_golden.npz (keys "<case>/lw/<output>", "<case>/sw/<output>").

The stored numbers are the pin: tests/test_refexec_pin_cpu.py holds the C restatement (oracle/*.c) to them, and
tests/test_refexec_pin_gpu.py the CUDA path, on the GPU box, where /root/reference does not exist.

Run time: about five minutes of pure-Python arithmetic.  Usage: python tests/golden/make_golden_from_refexec.py [case ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "rrtmg_refexec_golden.npz")


def run_case(name, c):
    from geosradiation_gridcomp_b200.synthetic import make_columns
    from oracle.refexec import run
    from refexec_cases import LW_OUT, LW_TAPS, SW_OUT, SW_TAPS
    s = make_columns(c["ncol"], c["nlay"], seed=c["seed"])
    out = {}
    ns = run.namespace(c["ih"])
    defaults = [float(ns["M_cloud_subcol_gen"].__dict__[k]) for k in run.CORR_NAMES]
    if c["corr"] is not None:
        run.initialize_cloud_subcol_gen(c["corr"], c["ih"])
    try:
        if c["lw"] is not None:
            r = run.rrtmg_lw_taps(s, ih=c["ih"], **c["lw"]) if c["taps"] else run.rrtmg_lw(s, ih=c["ih"], **c["lw"])
            for k in LW_OUT + (LW_TAPS if c["taps"] else ()):
                out[f"{name}/lw/{k}"] = r[k]
        if c["sw"] is not None:
            r = run.rrtmg_sw_taps(s, ih=c["ih"], **c["sw"]) if c["taps"] else run.rrtmg_sw(s, ih=c["ih"], **c["sw"])
            assert r["ret"][-1] == 0, r["ret"]
            for k in SW_OUT + (tuple(SW_TAPS) if c["taps"] else ()) + (("radval",) if c["sw"].get("radval") else ()):
                out[f"{name}/sw/{k}"] = r[k]
    finally:
        if c["corr"] is not None:
            run.initialize_cloud_subcol_gen(defaults, c["ih"])
    return out


KISS_SEEDS = ((1, 2, 3, 4), (-2147483648, 2147483647, 65535, 65536), (123456789, 362436069, 521288629, 916191069),
              (20260118, -77, 4095, -65536))
KISS_DRAWS = 600


def run_kiss():
    """The reference's own rng_kiss (SH/cloud_subcol_gen.F90:546-576) on a few seed quadruples: the seeds after every
    call are not kept, the real*8 ran_num is; the integer `kiss` is recovered exactly from it by the tests."""
    from oracle.refexec import run
    ns = run.namespace()
    out = np.zeros((len(KISS_SEEDS), KISS_DRAWS))
    for i, sd in enumerate(KISS_SEEDS):
        s1, s2, s3, s4 = sd
        for d in range(KISS_DRAWS):
            s1, s2, s3, s4, r = ns["P_cloud_subcol_gen__rng_kiss"](s1, s2, s3, s4, 0.0)
            out[i, d] = r
    return {"kiss/seeds": np.array(KISS_SEEDS, dtype=np.int64), "kiss/ran_num": out}


REFRESH_NCOL, REFRESH_SEED = 14, 73
RAD_GRAV, RAD_CP = 9.80665, 1004.68506    # MAPL_GRAV, MAPL_CP
RAD_EXPORTS = ("dtdt", "radlw", "radsw")
IRR_EXPORTS = ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc", "sfcem", "cldtt", "cldhi", "cldmd", "cldlo", "olrb", "dolrb_dts")
SOL_EXPORTS = ("fsw", "fsc", "fswu", "fscu", "cldts", "cldhs", "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp", "nirr",
               "nirf", "parr", "parf", "uvrr", "uvrf", "fswband")
IRR_PREPARED = ("play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
                "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel", "tauaer_lw", "zm", "alat")
SOL_PREPARED = ("play", "plev", "tlay", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "o2vmr", "cld", "ciwp", "clwp", "rei", "rel",
                "zm", "tauaer", "ssaaer", "asmaer")


def rad_fsw(n, fsw_normalised):
    """The net SW flux in W/m2 the parent component sees: the refresh returns fluxes normalised by the TOA insolation
    (normFlx = 1, SOL:6346) and the Solar Update scales them back by the instantaneous insolation."""
    return fsw_normalised * (float(n["sc"]) * np.asarray(n["zt"], dtype=np.float64)[:, None])


def run_refresh():
    """A whole refresh of each driver from the reference's text: the Run-phase glue LINES of GEOS_IrradGridComp.F90 /
    GEOS_SolarGridComp.F90 (oracle/refexec/glue.py) around the RRTMG sources: native state in, native exports out."""
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    from oracle.refexec import glue
    n = make_native_state(REFRESH_NCOL, 72, seed=REFRESH_SEED)
    out = {}
    s, _, f = glue.irrad_refresh(n)
    out.update({f"refresh/irr_prepared/{k}": s[k] for k in IRR_PREPARED})
    out.update({f"refresh/irr/{k}": f[k] for k in IRR_EXPORTS})
    out["refresh/irr_prepared/cloudLM"], out["refresh/irr_prepared/cloudMH"] = np.int64(s["cloudLM"]), np.int64(s["cloudMH"])
    flw = f["flxu"] + f["flxd"]            # the Irrad export FLX the parent component reads (IRR:3604: upward negative)
    s, _, f = glue.solar_refresh(n)
    out.update({f"refresh/sol_prepared/{k}": s[k] for k in SOL_PREPARED})
    out.update({f"refresh/sol/{k}": f[k] for k in SOL_EXPORTS})
    # the parent component's heating rates from the two refreshes (GEOS_RadiationGridComp.F90:801-814)
    hr = glue.heating_rates(n["ple"], flw, rad_fsw(n, f["fsw"]), RAD_GRAV, RAD_CP)
    out.update({f"refresh/rad/{k}": hr[k] for k in RAD_EXPORTS})
    return out


if __name__ == "__main__":
    from oracle.refexec import run
    from refexec_cases import CASES, INTEGER_KEYS
    if not run.available():
        print("no reference tree at", run.REF, file=sys.stderr)
        sys.exit(4)
    only = sys.argv[1:]
    res = dict(np.load(OUT)) if (only and os.path.exists(OUT)) else {}
    t0 = time.time()
    run.namespace()
    print(f"translated and initialised in {time.time() - t0:.1f} s ({len(run.sources())} reference files)")
    for name, c in CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        for k, v in run_case(name, c).items():
            v = np.asarray(v)
            if k.rsplit("/", 1)[1] in INTEGER_KEYS:
                v = v.astype(np.uint8 if k.endswith("cldymc") else np.int32)
            res[k] = v
        print(f"{name}: {time.time() - t:.1f} s")
    if not only or "kiss" in only:
        res.update(run_kiss())
    if not only or "refresh" in only:
        t = time.time()
        res.update(run_refresh())
        print(f"refresh: {time.time() - t:.1f} s")
    res["real_bytes"] = np.int64(8)
    res["reference_files"] = np.int64(len(run.sources()))
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(res), "arrays")
