"""The cases of tests/golden/rrtmg_refexec_golden.npz: inputs (a seeded synthetic state + the options of one driver
call) shared by the generator (make_golden_from_refexec.py, which runs the REFERENCE'S OWN SOURCE through
oracle/refexec) and by the tests that hold the C restatement (tests/test_refexec_pin_cpu.py) and the CUDA path
(tests/test_refexec_pin_gpu.py) to the stored numbers.

Every case: name -> dict(ncol, nlay, seed, ih, corr, lw=<kwargs or None>, sw=<kwargs or None>, taps=<bool>).
Options follow the reference's dummy arguments (LW/src/rrtmg_lw_rad.F90:15-23, SW/src/rrtmg_sw_rad.F90:68-124)."""
import numpy as np

CORR = (0.6, 3.0, 5.0, -30.0, 0.3, 1.1, 10.0, 35.0)   # non-default initialize_cloud_subcol_gen arguments


def _case(ncol, nlay=72, seed=20260118, ih=1, corr=None, lw=None, sw=None, taps=False):
    return dict(ncol=ncol, nlay=nlay, seed=seed, ih=ih, corr=corr, lw=lw, sw=sw, taps=taps)


CASES = {
    # GEOS defaults (iceflag 3, liqflag 1, isolvar 0, iaer 10, normalised fluxes); `taps` adds every intermediate the oracle taps
    "default": _case(40, lw=dict(), sw=dict(do_drfband=True)),
    "taps": _case(8, seed=3, lw=dict(), sw=dict(normFlx=0), taps=True),
    "l181": _case(8, nlay=181, seed=3, lw=dict(), sw=dict(do_drfband=True)),
    # partition sizes that do not divide the column count, unnormalised SW fluxes
    "parts": _case(11, seed=11, lw=dict(psize=5), sw=dict(rpart=3, normFlx=0)),
    # LW cloud optics options (LW/src/rrtmg_lw_cldprmc.F90:66-268)
    "lw_ice0": _case(10, seed=67, lw=dict(iceflg=0)),
    "lw_ice1": _case(10, seed=67, lw=dict(iceflg=1)),
    "lw_ice2": _case(10, seed=67, lw=dict(iceflg=2)),
    "lw_ice4": _case(10, seed=67, lw=dict(iceflg=4)),
    "lw_no_dudts": _case(6, seed=5, lw=dict(dudTs=False)),
    # SW cloud optics options (SW/src/rrtmg_sw_cldprmc.F90:150-300) and aerosol switch
    "sw_ice1": _case(10, seed=31, sw=dict(iceflg=1, normFlx=0)),
    "sw_ice2": _case(10, seed=31, sw=dict(iceflg=2, normFlx=0)),
    "sw_ice4": _case(10, seed=31, sw=dict(iceflg=4, normFlx=0)),
    "sw_noaer": _case(8, seed=31, sw=dict(iaer=0, normFlx=0)),
    # solar variability modes (SW/src/rrtmg_sw_rad.F90:880-1118, NRLSSI2)
    "sw_isolvar_m1": _case(6, seed=31, sw=dict(isolvar=-1, normFlx=0)),
    "sw_isolvar_m1_bndscl": _case(6, seed=31, sw=dict(isolvar=-1, normFlx=0, bndscl=np.linspace(0.9, 1.1, 14))),
    "sw_isolvar_1": _case(6, seed=31, sw=dict(isolvar=1, normFlx=0, solcycfrac=0.3)),
    "sw_isolvar_1_ind": _case(6, seed=31, sw=dict(isolvar=1, normFlx=0, solcycfrac=0.7, indsolvar=[1.2, 0.8])),
    "sw_isolvar_2": _case(6, seed=31, sw=dict(isolvar=2, normFlx=0, indsolvar=[0.16, 1000.0])),
    "sw_isolvar_3": _case(6, seed=31, sw=dict(isolvar=3, normFlx=0, bndscl=np.linspace(1.05, 0.95, 14))),
    # McICA: homogeneous condensate, gamma-distributed condensate, non-default decorrelation lengths
    "mcica_ih0": _case(12, seed=41, ih=0, lw=dict(), sw=dict()),
    "mcica_ih2": _case(12, seed=41, ih=2, lw=dict(), sw=dict()),
    "mcica_corr": _case(12, seed=61, corr=CORR, lw=dict(), sw=dict()),
    # the SOLAR_RADVAL build of rrtmg_sw (GEOSsolar_GridComp/CMakeLists.txt:18-20; SW/src/rrtmg_sw_rad.F90:85-122,
    # rrtmg_sw_cldprmc.F90:38-47, rrtmg_sw_spcvmc.F90:681-1105): the 120 phase-split PAR super-layer diagnostics
    "radval": _case(16, seed=31, sw=dict(radval=True, normFlx=0)),
    "radval_ice1": _case(10, seed=67, sw=dict(radval=True, iceflg=1)),
    "radval_ice2_m1": _case(10, seed=67, sw=dict(radval=True, iceflg=2, isolvar=-1)),
    "radval_ice4_ih2": _case(10, seed=41, ih=2, sw=dict(radval=True, iceflg=4)),
    "radval_l181": _case(6, nlay=181, seed=3, sw=dict(radval=True)),
}

LW_OUT = ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs", "olrb", "dolrb_dTs", "clearCounts")
SW_OUT = ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband", "drband",
          "dfband", "cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp", "clearCounts")
LW_TAPS = ("jp", "jt", "jt1", "indself", "indfor", "indminor", "laytrop", "fac00", "fac01", "fac10", "fac11", "cldymc",
           "ciwpmc", "clwpmc", "taug", "pfracs", "taucmc", "pwvcm")
# golden key -> the oracle's tap name (oracle/binding.py Taps reuses the LW field names for the SW arrays)
SW_TAPS = {"jp": "jp", "jt": "jt", "jt1": "jt1", "indself": "indself", "indfor": "indfor", "laytrop": "laytrop",
           "fac00": "fac00", "fac01": "fac01", "fac10": "fac10", "fac11": "fac11", "cldymc": "cldymc", "taucmc": "taucmc",
           "taug": "taug", "taur": "pfracs", "ssi": "ssi"}
INTEGER_KEYS = ("clearCounts", "jp", "jt", "jt1", "indself", "indfor", "indminor", "laytrop", "cldymc")

# ---- what the two pin tests share -----------------------------------------------------------------------------------
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rrtmg_refexec_golden.npz")


def rel_err(a, b):
    """Largest elementwise relative difference, elements below 1e-12 of the array's largest magnitude measured
    against that floor (the idiom of the other parity tests)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))


def run_case(impl, name, set_mcica, reset_mcica, skip=()):
    """Run one case through `impl` (oracle.binding, or the CUDA host mirror: same call interface) and return
    {"lw/<key>": array, "sw/<key>": array} under the golden file's key names.  set_mcica(ih, corr) / reset_mcica()
    select the McICA options on that implementation; `skip` names intermediates the implementation does not hold."""
    from geosradiation_gridcomp_b200.synthetic import make_columns
    c = CASES[name]
    s = make_columns(c["ncol"], c["nlay"], seed=c["seed"])
    lw_fn = getattr(impl, "run_lw", None) or impl.rrtmg_lw
    sw_fn = getattr(impl, "run_sw", None) or impl.rrtmg_sw
    out = {}
    special = c["ih"] != 1 or c["corr"] is not None
    if special:
        set_mcica(c["ih"], c["corr"])
    try:
        if c["lw"] is not None:
            lw_taps = tuple(k for k in LW_TAPS if k not in skip) if c["taps"] else ()
            r = lw_fn(s, taps=lw_taps, **c["lw"])
            assert r.get("rc", 0) == 0
            for k in LW_OUT + lw_taps:
                out["lw/" + k] = r[k]
        if c["sw"] is not None:
            r = sw_fn(s, taps=tuple(SW_TAPS.values()) if c["taps"] else (), **c["sw"])
            assert r.get("rc", 0) == 0
            for k in SW_OUT:
                out["sw/" + k] = r[k]
            if c["sw"].get("radval"):
                out["sw/radval"] = r["radval"]
            if c["taps"]:
                for k, tap in SW_TAPS.items():
                    out["sw/" + k] = r[tap]
    finally:
        if special:
            reset_mcica()
    return out


def check_case(got, golden, name, tol, tol_taps=None):
    """Integers bit for bit, reals within tol (intermediates within tol_taps).  Returns (arrays compared, arrays
    that are bit-identical, worst relative difference)."""
    n = same = 0
    worst = 0.0
    for k, v in got.items():
        ref = golden[f"{name}/{k}"]
        key = k.split("/")[1]
        n += 1
        if key in INTEGER_KEYS:
            np.testing.assert_array_equal(np.asarray(v).astype(np.int64), ref.astype(np.int64), err_msg=f"{name}/{k}")
            same += 1
            continue
        v = np.asarray(v, dtype=np.float64)
        assert v.shape == ref.shape, (name, k, v.shape, ref.shape)
        same += bool(np.array_equal(v, ref))
        e = rel_err(v, ref)
        worst = max(worst, e)
        limit = tol_taps if (tol_taps is not None and key in LW_TAPS + tuple(SW_TAPS)) else tol
        assert e <= limit, f"{name}/{k}: relative difference {e:.3e} > {limit:.1e}"
    return n, same, worst


def heating_inputs(golden, n):
    """The arguments of rrtmgx_heating_rate (net upward flux at levels, surface first; interface pressures in hPa, surface
    first) for the LW and the SW fluxes of the golden refresh, from the native top-down arrays the parent component
    holds (GEOS_RadiationGridComp.F90:801-814: FLW, FSW net downward, PLE in Pa, level 0 at the model top)."""
    import make_golden_from_refexec as gen
    flw = golden["refresh/irr/flxu"] + golden["refresh/irr/flxd"]
    fsw = gen.rad_fsw(n, golden["refresh/sol/fsw"])
    plev = np.asfortranarray(np.asarray(n["ple"])[:, ::-1] / 100.0)
    return {"radlw": np.asfortranarray(-flw[:, ::-1]), "radsw": np.asfortranarray(-fsw[:, ::-1])}, plev


def heating_formula(fnet, plev, grav, cp):
    """What the parity tests use as the plain restatement of the epilogue (K/day, layer 1 at the surface)."""
    return (fnet[:, :-1] - fnet[:, 1:]) * (grav / cp) / ((plev[:, :-1] - plev[:, 1:]) * 100.0) * 86400.0
