"""The recipe that pins the oracle by reference OUTPUT (oracle/build_ref.sh, oracle/ref_recipe/,
tests/golden/make_golden_from_ref.py).  No Fortran compiler exists in this image or on the GPU boxes seen so far, so
what can be held here is the recipe itself: it names every source it needs, they exist in the reference tree, it
reports the missing compiler with its own exit status, and - the moment a compiler is there - it builds the reference,
generates the golden vectors and the C restatement is held to them."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "oracle", "build_ref.sh")
GOLD = os.path.join(ROOT, "tests", "golden", "rrtmg_ref_golden_L72.npz")
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
COMPILERS = ("gfortran", "ifx", "ifort", "flang", "flang-new", "nvfortran")
have_ref = os.path.isdir(os.path.join(REF, "GEOSirrad_GridComp"))
have_fc = any(shutil.which(c) for c in COMPILERS)


@pytest.mark.skipif(not have_ref, reason="no reference tree on this machine")
def test_recipe_lists_existing_sources_in_a_dry_run():
    r = subprocess.run(["bash", SCRIPT, "--dry-run"], env=dict(os.environ, FC="/bin/true"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if " -c " in l]
    srcs = [l.split(" -c ")[1].split(" -o ")[0] for l in lines]
    assert len(srcs) == 93
    assert all(os.path.isfile(s) for s in srcs)
    # the reference sources are compiled where they lie; only the stand-ins and the C wrappers come from this repository
    own = [s for s in srcs if not s.startswith(REF)]
    assert sorted(os.path.basename(s) for s in own) == ["mapl_stub.F90", "ref_capi.F90"]
    # every RRTMG source file of the path is in the list
    names = {os.path.basename(s) for s in srcs}
    for must in ("rrtmg_lw_rad.F90", "rrtmg_lw_rtrnmc.F90", "rrtmg_lw_taumol.F90", "rrtmg_lw_setcoef.F90", "rrtmg_lw_cldprmc.F90",
                 "rrtmg_sw_rad.F90", "rrtmg_sw_spcvmc.F90", "rrtmg_sw_taumol.F90", "rrtmg_sw_setcoef.F90", "rrtmg_sw_cldprmc.F90",
                 "NRLSSI2.F90", "cloud_subcol_gen.F90", "cloud_condensate_inhomogeneity.F90"):
        assert must in names, must


@pytest.mark.skipif(have_fc, reason="a Fortran compiler exists: the real build is exercised instead")
def test_missing_compiler_is_reported_not_hidden():
    r = subprocess.run(["bash", SCRIPT], capture_output=True, text=True, env={k: v for k, v in os.environ.items() if k != "FC"})
    assert r.returncode == 3
    assert "no Fortran compiler" in r.stderr
    assert not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgeosref.so"))


def test_stand_ins_cover_what_the_sw_sources_use():
    """Every MAPL / ESMF entity the two SW sources touch is provided by the stand-ins."""
    if not have_ref:
        pytest.skip("no reference tree on this machine")
    import re
    sw = os.path.join(REF, "GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src")
    text = open(os.path.join(sw, "rrtmg_sw_rad.F90")).read() + open(os.path.join(sw, "rrtmg_sw_spcvmc.F90")).read()
    code = "\n".join(l.split("!")[0] for l in text.splitlines())
    used = set(re.findall(r"\b(MAPL_\w+|_ASSERT|_FAIL|_RETURN|_VERIFY|__RC__|_RC|_SUCCESS)\b", code))
    stub = open(os.path.join(ROOT, "oracle/ref_recipe/mapl_stub.F90")).read() + \
        open(os.path.join(ROOT, "oracle/ref_recipe/MAPL_Generic.h")).read()
    used.discard("MAPL_Generic")   # the include file itself: provided as oracle/ref_recipe/MAPL_Generic.h
    for u in used:
        assert re.search(r"\b%s\b" % re.escape(u), stub, re.I), u


@pytest.mark.skipif(not (have_fc and have_ref) and not os.path.exists(GOLD),
                    reason="no reference output yet: no Fortran compiler here (oracle/build_ref.sh exits 3)")
def test_oracle_equals_reference_output():
    """THE pin: the C restatement against numbers the Fortran reference produced."""
    if have_fc and have_ref and not os.path.exists(GOLD):
        subprocess.check_call(["python", os.path.join(ROOT, "tests/golden/make_golden_from_ref.py")])
    g = np.load(GOLD)
    if int(g["real_bytes"]) != 8:
        pytest.skip("golden vectors were made with real(4)")
    from geosradiation_gridcomp_b200.synthetic import make_columns
    from oracle import binding as oracle
    s = make_columns(int(g["ncol"]), int(g["nlay"]), seed=int(g["seed"]))
    lw, sw = oracle.rrtmg_lw(s), oracle.rrtmg_sw(s, do_drfband=True)
    np.testing.assert_array_equal(lw["clearCounts"], g["lw_clearCounts"])
    np.testing.assert_array_equal(sw["clearCounts"], g["sw_clearCounts"])
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))
    for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs"):
        assert rel(lw[k], g["lw_" + k]) <= 1e-9, k
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf"):
        assert rel(sw[k], g["sw_" + k]) <= 1e-9, k
