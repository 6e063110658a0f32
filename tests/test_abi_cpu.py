"""The C-ABI library loads and exports every symbol include/rrtmgx.h declares; without a GPU the
product path refuses to run (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rrtmgx.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rrtmgx_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from geosradiation_gridcomp_b200 import host
    lib = host.lib()
    names = declared_functions()
    assert {"rrtmgx_init", "rrtmgx_lw_run", "rrtmgx_sw_run", "rrtmgx_finalize", "rrtmgx_set_mcica",
            "rrtmgx_heating_rate", "rrtmgx_strerror", "rrtmgx_launch_count"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_mirrors_match_header_field_order():
    """host.LwArgs / host.SwArgs must list the fields in the header's order (ctypes mirrors)."""
    from geosradiation_gridcomp_b200 import host
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for cname, mirror in (("RrtmgxLwArgs", host.LwArgs), ("RrtmgxSwArgs", host.SwArgs), ("RrtmgxTaps", host.Taps)):
        body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} %s;" % cname, src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(double|int32_t|int|void|uint8_t)\s*", "", decl)
            fields += [f.strip().lstrip("*").strip() for f in decl.split(",")]
        assert fields == [f[0] for f in mirror._fields_], cname


def test_no_cpu_fallback():
    import torch
    from geosradiation_gridcomp_b200 import host
    lib = host.lib()
    assert lib.rrtmgx_strerror(-1).decode().startswith("no usable CUDA device")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal path cannot be exercised")
    args = host.LwArgs()
    assert lib.rrtmgx_lw_run(C.byref(args)) == -2          # RRTMGX_ENOTINIT
    sargs = host.SwArgs()
    assert lib.rrtmgx_sw_run(C.byref(sargs)) == -2
    cfg = host.Config()
    cfg.device = -1
    cfg.inhomogeneity = 1
    assert lib.rrtmgx_init(C.byref(cfg)) == -1              # RRTMGX_ENODEVICE
    with pytest.raises(host.RrtmgxError):
        host.init()


def test_host_mirror_rejects_wrong_layout():
    from geosradiation_gridcomp_b200 import host
    a = np.zeros((4, 3))                                    # C order, not the reference layout
    with pytest.raises(ValueError):
        host._addr(a, False)
    with pytest.raises(ValueError):
        host._addr(np.zeros((4, 3), dtype=np.float32, order="F"), False)
    with pytest.raises(ValueError):
        host._addr(np.zeros(4), True)                       # device=True needs device memory
    assert host._addr(np.zeros((4, 3), order="F"), False) != 0


def test_product_does_not_import_the_oracle():
    """Nothing under the package may import, include, link or call the CPU oracle."""
    pkg = os.path.join(ROOT, "geosradiation_gridcomp_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.\.?oracle|#include\s+\"[^\"]*oracle|liboracle|\boracle_[a-z]+\s*\(")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dirpath, f)


def test_fortran_shim_types_match_header_field_order():
    """The ISO_C_BINDING derived types of the Fortran shim list the fields in the header's order."""
    from geosradiation_gridcomp_b200 import host
    f90 = open(os.path.join(ROOT, "geosradiation_gridcomp_b200", "fortran", "rrtmgx_c.F90")).read()
    for tname, mirror in (("rrtmgx_lw_args", host.LwArgs), ("rrtmgx_sw_args", host.SwArgs),
                          ("rrtmgx_irrad_args", host.IrradArgs), ("rrtmgx_solar_args", host.SolarArgs),
                          ("rrtmgx_lw_variants", host.LwVariants), ("rrtmgx_sw_no_aerosol", host.SwNoAerosol)):
        body = re.search(r"type, bind\(C\) :: %s\n(.*?)end type" % tname, f90, flags=re.S).group(1)
        fields = []
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" in line:
                fields += [f.strip() for f in line.split("::")[1].split(",")]
        assert [f.lower() for f in fields] == [f[0].lower() for f in mirror._fields_], tname


def test_ctypes_mirrors_match_the_compiled_header(tmp_path):
    """sizeof and the offset of the last field of every argument struct, as gcc lays out include/rrtmgx.h,
    against the ctypes mirrors of host.py (a drifted mirror would shift every pointer after the drift)."""
    import subprocess
    from geosradiation_gridcomp_b200 import host
    pairs = [("RrtmgxLwArgs", host.LwArgs, "dolrb_dTs"), ("RrtmgxSwArgs", host.SwArgs, "radval"),
             ("RrtmgxIrradArgs", host.IrradArgs, "dolrb_dts"), ("RrtmgxSolarArgs", host.SolarArgs, "cotlp"),
             ("RrtmgxIrradUpdateArgs", host.IrradUpdateArgs, "flnsc"), ("RrtmgxLwVariants", host.LwVariants, "duflx_dTs"),
             ("RrtmgxSwNoAerosol", host.SwNoAerosol, "fswband"), ("RrtmgxTaps", host.Taps, "ssi"),
             ("RrtmgxConfig", host.Config, "corr")]
    src = tmp_path / "probe.c"
    body = "".join(f'  printf("{n} %zu %zu\\n", sizeof({n}), offsetof({n}, {last}));\n' for n, _, last in pairs)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rrtmgx.h"\nint main(void) {\n' + body + "  return 0;\n}\n")
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for name, mirror, last in pairs:
        assert got[name] == (C.sizeof(mirror), getattr(mirror, last).offset), name


# ---- the Fortran shim keeps the reference's own dummy-argument lists (the drop-in claim, checked on the text) ---------
_REF_TREE = "/root/reference"


def _dummy_arguments(path, name, strip_radval=True):
    """Ordered dummy-argument names of `subroutine name(...)` in a free-form Fortran file (continuation lines joined,
    comments and the reference's compile-time-off SOLAR_RADVAL blocks dropped)."""
    import re
    out, skip = [], False
    for line in open(path, errors="ignore"):
        s = line.strip()
        if s.startswith("#ifdef SOLAR_RADVAL"):
            skip = True
        elif s.startswith("#endif") or s.startswith("#else"):
            skip = False
        elif not (skip and strip_radval) and not s.startswith("#"):
            out.append(line.split("!")[0].rstrip())
    text = "\n".join(out)
    m = re.search(r"subroutine\s+" + name + r"\s*\((.*?)\)", text, re.S | re.I)
    assert m, (path, name)
    return [a.strip().lower() for a in m.group(1).replace("&", " ").replace("\n", " ").split(",") if a.strip()]


@pytest.mark.skipif(not os.path.isdir(_REF_TREE), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("ref_file,shim_file,name", [
    ("GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/src/rrtmg_lw_rad.F90", "rrtmg_lw_rad.F90", "rrtmg_lw"),
    ("GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90", "rrtmg_sw_rad.F90", "rrtmg_sw"),
    ("GEOS_RadiationShared/cloud_subcol_gen.F90", "rrtmgx_init_mods.F90", "initialize_cloud_subcol_gen"),
    ("GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90", "rrtmgx_init_mods.F90", "set_inhomogeneity"),
])
def test_fortran_shim_has_the_reference_argument_lists(ref_file, shim_file, name):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = _dummy_arguments(os.path.join(_REF_TREE, ref_file), name)
    mine = _dummy_arguments(os.path.join(root, "geosradiation_gridcomp_b200", "fortran", shim_file), name)
    assert mine == ref, name


@pytest.mark.skipif(not os.path.isdir(_REF_TREE), reason="the reference tree is only present in the build container")
def test_fortran_shim_has_the_reference_argument_list_of_the_radval_build():
    """Compiled with -DSOLAR_RADVAL (GEOSsolar_GridComp/CMakeLists.txt:18-20) rrtmg_sw gains 120 dummies
    (SW/src/rrtmg_sw_rad.F90:85-122): the shim lists them in the reference's order, which is also the column order of
    RrtmgxSwArgs::radval (host.RADVAL_NAMES) - the shim copies column q to the q-th of them."""
    import re
    from geosradiation_gridcomp_b200 import host
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim_path = os.path.join(root, "geosradiation_gridcomp_b200", "fortran", "rrtmg_sw_rad.F90")
    ref = _dummy_arguments(os.path.join(_REF_TREE, "GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90"),
                           "rrtmg_sw", strip_radval=False)
    mine = _dummy_arguments(shim_path, "rrtmg_sw", strip_radval=False)
    assert mine == ref and len(ref) == len(_dummy_arguments(shim_path, "rrtmg_sw")) + host.NRADVAL
    i = ref.index("cotnlp") + 1
    assert tuple(ref[i:i + host.NRADVAL]) == host.RADVAL_NAMES
    copies = re.findall(r"^\s*(\w+) = radval\(:,(\d+)\)", open(shim_path).read(), flags=re.M)
    assert [(n, int(q)) for n, q in copies] == [(n, q + 1) for q, n in enumerate(host.RADVAL_NAMES)]
