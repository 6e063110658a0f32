#!/usr/bin/env python3
"""Extract the literal data tables of the reference RRTMG LW/SW + McICA sources into one
binary blob (`geosradiation_gridcomp_b200/data/rrtmg_tables.bin`).

The reference keeps its k-distribution, Planck, cloud-optics, reference-atmosphere and
condensate-inhomogeneity tables as Fortran array constructors inside source files
(e.g. LW/src/rrtmg_lw_k_g_03.F90, LW/src/rrtmg_lw_init.F90:1981-3269,
LW/src/rrtmg_lw_setcoef.F90:646-2218, SH/cloud_condensate_inhomogeneity.F90:127-70421).
This script parses those constructors (handling `&` continuations on either side, comment
lines inside constructors, `a:b` slices and lower bounds such as `13:59` / `16:29`) and the
module declarations that give each array its shape.  Decimal literals are parsed as fp64,
which is what a promoted-real (`-fdefault-real-8`) build of the reference would hold.

Only DATA is extracted; every computation on it (g-point reduction, lookup-table
construction, the band-29 irradnce scaling of SW/src/rrtmg_sw_k_g_29.F90:80-81) is restated
in code.  The blob is written once here (the reference tree is not available on the GPU box)
and committed; re-run with `python tools/extract_tables.py` to regenerate.

Blob layout (little endian):
  magic "RRTMGTB1", int32 n_entries, then per entry:
     char name[48], int32 dtype (0=f64,1=i32), int32 ndim, int32 dims[6], int64 offset, int64 nbytes
  followed by the raw array data (Fortran/column-major order, 8-byte aligned).
"""
import os
import re
import struct
import sys

import numpy as np

REF = os.environ.get("RRTMG_REFERENCE", "/root/reference")
LW = f"{REF}/GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model"
SW = f"{REF}/GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model"
SH = f"{REF}/GEOS_RadiationShared"
OUT = os.environ.get("RRTMG_TABLES_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                         "geosradiation_gridcomp_b200", "data", "rrtmg_tables.bin")


def logical_statements(path):
    """Yield Fortran statements with comments stripped and continuations joined."""
    stmts = []
    cur = None
    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.rstrip("\n")
            if line.lstrip().startswith("#"):
                continue
            # strip comments (no string literals containing '!' in the data we parse)
            if "!" in line:
                line = line[: line.index("!")]
            s = line.strip()
            if not s:
                continue
            if cur is not None:
                if s.startswith("&"):
                    s = s[1:].strip()
                if s.endswith("&"):
                    cur += " " + s[:-1].strip()
                else:
                    cur += " " + s
                    stmts.append(cur)
                    cur = None
            else:
                if s.endswith("&"):
                    cur = s[:-1].strip()
                else:
                    stmts.append(s)
    if cur is not None:
        stmts.append(cur)
    return stmts


def parse_number(tok):
    t = tok.strip().lower().replace("d", "e")
    t = re.sub(r"_\w+$", "", t)
    return float(t)


def split_top(s):
    """split on commas not inside parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


class Decls:
    """Array shapes (with lower bounds) and integer parameters from module files."""

    def __init__(self):
        self.params = {}
        self.arrays = {}  # name -> (dtype, [(lo,hi),...])

    def ev(self, expr):
        e = expr.strip().lower()
        return int(eval(e, {"__builtins__": {}}, self.params))

    def dims(self, spec):
        out = []
        for d in split_top(spec):
            d = d.strip()
            if ":" in d:
                lo, hi = d.split(":")
                out.append((self.ev(lo), self.ev(hi)))
            else:
                out.append((1, self.ev(d)))
        return out

    def load(self, path):
        for st in logical_statements(path):
            low = st.lower()
            m = re.match(r"^integer\s*,\s*parameter\s*::\s*(\w+)\s*=\s*([-\w+*/ ()]+)$", low)
            if m:
                try:
                    self.params[m.group(1)] = self.ev(m.group(2))
                except Exception:
                    pass
                continue
            m = re.match(r"^(real|integer)\s*(\*\d+)?\s*(,\s*dimension\s*\(([^)]*)\))?\s*(,\s*parameter)?\s*::\s*(.*)$", low)
            if not m:
                continue
            dtype = "f8" if m.group(1) == "real" else "i4"
            common = m.group(4)
            rest = m.group(6)
            if "=" in rest and m.group(5):
                # parameter arrays with initialisers are handled by the assignment parser
                rest = rest.split("=")[0]
            for item in split_top(rest):
                item = item.strip()
                mm = re.match(r"^(\w+)\s*(\((.*)\))?$", item)
                if not mm:
                    continue
                name = mm.group(1)
                spec = mm.group(3) if mm.group(3) else common
                try:
                    if spec:
                        self.arrays[name] = (dtype, self.dims(spec))
                    else:
                        self.arrays[name] = (dtype, [])
                except Exception:
                    pass


def assign_from_file(path, decls, store, prefix, wanted=None, rename=None):
    """Parse `name(subs) = (/ ... /)` and scalar assignments into numpy arrays."""
    for st in logical_statements(path):
        m = re.match(r"^(\w+)\s*(\(([^=]*)\))?\s*=\s*(\(/|\[)(.*)(/\)|\])\s*$", st, re.S)
        scalar = None
        if not m:
            ms = re.match(r"^(\w+)\s*=\s*([-+]?[0-9.]+([eEdD][-+]?\d+)?)\s*$", st)
            if not ms:
                continue
            name = ms.group(1).lower()
            scalar = parse_number(ms.group(2))
            subs = None
        else:
            name = m.group(1).lower()
            subs = m.group(3)
        if wanted is not None and name not in wanted:
            continue
        if name not in decls.arrays:
            continue
        dtype, dims = decls.arrays[name]
        key = prefix + (rename.get(name, name) if rename else name)
        if scalar is not None:
            if dims:
                continue
            store[key] = np.array([scalar], dtype="f8")
            continue
        vals = [parse_number(t) for t in split_top(m.group(5)) if t.strip()]
        shape = [hi - lo + 1 for lo, hi in dims]
        if key not in store:
            arr = np.full(shape, np.nan, dtype="f8", order="F")
            store[key] = arr
        arr = store[key]
        if subs is None:
            sl = [slice(None)] * len(dims)
        else:
            parts = [p.strip() for p in split_top(subs)]
            assert len(parts) == len(dims), (path, st[:80])
            sl = []
            for p, (lo, hi) in zip(parts, dims):
                if p == ":":
                    sl.append(slice(None))
                elif ":" in p:
                    a, b = p.split(":")
                    sl.append(slice(decls.ev(a) - lo, decls.ev(b) - lo + 1))
                else:
                    sl.append(decls.ev(p) - lo)
        sub = arr[tuple(sl)]
        v = np.array(vals, dtype="f8")
        assert v.size == sub.size, (path, name, subs, v.size, sub.size)
        arr[tuple(sl)] = v.reshape(sub.shape, order="F")


def finish(store, int_names=()):
    for k, v in list(store.items()):
        if np.isnan(v).any():
            raise SystemExit(f"table {k} not completely filled ({np.isnan(v).sum()} holes)")
        base = k.split(".")[-1]
        if base in int_names:
            store[k] = np.asfortranarray(v.astype("i4"))


def main():
    store = {}

    # ---------------- longwave ----------------
    lw_par = Decls()
    lw_par.load(f"{LW}/modules/parrrtm.F90")
    for b in range(1, 17):
        d = Decls()
        d.params = dict(lw_par.params)
        d.load(f"{LW}/modules/rrlw_kg{b:02d}.F90")
        wanted = {n for n in d.arrays if n.endswith("o") or n.endswith("o_mn2") or
                  re.match(r"k[ab]o_m\w+", n)}
        assign_from_file(f"{LW}/src/rrtmg_lw_k_g_{b:02d}.F90", d, store, f"lw.kg{b:02d}.", wanted)
    d = Decls()
    d.params = dict(lw_par.params)
    d.params["ntbl"] = 10000
    for mod in ("rrlw_ref", "rrlw_wvn", "rrlw_cld"):
        d.load(f"{LW}/modules/{mod}.F90")
    assign_from_file(f"{LW}/src/rrtmg_lw_setcoef.F90", d, store, "lw.ref.",
                     {"pref", "preflog", "tref", "chi_mls"})
    assign_from_file(f"{LW}/src/rrtmg_lw_setcoef.F90", d, store, "lw.wvn.",
                     {"totplnk", "totplk16", "totplnkderiv", "totplk16deriv"})
    assign_from_file(f"{LW}/src/rrtmg_lw_init.F90", d, store, "lw.cld.",
                     {"absice0", "absice1", "absice2", "absice3", "absice4", "absliq1"})
    assign_from_file(f"{LW}/src/rrtmg_lw_init.F90", d, store, "lw.wvn.",
                     {"ngc", "ngs", "ngm", "ngn", "ngb", "wt", "nspa", "nspb"})

    # ---------------- shortwave ----------------
    sw_par = Decls()
    sw_par.load(f"{SW}/modules/parrrsw.F90")
    for b in range(16, 30):
        d = Decls()
        d.params = dict(sw_par.params)
        d.load(f"{SW}/modules/rrsw_kg{b}.F90")
        wanted = {n for n in d.arrays if n.endswith("o")}
        if d.arrays.get("rayl", (None, [1]))[1] == []:
            wanted.add("rayl")
        assign_from_file(f"{SW}/src/rrtmg_sw_k_g_{b}.F90", d, store, f"sw.kg{b}.", wanted)
    d = Decls()
    d.params = dict(sw_par.params)
    for mod in ("rrsw_ref", "rrsw_wvn", "rrsw_cld"):
        d.load(f"{SW}/modules/{mod}.F90")
    assign_from_file(f"{SW}/src/rrtmg_sw_setcoef.F90", d, store, "sw.ref.",
                     {"pref", "preflog", "tref"})
    assign_from_file(f"{SW}/src/rrtmg_sw_init.F90", d, store, "sw.cld.",
                     {"extliq1", "ssaliq1", "asyliq1", "extice2", "ssaice2", "asyice2",
                      "extice3", "ssaice3", "asyice3", "fdlice3", "extice4", "ssaice4",
                      "asyice4", "abari", "bbari", "cbari", "dbari", "ebari", "fbari"})
    assign_from_file(f"{SW}/src/rrtmg_sw_init.F90", d, store, "sw.wvn.",
                     {"ngc", "ngs", "ngm", "ngn", "ngb", "wt", "nspa", "nspb", "icxa"})

    # ---------------- NRLSSI2 average-cycle index tables ----------------
    d = Decls()
    d.params["nsolfrac"] = 134
    d.arrays["mgavgcyc"] = ("f8", [(1, 134)])
    d.arrays["sbavgcyc"] = ("f8", [(1, 134)])
    for st in logical_statements(f"{SW}/src/NRLSSI2.F90"):
        m = re.match(r"^real\s*,\s*parameter\s*::\s*(mgavgcyc|sbavgcyc)\s*\(nsolfrac\)\s*=\s*\(/(.*)/\)$",
                     st, re.S | re.I)
        if m:
            vals = [parse_number(t) for t in split_top(m.group(2)) if t.strip()]
            assert len(vals) == 134
            store["sw.nrlssi2." + m.group(1).lower()] = np.array(vals, dtype="f8")

    # ---------------- McICA condensate-inhomogeneity tables ----------------
    # two subroutines fill the same `xcw` name: split the statement stream on the subroutine names
    path = f"{SH}/cloud_condensate_inhomogeneity.F90"
    d = Decls()
    d.arrays["xcw"] = ("f8", [(1, 1000), (1, 140)])
    stmts = logical_statements(path)
    which = None
    tmp = {"beta": {}, "gamma": {}}
    for st in stmts:
        low = st.lower()
        if low.startswith("subroutine tabulate_xcw_beta"):
            which = "beta"
        elif low.startswith("subroutine tabulate_xcw_gamma"):
            which = "gamma"
        elif low.startswith("end subroutine"):
            which = None
        elif which and low.startswith("xcw"):
            m = re.match(r"^xcw\s*\(\s*:\s*,\s*(\d+)\s*\)\s*=\s*\(/(.*)/\)$", st, re.S | re.I)
            assert m, st[:60]
            col = int(m.group(1))
            vals = [parse_number(t) for t in split_top(m.group(2)) if t.strip()]
            assert len(vals) == 1000
            tmp[which][col] = vals
    for w in ("beta", "gamma"):
        arr = np.empty((1000, 140), dtype="f8", order="F")
        for c in range(1, 141):
            arr[:, c - 1] = tmp[w][c]
        store[f"mcica.xcw_{w}"] = arr

    finish(store, int_names={"ngc", "ngs", "ngm", "ngn", "ngb", "nspa", "nspb", "icxa"})

    # ---------------- write blob ----------------
    names = sorted(store)
    hdr = 8 + 4 + len(names) * (48 + 4 + 4 + 24 + 8 + 8)
    off = (hdr + 7) // 8 * 8
    entries, blobs = [], []
    for n in names:
        a = store[n]
        dt = 1 if a.dtype == np.int32 else 0
        raw = np.asfortranarray(a).tobytes(order="F")
        dims = list(a.shape) + [0] * (6 - a.ndim)
        entries.append((n, dt, a.ndim, dims, off, len(raw)))
        blobs.append((off, raw))
        off = (off + len(raw) + 7) // 8 * 8
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(b"RRTMGTB1")
        f.write(struct.pack("<i", len(names)))
        for n, dt, nd, dims, o, nb in entries:
            f.write(n.encode().ljust(48, b"\0"))
            f.write(struct.pack("<ii6iqq", dt, nd, *dims, o, nb))
        for o, raw in blobs:
            f.seek(o)
            f.write(raw)
    tot = sum(len(r) for _, r in blobs)
    print(f"wrote {len(names)} tables, {tot/1e6:.2f} MB -> {os.path.normpath(OUT)}")
    if "-v" in sys.argv:
        for n, dt, nd, dims, o, nb in entries:
            print(f"  {n:32s} {'i4' if dt else 'f8'} {dims[:nd]}")


if __name__ == "__main__":
    main()
