#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` report:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_source_summary.py x.csv [kernel-substring]
For every kernel: total stall samples by reason, and the source lines (file:line) that hold the most samples, with the
instructions executed on them and their dominant stall reasons."""
import collections
import csv
import sys


def main(path, want=None, top=40):
    rows = csv.reader(open(path, newline=""))
    file_path = func = None
    hdr = None
    per_func = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            file_path = r[1]; continue
        if r[0] == "Function Name":
            func = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or func is None or r[0] in ("", None):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, "0").replace(",", ""))
            except ValueError:
                return 0.0
        F = per_func.setdefault(func, {"lines": [], "stalls": collections.Counter(), "samples": 0., "inst": 0.})
        samples = num("# Samples")
        inst = num("Instructions Executed")
        st = {k[6:]: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        F["lines"].append((samples, inst, file_path.split("/")[-1], line, r[1].strip()[:110], st))
        F["samples"] += samples; F["inst"] += inst
        for k, v in st.items():
            F["stalls"][k] += v
    for func, F in per_func.items():
        if want and want not in func:
            continue
        print("=" * 20, func[:100])
        tot = max(F["samples"], 1.)
        print(f"samples {F['samples']:.0f}, warp instructions {F['inst']:.3e}")
        print("stall reasons: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in F["stalls"].most_common(9)))
        for samples, inst, fn, line, src, st in sorted(F["lines"], key=lambda x: -x[0])[:top]:
            main3 = ", ".join(f"{k} {100 * v / max(samples, 1):.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v)
            print(f"{100 * samples / tot:5.1f}% {100 * inst / max(F['inst'], 1):5.1f}%i  {fn}:{line:<5d} {src[:90]}\n           [{main3}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
