"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        k = r[ki][:78]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'total ms':>10} {'n':>4} {'share':>6}  kernel   (gpu__time_duration.sum, serialised by ncu)")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{v / 1e6:10.3f} {n:4d} {100 * v / tot:5.1f}%  {k}")
    print(f"{tot / 1e6:10.3f} ms total over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
