"""Summarise an ncu --csv launch list that carries several metrics per launch (one row per metric)."""
import collections
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, mi, vi, idi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        per.setdefault((r[idi], r[ki]), {})[r[mi]] = v
    return per


def main(path, ncol=None, steps=None):
    per = load(path)
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k[:64], collections.defaultdict(float))
        a["n"] += 1
        for kk, v in m.items():
            a[kk] += v
    T = "gpu__time_duration.sum"
    tot = sum(a[T] for a in agg.values())
    rd = sum(a["dram__bytes_read.sum"] for a in agg.values())
    wr = sum(a["dram__bytes_write.sum"] for a in agg.values())
    print(f"{'ms':>8} {'n':>4} {'share':>6} {'regs':>5} {'warps%':>6} {'fp64%':>6} {'rdGB':>7} {'wrGB':>7} {'GB/s':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][T]):
        n, t = a["n"], a[T]
        print(f"{t / 1e6:8.2f} {int(n):4d} {100 * t / tot:5.1f}% {a['launch__registers_per_thread'] / n:5.0f} "
              f"{a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n:6.1f} "
              f"{a['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'] / n:6.1f} "
              f"{a['dram__bytes_read.sum'] / 1e9:7.2f} {a['dram__bytes_write.sum'] / 1e9:7.2f} "
              f"{(a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']) / max(t, 1):7.1f}  {k}")
    print(f"{tot / 1e6:8.2f} ms in {int(sum(a['n'] for a in agg.values()))} launches; DRAM read {rd / 1e9:.2f} GB, write {wr / 1e9:.2f} GB")
    return agg, tot, rd, wr


if __name__ == "__main__":
    main(sys.argv[1])
