#!/usr/bin/env python
"""profiles/kernel_counters.json from an ncu launch list of one bench step:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none --csv --log-file launches.csv python bench.py --steps 1 --warmup 1 --ncol N --no-e2e --no-cpu
    python tools/ncu_counters.py launches.csv N [out.json]

Per kernel (named like the library's own profile tags, so bench.py can join them with its in-run CUDA-event times):
DRAM bytes and FP64 thread-instructions PER COLUMN AND LAUNCH, averaged over the launches of the capture.  bench.py
multiplies them by the columns of a launch and divides by the launch time it measures itself: `roofline.traffic`
and `roofline.fp64` are then per-launch figures of THIS run's timing on the counters of the committed capture."""
import collections
import csv
import json
import re
import subprocess
import sys


def tag_of(name):
    base = re.sub(r"^(void )?(rrtmgx::)?(<unnamed>::)?", "", name)
    base = re.sub(r"\((int|bool)\)", "", base)   # ncu prints template arguments with or without their casts
    m = re.match(r"sw_band_kernel<(\d+), (\d+), (\d+), (\d+), (\d|true|false)>", base)
    if m:
        return f"{'sw_up_kernel' if m.group(5) in ('1', 'true') else 'sw_band_kernel'}<{m.group(1)},gn{m.group(2)},r{m.group(3)},c{m.group(4)}>"
    m = re.match(r"lw_band_kernel<(\d+), (\d+), (\d+), (\d+)>", base)
    if m:
        return f"lw_band_kernel<{m.group(1)},gn{m.group(2)},r{m.group(3)},c{m.group(4)}>"
    m = re.match(r"sw_down_kernel<(\d+), (\d+)>", base)
    if m:
        return f"sw_down_kernel<{m.group(1)},gn{m.group(2)}>"
    m = re.match(r"mcica_kernel<(rrtmgx::)?(\w+)>", base)
    if m:
        return f"mcica_kernel<{m.group(2)}>"
    return re.sub(r"[<(].*", "", base)


def main(path, ncol_per_launch, out):
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
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        if "rrtmgx" not in k:
            continue
        a = agg.setdefault(tag_of(k), collections.defaultdict(float))
        a["n"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.)
        a["dram"] += m.get("dram__bytes_read.sum", 0.) + m.get("dram__bytes_write.sum", 0.)
        a["fp64"] += m.get("sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum", 0.)
        a["pipe"] += m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.)
    try:
        head = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
    except Exception:
        head = None
    tot_ns = sum(a["ns"] for a in agg.values()) or 1.
    res = {"source": path, "commit_of_capture": head, "columns_per_launch": ncol_per_launch,
           "what": "ncu counters per COLUMN and LAUNCH (averages over the capture's launches): dram = dram__bytes_read.sum + "
                   "dram__bytes_write.sum, fp64 = sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum; "
                   "share = kernel's part of the serialised, cold-cache launch list",
           "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = a["n"]
        res["kernels"][k] = {"launches": int(n), "ms_per_launch_under_ncu": a["ns"] / n / 1e6, "share": a["ns"] / tot_ns,
                             "dram_bytes_per_column": a["dram"] / n / ncol_per_launch,
                             "fp64_thread_instructions_per_column": a["fp64"] / n / ncol_per_launch,
                             "fp64_pipe_active_pct": a["pipe"] / n}
    # a step = one chunk through every kernel: each kernel once, the kernels shared by the LW and the SW path twice
    twice = ("check_negative_kernel", "mcica_threshold_kernel", "mcica_prep_kernel", "cloudy_flag_kernel")
    per_chunk = lambda k: 2 if k in twice else 1
    res["step"] = {"dram_bytes_per_column": sum(v["dram_bytes_per_column"] * per_chunk(k) for k, v in res["kernels"].items()),
                   "fp64_thread_instructions_per_column":
                       sum(v["fp64_thread_instructions_per_column"] * per_chunk(k) for k, v in res["kernels"].items()),
                   "kernels_per_chunk": sum(per_chunk(k) for k in res["kernels"])}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res["step"]))
    for k, v in list(res["kernels"].items())[:12]:
        print(f"{v['share'] * 100:5.1f}%  {v['ms_per_launch_under_ncu']:7.3f} ms  dram {v['dram_bytes_per_column'] / 1e3:8.1f} KB/col  "
              f"fp64 {v['fp64_thread_instructions_per_column'] / 1e3:8.1f} k/col  pipe {v['fp64_pipe_active_pct']:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "profiles/kernel_counters.json")
