"""Times the device-resident LW+SW step (the bench's step: LW and SW on two streams) under several environment
configurations of the library in ONE process (columns generated once; the library is finalised and re-initialised per
configuration, which is when it reads its RRTMGX_* knobs).

    python tools/sweep.py NCOL NLAY "NAME:VAR=VAL,VAR=VAL" "NAME2:..." [--profile] [--only lw|sw]
Prints one JSON line per configuration: ms per step (both / LW alone / SW alone) and, with --profile, the per-kernel
device times of one serialised step grouped by kernel family."""
import json
import os
import re
import sys

import torch

sys.path.insert(0, ".")
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200 import devstate, host
from geosradiation_gridcomp_b200.synthetic import make_columns


def timed(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = [a for a in sys.argv[1:] if a.startswith("--")]
    ncol, nlay = int(args[0]), int(args[1])
    cfgs = args[2:] or ["default:"]
    base = make_columns(ncol, nlay, seed=20260121)
    managed = set()
    for cfg in cfgs:
        name, _, kv = cfg.partition(":")
        env = dict(x.split("=", 1) for x in kv.split(",") if x)
        for k in managed:
            os.environ.pop(k, None)
        os.environ.update(env)
        managed |= set(env)
        host.finalize()
        pkg.init()
        d = devstate.to_device(base)
        o = devstate.alloc_outputs(ncol, nlay)
        st_lw, st_sw = torch.cuda.Stream(), torch.cuda.Stream()
        lw = devstate.lw_runner(d, o, device=True, sync=False, stream=st_lw.cuda_stream)
        sw = devstate.sw_runner(d, o, device=True, sync=False, stream=st_sw.cuda_stream)
        cur = torch.cuda.current_stream()

        def run(which):
            def step():
                ev = torch.cuda.Event(); ev.record(cur)
                if "lw" in which:
                    st_lw.wait_event(ev); lw(); cur.wait_stream(st_lw)
                if "sw" in which:
                    st_sw.wait_event(ev); sw(); cur.wait_stream(st_sw)
            return step
        out = {"cfg": name, "env": env, "ncol": ncol, "nlay": nlay}
        only = [f.split("=")[1] for f in flags if f.startswith("--only=")]
        if not only:
            out["both_ms"] = timed(run(("lw", "sw")))
            out["Mcol_s"] = ncol / out["both_ms"] / 1e3
        if not only or only[0] == "lw":
            out["lw_ms"] = timed(run(("lw",)))
        if not only or only[0] == "sw":
            out["sw_ms"] = timed(run(("sw",)))
        rc = (host.lw_status(), host.sw_status())
        out["status"] = rc
        if "--profile" in flags:
            host.profile(True)
            run(("lw", "sw"))()
            torch.cuda.synchronize()
            host.profile(False)
            fam = {}
            for k, (n, ms) in host.profile_report().items():
                f = re.sub(r"<.*", "", k)
                fam[f] = fam.get(f, 0.) + ms
            out["families_ms"] = {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}
        # checksum of the fluxes: configurations must agree (to rounding of the summation order)
        torch.cuda.synchronize()
        out["sum_swdflx"] = float(o["swdflx"].sum().item())
        out["sum_swuflx"] = float(o["swuflx"].sum().item())
        out["sum_uflx"] = float(o["uflx"].sum().item())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
