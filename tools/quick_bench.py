"""Scratch timing of the LW/SW device path (device-resident inputs); not the contract bench."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200.synthetic import make_columns
from geosradiation_gridcomp_b200 import devstate

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 72
which = sys.argv[3] if len(sys.argv) > 3 else "lw"
pkg.init()
t = time.time(); s = make_columns(ncol, nlay); print("gen", time.time() - t)
d = devstate.to_device(s)
for path in which.split("+"):
    run = devstate.lw_runner(d) if path == "lw" else devstate.sw_runner(d)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = pkg.host.launch_count()
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{path}: {ncol} cols x {nlay}: {ms:.2f} ms/call -> {ncol / ms * 1e3:.0f} col/s; launches/call {(pkg.host.launch_count()-n0)//3}")
