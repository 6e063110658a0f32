"""End-to-end (host arrays through the C ABI) timing only; RRTMGX_HOST_CHUNK picks the staging chunk."""
import sys, time, threading
import torch
sys.path.insert(0, ".")
import bench
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200 import devstate

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 194400
pkg.init()
s = bench.make_state(ncol, 72, 20260121, 0, 16)
hp = devstate.to_device(s, pinned=True)
ho = devstate.alloc_outputs(ncol, 72, pinned=True)
h_lw = devstate.lw_runner(hp, ho, device=False)
h_sw = devstate.sw_runner(hp, ho, device=False)
def step():
    t = threading.Thread(target=h_sw); t.start(); h_lw(); t.join()
step(); torch.cuda.synchronize()
for name, fn in (("lw+sw", step), ("lw", h_lw), ("sw", h_sw)):
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt*1e3:.1f} ms/step -> {ncol/dt:.0f} col/s")
