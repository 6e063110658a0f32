"""End-to-end (pinned host arrays through the C ABI) rate of one LW + one SW refresh as a function of the
host staging chunk and the number of staging sets (RRTMGX_HOST_CHUNK, RRTMGX_STAGES, read at rrtmgx_init), beside the raw H2D rate of the box
(tools/pcie_probe.py): how much of the link the chunk pipeline of api.cu keeps busy.  One JSON line per setting.
PROBE_F32=1: real*4 host arrays (RRTMGX_F32_ARRAYS, half the bytes: the kernels, not the link, are then the longer
stage of the pipeline, and a chunk too small to fill the SMs costs more than it does with fp64 arrays)."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from geosradiation_gridcomp_b200 import devstate, host


def main():
    ncol, nlay = int(os.environ.get("PROBE_NCOL", 194400)), 72
    s = bench.make_state(ncol, nlay, 20260121, 0, 16)
    f32 = os.environ.get("PROBE_F32", "0") == "1"
    hp = devstate.to_device(s, pinned=True, real4=f32)
    ho = devstate.alloc_outputs(ncol, nlay, pinned=True, real4=f32)
    h2d = ((36 * nlay + 20) * 8 + (56 * nlay + 7) * 8) * ncol // (2 if f32 else 1)
    chunks = [int(v) for v in os.environ.get("PROBE_CHUNKS", "4096,8192,12288,16384,24576,32768").split(",")]
    stages = [int(v) for v in os.environ.get("PROBE_STAGES", "3").split(",")]
    modes = [m == "1" for m in os.environ.get("PROBE_MODES", "1,0").split(",")]
    nsteps = int(os.environ.get("PROBE_STEPS", 3))
    for chunk, nst in [(c, n) for c in chunks for n in stages]:
        os.environ["RRTMGX_HOST_CHUNK"] = str(chunk)
        os.environ["RRTMGX_STAGES"] = str(nst)
        host.finalize()
        host.init()
        h_lw = devstate.lw_runner(hp, ho, device=False, f32=f32)
        h_sw = devstate.sw_runner(hp, ho, device=False, f32=f32)

        def step(concurrent=True):
            if concurrent:
                t = threading.Thread(target=h_sw)
                t.start(); h_lw(); t.join()
            else:
                h_lw(); h_sw()
        out = {"host_chunk": chunk, "stages": nst, "real4_arrays": f32}
        for mode in modes:
            step(mode); torch.cuda.synchronize()
            n = nsteps
            t0 = time.perf_counter()
            for _ in range(n):
                step(mode)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            key = "two_threads" if mode else "lw_then_sw"
            out[key] = {"columns_per_s": round(ncol / dt), "ms": round(dt * 1e3, 1), "h2d_gbs": round(h2d / dt / 1e9, 1)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
