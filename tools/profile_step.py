"""Per-kernel device times of one LW+SW step (CUDA events around every launch, serialised).

    python tools/profile_step.py [ncol] [nlay] [steps]      (RRTMGX_LW_GN / RRTMGX_SW_GN pick variants)
Prints one JSON line: per kernel launches and ms per step."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200 import devstate, host
from geosradiation_gridcomp_b200.synthetic import make_columns


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 72
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    pkg.init()
    base = make_columns(ncol, nlay, seed=20260121)   # all columns distinct (a tiled slab would repeat table rows)
    d = devstate.to_device(base)
    o = devstate.alloc_outputs(ncol, nlay)
    lw, sw = devstate.lw_runner(d, o), devstate.sw_runner(d, o)
    for _ in range(2):
        lw(); sw()
    torch.cuda.synchronize()
    host.profile(True)
    for _ in range(steps):
        lw(); sw()
    torch.cuda.synchronize()
    host.profile(False)
    rep = host.profile_report()
    tot = sum(ms for _, ms in rep.values())
    out = {"ncol": ncol, "nlay": nlay, "steps": steps, "total_ms_per_step": tot / steps,
           "kernels": {k: {"launches": n, "ms_per_step": ms / steps} for k, (n, ms) in rep.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
