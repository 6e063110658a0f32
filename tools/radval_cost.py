"""Device cost of the SOLAR_RADVAL build of rrtmg_sw against the default build: per-kernel CUDA-event times of one
serialised SW call over NCOL columns x L72 (device-resident arrays), with and without RrtmgxSwArgs::radval."""
import json
import sys

import torch

sys.path.insert(0, ".")
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200 import devstate, host
from geosradiation_gridcomp_b200.synthetic import make_columns


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    pkg.init()
    s = make_columns(ncol, 72, seed=20260121)
    d = devstate.to_device(s)
    o = devstate.alloc_outputs(ncol, 72)
    rv = torch.zeros((host.NRADVAL, ncol), dtype=torch.float64, device="cuda")
    runners = {False: devstate.sw_runner(d, o), True: devstate.sw_runner(d, o, radval=rv)}

    def call(radval):
        runners[radval]()

    for radval in (False, True):
        call(radval); call(radval)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            call(radval)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        host.profile(True)
        call(radval)
        torch.cuda.synchronize()
        host.profile(False)
        rep = {k: round(v[1], 3) for k, v in host.profile_report().items()
               if k.startswith(("mcica", "sw_radval", "sw_cldcoef", "sw_surface"))}
        print(json.dumps({"radval": radval, "ncol": ncol, "sw_call_ms": round(ms, 2), "kernels_ms": rep}), flush=True)


if __name__ == "__main__":
    main()
