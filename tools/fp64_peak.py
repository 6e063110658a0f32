#!/usr/bin/env python
"""Measures the FP64 issue roof of the GPU (tools/probes/fp64_peak.cu: DFMA / DADD / DMUL / alternating DMUL-DADD loops)
and writes profiles/fp64_peak.json.  The product kernels are built with --fmad=false, so their roof is the "mix" figure
(one FP64 instruction per lane and issue slot, no contraction); `dfma_flops` is the datasheet-style 2 flop/FMA number."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "probes", "fp64_peak.cu")
EXE = os.path.join(ROOT, "tools", "probes", "fp64_peak")


def build():
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "-o", EXE, SRC])


def measure():
    build()
    out = subprocess.check_output([EXE], text=True)
    return json.loads(out.strip().splitlines()[-1])


if __name__ == "__main__":
    if "--build-only" in sys.argv:
        build()
        sys.exit(0)
    r = measure()
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fp64_peak.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(r, f, indent=1)
    print(json.dumps(r))
