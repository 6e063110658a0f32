"""Multi-GPU path on hardware (needs >= 2 GPUs on the box; skipped on the 1-GPU boxes the driver's `-m gpu` run
uses).  One process per GPU under torch.distributed.run, NCCL: every rank runs its own column slab through the
library and the sampled fluxes, gathered over NCCL on rank 0, must equal rank 0's recomputation of the same GLOBAL
columns bit for bit (bench.py's verification block; SURVEY.md section 8e: no collective on the data path)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs on one box")
@pytest.mark.parametrize("extra", [["--ncol", "3000"], ["--ncol", "700", "--nlay", "181"]])
def test_slabs_over_nccl_equal_a_single_rank(extra):
    n = min(_ngpu(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "1", "--warmup", "1",
           "--no-e2e", "--no-cpu", "--verify-cols", "128"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == n
    v = line["verify"]
    assert v["ranks"] == n and v["columns_per_rank"] == 128 and v["transport"] == "nccl all_gather"
    assert v["bit_exact"] is True and v["max_abs_diff"] == 0.0
