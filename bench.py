#!/usr/bin/env python
"""bench.py -- columns/s of RRTMG LW+SW with McICA on synthetic C180 L72 columns.

One "step" = one rrtmg_lw call + one rrtmg_sw call (McICA subcolumns, cloud optics, gas
optics, radiative transfer) over the same batch of columns, i.e. one radiation refresh of the
GEOS Run phase.  The workload at N=1 is BASELINE.json configs[2]: the full C180
cube-sphere, 194 400 columns x 72 layers; at N>1 every rank owns its own 194 400-column slab
of a larger grid (weak scaling, no data-path collective; columns are independent).

  value     whole-job columns/s with every boundary array already resident in HBM
  e2e       the same metric through the C ABI with HOST (pinned) arrays: the library stages
            chunks host->device, computes, and copies the fluxes back inside the timed region
  roofline  the DOMINANT KERNEL (the slowest SW band kernel: gas optics + two-stream + adding fused): algorithmic
            boundary bytes of one launch / its average launch duration in this run (CUDA events on its own
            stream) against the measured HBM copy bandwidth; `traffic` = its measured DRAM bytes per launch;
            `fp64` = its FP64 thread-instructions/s against the measured DADD/DMUL and DFMA issue rates;
            `step` = the same for the whole refresh (59 560 algorithmic B per column at L72)
  cpu_baseline  the C oracle (oracle/, the CPU restatement of the reference Fortran; the
            reference itself cannot be compiled here: no Fortran compiler) on a bounded sample

`--impl reference` times the oracle port alone on the host cores (rank 0 only).
`--config 4|5` runs BASELINE.json's configs[3] / [4] instead (C720 x L72, C360 x L181: the whole grid is cut into one
contiguous column slab per rank, strong scaling); every run ends with a verification that is not timed: each rank's
first columns are gathered on rank 0 (NCCL when N > 1) and compared bit for bit with rank 0's own recomputation of
the same global columns through the host-array interface.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "columns/s RRTMG LW+SW (C180 L72, McICA)"
UNIT = "columns/s"


def algorithmic_bytes_per_column(nlay):
    """every boundary input read once, every output written once, fp64."""
    lw = (36 * nlay + 20) * 8 + (6 * (nlay + 1) + 32) * 8 + 16
    sw = (56 * nlay + 7) * 8 + (4 * (nlay + 1) + 28) * 8 + 16
    return lw, sw


def _gen_slab(args):
    from geosradiation_gridcomp_b200.synthetic import make_columns
    ncol, nlay, seed, col0 = args
    return make_columns(ncol, nlay, seed=seed, col0=col0)


def make_state(ncol, nlay, seed, col0, workers):
    """synthetic.make_columns for [col0, col0+ncol), generated in slabs by a process pool (the
    generator is a pure function of (seed, global column index))."""
    from geosradiation_gridcomp_b200.synthetic import make_columns
    slab = 8192
    if ncol <= slab or workers <= 1:
        return make_columns(ncol, nlay, seed=seed, col0=col0)
    jobs = [(min(slab, ncol - c), nlay, seed, col0 + c) for c in range(0, ncol, slab)]
    import multiprocessing as mp
    # spawn, not fork: callers may already hold a CUDA context and helper threads
    with mp.get_context("spawn").Pool(min(workers, len(jobs))) as pool:
        parts = pool.map(_gen_slab, jobs)
    out = dict(parts[0])
    for k, v in parts[0].items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.ndim >= 1 and v.shape[0] == parts[0]["ncol"]:
            out[k] = np.asfortranarray(np.concatenate([p[k] for p in parts], axis=0))
    out["ncol"] = ncol
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.2)] or \
               [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(sample_cols, nlay, seed, with_sw):
    """columns/s of the oracle port (all host threads) on a bounded sample of the workload."""
    from oracle import binding as oracle
    s = make_state(sample_cols, nlay, seed, 0, min(16, os.cpu_count() or 1))
    oracle.lib()
    warm = make_state(min(1024, sample_cols), nlay, seed, 0, 1)   # page in tables, spin up the OpenMP team
    oracle.rrtmg_lw(warm)
    if with_sw:
        oracle.rrtmg_sw(warm)
    t0 = time.perf_counter()
    r = oracle.rrtmg_lw(s)
    assert r["rc"] == 0
    if with_sw:
        r = oracle.rrtmg_sw(s)
        assert r["rc"] == 0
    dt = time.perf_counter() - t0
    return sample_cols / dt, oracle.num_threads(), dt


def oracle_has_sw():
    from oracle import binding as oracle
    return hasattr(oracle.lib(), "oracle_rrtmg_sw") and os.path.exists(os.path.join(ROOT, "oracle", "sw.c"))


def run_reference(a):
    """--impl reference: the CPU implementation of the path (oracle port; the Fortran reference
    cannot be built in this image) with all host threads, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 per rank; rank 0 alone runs this arm, with every host thread
    if "RANK" in os.environ or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    with_sw = oracle_has_sw()
    sample = a.cpu_sample
    rates, secs = [], []
    for i in range(a.warmup + a.steps):
        r, threads, dt = cpu_oracle_rate(sample, a.nlay, a.seed + i, with_sw)
        if i >= a.warmup:
            rates.append(r); secs.append(dt)
    v = sample * len(secs) / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(a, with_sw), cpu_sample_columns_per_step=sample,
                       note="each step of this arm is a bounded sample of the workload (cpu_sample_columns_per_step "
                            "columns), the rate is per column"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                         "sample": f"{sample} columns x L{a.nlay} per step, {'LW+SW' if with_sw else 'LW only'}, "
                                   f"OpenMP over column partitions"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


GRIDS = {3: ("C180 cube-sphere", 6 * 180 * 180, 72), 4: ("C720 cube-sphere", 6 * 720 * 720, 72),
         5: ("C360 cube-sphere", 6 * 360 * 360, 181)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def workload_config(a, with_sw, world=1):
    name, total, _ = GRIDS[a.config]
    if a.config == 3 and a.ncol == total:
        what = f"the full {name} ({a.ncol} columns) x L{a.nlay} per GPU"
    elif a.config in (4, 5) and a.ncol * world >= total:
        what = f"{name} ({total} columns) x L{a.nlay} cut into {world} contiguous slabs of {a.ncol} columns, one per GPU"
    else:
        what = f"a slab of {a.ncol} synthetic columns x L{a.nlay} per GPU (not a whole grid)"
    return {"workload": f"RRTMG {'LW+SW' if with_sw else 'LW (SW not built)'} with McICA, {what}, "
                        f"all columns sunlit, 40% clear-sky columns",
            "baseline_config": a.config, "ncol_per_gpu": a.ncol, "nlay": a.nlay, "ngpt_lw": 140, "ngpt_sw": 112,
            "l2_policy": "inputs (>10 GB per step) far exceed the 126 MB L2; no explicit flush",
            "precision": "fp64 boundary arrays and arithmetic (promoted-real contract)",
            "paths": getattr(a, "paths", "concurrent")}


def bind_near_gpu(local):
    """Multi-rank runs: keep this rank's host threads (and, by first touch, its pinned staging arrays) on the
    NUMA node its GPU hangs off, so eight ranks do not pull their host arrays across the socket link.
    Returns the node or None when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE.json config: 3 = C180 L72 per GPU (weak scaling, the default), 4 = C720 L72 and "
                         "5 = C360 L181 cut into one slab per GPU (strong scaling)")
    ap.add_argument("--ncol", type=int, default=None, help="columns per GPU (default: from --config)")
    ap.add_argument("--nlay", type=int, default=None)
    ap.add_argument("--verify-cols", type=int, default=256, help="columns per rank checked after the timed region")
    ap.add_argument("--seed", type=int, default=20260121)
    ap.add_argument("--cpu-sample", type=int, default=65536, help="columns in the CPU baseline sample")
    ap.add_argument("--paths", default="concurrent", choices=["concurrent", "serial"],
                    help="device-resident arm: LW and SW of a step on two streams at once (default) or SW after LW")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = max(a.warmup, 1)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    _, total, lay = GRIDS[a.config]
    if a.nlay is None:
        a.nlay = lay
    if a.ncol is None:
        a.ncol = total if a.config == 3 else -(-total // max(world_env, a.gpus if a.impl == "reference" else 1))

    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import geosradiation_gridcomp_b200 as pkg
    from geosradiation_gridcomp_b200 import devstate, host
    pkg.init(device=local)
    with_sw = os.path.exists(os.path.join(ROOT, "geosradiation_gridcomp_b200", "csrc", "sw.cu"))

    ncol, nlay = a.ncol, a.nlay
    workers = max(1, (os.cpu_count() or 8) // max(world, 1))
    s = make_state(ncol, nlay, a.seed, rank * ncol, min(workers, 16))

    # ---- device-resident arm ------------------------------------------------------------------
    d = devstate.to_device(s)
    o = devstate.alloc_outputs(ncol, nlay)
    st_lw, st_sw = torch.cuda.Stream(), torch.cuda.Stream()
    run_lw = devstate.lw_runner(d, o, device=True, sync=False, stream=st_lw.cuda_stream)
    run_sw = devstate.sw_runner(d, o, device=True, sync=False, stream=st_sw.cuda_stream) if with_sw else None
    cur = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(cur)
        st_lw.wait_event(ev)
        run_lw()
        if run_sw:
            if a.paths == "serial":
                st_sw.wait_stream(st_lw)   # experiment (profiles/t1_j_*): one path at a time
            else:
                st_sw.wait_event(ev)
            run_sw()
        cur.wait_stream(st_lw)
        if run_sw:
            cur.wait_stream(st_sw)

    def check():
        rc = host.lw_status()
        if rc:
            raise SystemExit(f"rrtmgx_lw_run failed: {rc}")
        if run_sw:
            rc = host.sw_status()
            if rc:
                raise SystemExit(f"rrtmgx_sw_run failed: {rc}")

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    check()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = host.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(cur)
    for _ in range(a.steps):
        step()
    e1.record(cur)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    check()
    launches = host.launch_count() - n0
    dev_ms = e0.elapsed_time(e1)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    value = world * ncol * a.steps / (dev_ms * 1e-3)

    # ---- verification, not timed: every rank's first columns, gathered on rank 0 (NCCL when world > 1), against rank
    # 0's own recomputation of those global columns through the HOST-array interface (other chunking, other leading
    # dimension, other column grouping): bit for bit.  synthetic.make_columns is a pure function of (seed, global column).
    verify = None
    if a.verify_cols > 0:
        from geosradiation_gridcomp_b200 import sharding
        from geosradiation_gridcomp_b200.synthetic import make_columns
        K = min(a.verify_cols, ncol)
        names = ("uflx", "dflx", "swdflx", "swuflx") if with_sw else ("uflx", "dflx")
        got = {}
        for k in names:
            local = np.asfortranarray(o[k][:, :K].t().contiguous().cpu().numpy())   # (K, nlay+1)
            got[k] = sharding.gather_columns(local, K * world, dist if world > 1 else None)
        if rank == 0:
            worst, same = 0.0, True
            for r in range(world):
                sr = make_columns(K, nlay, seed=a.seed, col0=r * ncol)
                ref = dict(host.run_lw(sr))
                if with_sw:
                    ref.update(host.run_sw(sr))
                for k in names:
                    g = got[k][r * K:(r + 1) * K]
                    same = same and np.array_equal(g, ref[k])
                    worst = max(worst, float(np.max(np.abs(g - ref[k]))))
            verify = {"ranks": world, "columns_per_rank": K, "arrays": list(names), "bit_exact": bool(same),
                      "max_abs_diff": worst, "transport": "nccl all_gather" if world > 1 else "none (one rank)",
                      "against": "rank 0 recomputing the same global columns from host arrays"}
            if not same:
                print(f"bench.py: verification FAILED: {verify}", file=sys.stderr)

    # one extra, untimed step with the library's per-kernel CUDA-event timing (events on each
    # kernel's own stream; serialises the step) to attribute device time to kernels
    kernels = None
    if rank == 0:
        host.profile(True)
        step()
        torch.cuda.synchronize()
        host.profile(False)
        rep = host.profile_report()
        tot = sum(ms for _, ms in rep.values()) or 1.0
        top = sorted(rep.items(), key=lambda kv: -kv[1][1])[:6]
        kernels = {"serialised_step_ms": tot,
                   "top": [{"name": k, "launches_per_step": n, "avg_launch_ms": ms / n, "share": ms / tot}
                           for k, (n, ms) in top]}
        # the dominant kernel family on its own: the SW band kernels (taumol + reftra + vrtqdr fused).  Algorithmic
        # bytes per (column, launch) = what the kernel reads and writes at its boundary: 14 setcoef planes + the
        # packed indices and 3 aerosol planes per layer, 4 partial flux profiles, the surface sums, 5 column scalars;
        # the two-sweep scratch it streams through HBM besides (profiles/traffic.json) is what `traffic` shows.
        sw = {k: v for k, v in rep.items() if k.startswith("sw_band_kernel")}
        if sw:
            name, (n, ms) = max(sw.items(), key=lambda kv: kv[1][1])
            cols = ncol / n                                  # columns of one launch (one chunk)
            abytes = ((14 * 8 + 4) + 3 * 8) * nlay + 4 * 8 * (nlay + 1) + 10 * 8
            kernels["dominant"] = {"name": name, "family_share": sum(v[1] for v in sw.values()) / tot,
                                   "avg_launch_ms": ms / n, "columns_per_launch": cols,
                                   "algorithmic_bytes_per_column": abytes,
                                   "achieved_gbs": abytes * cols / (ms / n * 1e-3) / 1e9}

    if world > 1:
        dist.barrier()   # rank 0's serialised profile step above runs on an otherwise idle node
    # ---- end-to-end arm: host (pinned) arrays through the C ABI --------------------------------
    e2e = None
    if not a.no_e2e:
        del d
        torch.cuda.empty_cache()
        hp = devstate.to_device(s, pinned=True)
        ho = devstate.alloc_outputs(ncol, nlay, pinned=True)
        h_lw = devstate.lw_runner(hp, ho, device=False)
        h_sw = devstate.sw_runner(hp, ho, device=False) if with_sw else None
        th = None

        def timed_e2e(h_lw, h_sw):
            def e2e_step():
                # LW and SW calls are independent; issue them from two host threads (the library
                # pipelines H2D / compute / D2H per path on its own streams)
                if h_sw:
                    err = []

                    def sw_call():
                        try:
                            h_sw()
                        except BaseException as e:   # re-raised on the main thread below
                            err.append(e)
                    t = threading.Thread(target=sw_call)
                    t.start()
                    h_lw()
                    t.join()
                    if err:
                        raise err[0]
                else:
                    h_lw()
            e2e_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            n = max(1, min(a.steps, 5))
            t0 = time.perf_counter()
            for _ in range(n):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return world * ncol * n / float(tt.item()), n

        rate, n_e2e = timed_e2e(h_lw, h_sw)
        lwb, swb = algorithmic_bytes_per_column(nlay)
        h2d = ((36 * nlay + 20) * 8 + ((56 * nlay + 7) * 8 if with_sw else 0)) * ncol
        d2h = ((6 * (nlay + 1) + 32) * 8 + 16 + (((4 * (nlay + 1) + 28) * 8 + 16) if with_sw else 0)) * ncol
        e2e = {"value": rate, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "how": "C-ABI calls with pinned host arrays; chunked H2D, kernels and D2H inside the timed region"}
        # for information, not the headline: the same calls with real*4 host arrays (RRTMGX_F32_ARRAYS, the
        # production kind of GEOS): half the bytes over PCIe, the same fp64 kernels
        del hp, ho, h_lw, h_sw
        hp4 = devstate.to_device(s, pinned=True, real4=True)
        ho4 = devstate.alloc_outputs(ncol, nlay, pinned=True, real4=True)
        r4, _ = timed_e2e(devstate.lw_runner(hp4, ho4, device=False, f32=True),
                          devstate.sw_runner(hp4, ho4, device=False, f32=True) if with_sw else None)
        e2e["real4_host_arrays"] = {"value": r4, "unit": UNIT, "h2d_bytes_per_step": int(h2d // 2),
                                    "note": "RRTMGX_F32_ARRAYS: arrays widened on the device, arithmetic fp64; "
                                            "informational, the headline e2e above moves fp64 arrays"}
        del hp4, ho4
        # the interface GEOS would really call: one fused refresh per path from the drivers' NATIVE state (top-down,
        # Pa, kg/kg, real*4), rrtmgx_irrad_refresh / rrtmgx_solar_refresh: flip, units, TLEV, ZM, aerosol
        # normalisation, RRTMG, unflip and the cloud-fraction / COT epilogue all on the device (IRR:3237-3547,
        # SOL:6113-6447).  Fewer and narrower arrays cross PCIe; same fp64 kernels.  On a bounded slab (the
        # native state is generated single-threaded), rate per column.
        try:
            from geosradiation_gridcomp_b200.synthetic import make_native_state
            ng = min(ncol, 65536)
            nat = make_native_state(ng, nlay, seed=a.seed, col0=rank * ncol)
            L = nlay

            def pin(v):
                t = torch.from_numpy(np.ascontiguousarray(v.T if v.ndim > 1 else v)).to(torch.float32).pin_memory()
                return t
            keep = {k: pin(v) for k, v in nat.items() if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.size >= ng}
            n4 = dict(nat)
            n4.update({k: t.data_ptr() for k, t in keep.items()})
            zo = lambda *sh: torch.zeros(tuple(reversed(sh)), dtype=torch.float32).pin_memory()
            from geosradiation_gridcomp_b200.host import _IRR_OUT, _SOL_OUT
            irr_o = {k: zo(ng, L + 1) for k in _IRR_OUT[:6]}
            irr_o.update({k: zo(ng) for k in _IRR_OUT[6:11]})
            irr_o["olrb"], irr_o["dolrb_dts"] = zo(16, ng), zo(16, ng)
            sol_o = {k: zo(ng, L + 1) for k in _SOL_OUT[:4]}
            sol_o.update({k: zo(ng) for k in _SOL_OUT[4:10] + _SOL_OUT[11:]})
            sol_o["fswband"] = zo(ng, 14)
            irr_p = {k: t.data_ptr() for k, t in irr_o.items()}
            sol_p = {k: t.data_ptr() for k, t in sol_o.items()}

            def glue_step():
                err = []

                def sw_call():
                    try:
                        host.solar_refresh(n4, out=sol_p, f32=True)
                    except BaseException as e:
                        err.append(e)
                t = threading.Thread(target=sw_call) if with_sw else None
                if t:
                    t.start()
                host.irrad_refresh(n4, out=irr_p, f32=True)
                if t:
                    t.join()
                if err:
                    raise err[0]
            glue_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ngl = max(1, min(a.steps, 5))
            t0 = time.perf_counter()
            for _ in range(ngl):
                glue_step()
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            in_b = ((15 * L + (L + 1) + 4 + 32 * L) + ((10 * L + (L + 1) + 7 + 42 * L) if with_sw else 0)) * 4
            out_b = ((6 * (L + 1) + 5 + 32) + ((4 * (L + 1) + 6 + 14 + 8) if with_sw else 0)) * 4
            e2e["native_real4_glue"] = {
                "value": world * ng * ngl / float(tt.item()), "unit": UNIT, "columns_per_gpu": ng,
                "h2d_bytes_per_column": in_b, "d2h_bytes_per_column": out_b,
                "note": "rrtmgx_irrad_refresh + rrtmgx_solar_refresh from the drivers' native real*4 state (pinned host "
                        "arrays), glue fused on the device; informational, the headline e2e above moves the fp64 "
                        "argument lists of rrtmg_lw / rrtmg_sw"}
            del keep, irr_o, sol_o
        except Exception as e:   # informational arm: never takes the bench line down
            e2e["native_real4_glue"] = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    lwb, swb = algorithmic_bytes_per_column(nlay)
    bpc = lwb + (swb if with_sw else 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = (value / world) * bpc / 1e9
    # Counters of the kernels (DRAM bytes, FP64 thread-instructions per column and launch) come from the committed
    # ncu launch list of this bench command (profiles/kernel_counters.json, tools/ncu_counters.py; a run under the
    # profiler cannot be a timing run): they are COUNTS, independent of the run; every TIME below is this run's own.
    counters, fp64_peak = None, None
    try:
        counters = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    except (OSError, ValueError):
        pass
    try:
        fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
    except (OSError, ValueError):
        pass
    step_s = dev_ms / a.steps * 1e-3
    traffic = None
    fp64 = None
    cnt_src = None
    if counters and nlay == 72:
        cnt_src = (f"profiles/kernel_counters.json (ncu launch list of commit {counters.get('commit_of_capture')}, "
                   f"{counters.get('columns_per_launch')} columns per launch) x this run's columns; times are this run's")
        traffic = counters["step"]["dram_bytes_per_column"] * ncol
        if fp64_peak:
            inst = counters["step"]["fp64_thread_instructions_per_column"] * ncol
            pk = float(fp64_peak["mix_dmul_dadd_register_operands"])
            fp64 = {"achieved_instr_s": inst / step_s, "peak": pk, "frac": inst / step_s / pk,
                    "unit": "fp64 thread-instructions/s",
                    "peak_source": "profiles/fp64_peak.json: DMUL/DADD with register operands, measured on a B200 of this "
                                   "pool (tools/fp64_peak.py); DFMA on three distinct register pairs peaks at "
                                   f"{float(fp64_peak['dfma_register_operands']):.3e}",
                    "counts_from": cnt_src}
    if kernels and "dominant" in kernels:
        dom = kernels["dominant"]
        dom["frac"] = dom["achieved_gbs"] / peak
        kc = (counters or {}).get("kernels", {}).get(dom["name"]) if nlay == 72 else None
        if kc:
            launch_s = dom["avg_launch_ms"] * 1e-3
            dom["traffic"] = kc["dram_bytes_per_column"] * dom["columns_per_launch"]
            dom["traffic_gbs"] = dom["traffic"] / launch_s / 1e9
            dom["traffic_frac"] = dom["traffic_gbs"] / peak
            if fp64_peak:
                inst = kc["fp64_thread_instructions_per_column"] * dom["columns_per_launch"]
                pk = float(fp64_peak["mix_dmul_dadd_register_operands"])
                pk_fma = float(fp64_peak["dfma_register_operands"])
                dom["fp64"] = {"achieved_instr_s": inst / launch_s, "peak": pk, "frac": inst / launch_s / pk,
                               "peak_dfma": pk_fma, "frac_of_dfma_peak": inst / launch_s / pk_fma,
                               "unit": "fp64 thread-instructions/s",
                               "pipe_active_pct_under_ncu": kc.get("fp64_pipe_active_pct"),
                               "peak_source": "profiles/fp64_peak.json (tools/fp64_peak.py on a B200 of this pool): "
                                              "DADD/DMUL with register operands; DFMA on three distinct register pairs"}
            dom["counts_from"] = cnt_src
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
    # whole step: the fixed pipeline of LW+SW kernels of one refresh; algorithmic bytes = boundary inputs read once +
    # outputs written once
    step_roof = {"achieved": achieved, "frac": achieved / peak, "unit": "GB/s", "algorithmic_bytes_per_column": bpc,
                 "traffic": traffic, "traffic_source": cnt_src, "fp64": fp64}
    if traffic is not None:   # the measured DRAM bytes over the measured step: how busy HBM really is (scratch included)
        step_roof["traffic_gbs"] = traffic / step_s / 1e9
        step_roof["traffic_frac"] = step_roof["traffic_gbs"] / peak
    dom = (kernels or {}).get("dominant")
    if dom:
        # the contract's roofline object: the DOMINANT KERNEL, algorithmic bytes of one launch over its average launch
        # duration of THIS run (CUDA events on the kernel's own stream); `traffic` = its measured DRAM bytes per launch
        roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                    "traffic": dom.get("traffic"), "kernel": dom["name"],
                    "kernel_family_share_of_step": dom["family_share"], "avg_launch_ms": dom["avg_launch_ms"],
                    "columns_per_launch": dom["columns_per_launch"],
                    "algorithmic_bytes_per_column_and_launch": dom["algorithmic_bytes_per_column"],
                    "traffic_gbs": dom.get("traffic_gbs"), "traffic_frac": dom.get("traffic_frac"),
                    "traffic_source": dom.get("counts_from"), "fp64": dom.get("fp64")}
    else:   # no per-kernel attribution (not rank 0 of the profile step): the whole step
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "whole step", "fp64": fp64}
    roofline.update({"peak_source": peak_src, "step": step_roof, "kernels": kernels,
                     "note": "an FP64 code: HBM is the roof the contract names, the FP64 issue rate the one closer to "
                             "binding (`fp64`: thread-instructions/s against the measured DADD/DMUL and DFMA rates); "
                             "neither is reached, the band kernels are latency-bound recurrences (DESIGN.md section 4)"})
    cpu = None
    if not a.no_cpu and world == 1:
        r, threads, dt = cpu_oracle_rate(a.cpu_sample, nlay, a.seed, with_sw and oracle_has_sw())
        cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"{a.cpu_sample} columns x L{nlay} of the same synthetic workload, "
                         f"{'LW+SW' if with_sw else 'LW only'}, {dt:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(a, with_sw, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "verify": verify,
    }
    if a.config != 3:
        line["scaling"] = "strong"
    if e2e is not None and world > 1:
        e2e["host_numa_node_rank0"] = numa   # ranks are bound to the NUMA node of their GPU (bind_near_gpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
