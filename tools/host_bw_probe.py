"""Aggregate host -> device copy rate of ONE box when every rank pulls pinned host memory at the same time (run
under torch.distributed.run, one rank per GPU): what bounds the end-to-end (host-array) arm of bench.py at N > 1.
Each rank copies 1 GiB of pinned fp64 to its GPU `reps` times after a barrier; rank 0 prints one JSON line with the
per-rank rates alone (ranks one after the other) and together."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 27
    h = torch.empty(n, dtype=torch.float64).pin_memory(); h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float64, device="cuda")

    def timed(reps=4):
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n * 8 / (time.perf_counter() - t0) / 1e9

    together = torch.tensor([timed()], dtype=torch.float64, device="cuda")
    alone = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):          # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r] = timed()
    if world > 1:
        parts = [torch.zeros_like(together) for _ in range(world)]
        dist.all_gather(parts, together)
        dist.all_reduce(alone)
        tog = [float(p.item()) for p in parts]
    else:
        tog = [float(together.item())]
    if rank == 0:
        numa = None
        try:
            numa = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
        except OSError:
            pass
        print(json.dumps({"ranks": world, "h2d_gbs_each_rank_alone": [round(float(x), 1) for x in alone.tolist()],
                          "h2d_gbs_each_rank_all_at_once": [round(x, 1) for x in tog],
                          "h2d_gbs_aggregate_all_at_once": round(sum(tog), 1), "host_numa_nodes": numa,
                          "cpus": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
