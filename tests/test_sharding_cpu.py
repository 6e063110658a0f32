"""Column-slab sharding (the N>1 path) on CPU: world_size-2 gloo, oracle as the per-rank compute."""
import os
import socket

import numpy as np
import pytest

from geosradiation_gridcomp_b200 import sharding
from geosradiation_gridcomp_b200.synthetic import make_columns


def test_slab_bounds_cover_the_grid():
    for ncol, world in ((194400, 8), (10, 4), (7, 8), (1, 1), (3110400, 8)):
        spans = [sharding.slab_bounds(ncol, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == ncol
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
    with pytest.raises(ValueError):
        sharding.slab_bounds(10, 2, 2)


def test_generator_slabs_equal_the_full_grid():
    full = make_columns(50, 72, seed=9)
    part = make_columns(20, 72, seed=9, col0=17)
    for k in ("play", "tlay", "cldf", "tauaer_sw", "coszen", "emis"):
        np.testing.assert_array_equal(part[k], full[k][17:37])
    sl = sharding.slab_of(full, 17, 37)
    for k in ("play", "tlay", "cldf", "tauaer_sw", "coszen", "emis"):
        np.testing.assert_array_equal(sl[k], part[k])
    assert sl["ncol"] == 20 and sl["cloudLM"] == full["cloudLM"]


def _worker(rank, world, port, ncol, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as oracle
    c0, c1 = sharding.slab_bounds(ncol, world, rank)
    s = make_columns(c1 - c0, 72, seed=77, col0=c0)
    lw = oracle.rrtmg_lw(s)
    sw = oracle.rrtmg_sw(s)
    res = {}
    for k, v in (("uflx", lw["uflx"]), ("dflx", lw["dflx"]), ("swdflx", sw["swdflx"]), ("swuflx", sw["swuflx"]),
                 ("cc", lw["clearCounts"].astype(np.int32))):
        res[k] = sharding.gather_columns(np.asfortranarray(v), ncol, dist)
    tmax = sharding.max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        np.savez(out_path, tmax=tmax, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slabs_reproduce_the_single_process_result(tmp_path):
    import torch.multiprocessing as mp
    from oracle import binding as oracle
    ncol = 45                      # ragged: 23 + 22 columns
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, ncol, out), nprocs=2, join=True)
    g = np.load(out)
    s = make_columns(ncol, 72, seed=77)
    lw, sw = oracle.rrtmg_lw(s), oracle.rrtmg_sw(s)
    np.testing.assert_array_equal(g["uflx"], lw["uflx"])
    np.testing.assert_array_equal(g["dflx"], lw["dflx"])
    np.testing.assert_array_equal(g["swdflx"], sw["swdflx"])
    np.testing.assert_array_equal(g["swuflx"], sw["swuflx"])
    np.testing.assert_array_equal(g["cc"], lw["clearCounts"])
    assert float(g["tmax"]) == 11.0


def test_lit_slabs_balance_the_daytime_columns():
    """The Solar driver balances its LIT soundings across ranks (SOL:3686-3712); lit_slab splits the lit-column list,
    pack / unpack restate PackIt / UnPackIt (SOL:7753-7799)."""
    from geosradiation_gridcomp_b200 import sharding
    from geosradiation_gridcomp_b200.synthetic import make_columns
    s = make_columns(1000, 8, seed=3, lit=False)
    zth = s["coszen"]
    lit = sharding.lit_columns(zth)
    assert 300 < len(lit) < 700 and (zth[lit] > 0).all() and (np.diff(lit) > 0).all()
    parts = [sharding.lit_slab(zth, 8, r) for r in range(8)]
    np.testing.assert_array_equal(np.concatenate(parts), lit)
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 8 and max(sizes) == -(-len(lit) // 8)
    # a contiguous split of the GRID would not balance: the spread of lit columns per grid slab is wider
    grid = [int((zth[slice(*sharding.slab_bounds(1000, 8, r))] > 0).sum()) for r in range(8)]
    assert max(grid) - min(grid) > max(sizes) - min(sizes)
    p = sharding.pack_columns(s, lit)
    assert p["ncol"] == len(lit) and p["play"].shape == (len(lit), 8) and (p["coszen"] > 0).all()
    back = sharding.unpack_columns(p["play"], lit, 1000, default=-1.0)
    np.testing.assert_array_equal(back[lit], s["play"][lit])
    night = np.setdiff1d(np.arange(1000), lit)
    assert (back[night] == -1.0).all()
