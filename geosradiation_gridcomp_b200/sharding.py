"""Column-slab sharding of a cube-sphere grid across the GPUs of one box.

Columns are independent on the whole RRTMG path (McICA seeds come from the
column's own pressures, partitions only block for cache), so a grid is split into contiguous
column slabs, one per rank, with no collective on the data path.  torch.distributed (NCCL on
GPUs, gloo on CPU) is used only to gather fluxes for verification and to reduce timings.
"""
import numpy as np


def slab_bounds(ncol, world_size, rank):
    """[col0, col1) of `rank`'s contiguous slab: ceil(ncol/world_size) columns per rank, the
    last ranks may get fewer (or none)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    per = -(-ncol // world_size)
    col0 = min(rank * per, ncol)
    return col0, min(col0 + per, ncol)


def slab_of(state, col0, col1):
    """The [col0, col1) column slab of a boundary-array dict (arrays whose leading dimension is
    the column index are sliced; scalars and band flags pass through)."""
    ncol = state["ncol"]
    out = {}
    for k, v in state.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == ncol and v.dtype != np.int32:
            out[k] = np.asfortranarray(v[col0:col1])
        else:
            out[k] = v
    out["ncol"] = col1 - col0
    return out


def gather_columns(local, ncol, dist=None, dst=0):
    """Gather per-rank column slabs of one output array (column index first) on rank `dst`
    (verification only).  `local` is a numpy array for this rank's slab_bounds(); returns the
    full (ncol, ...) array on `dst`, None elsewhere.  With dist=None (single process) returns
    `local` unchanged."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-ncol // world)
    tail = local.shape[1:]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    # equal-sized padded slabs so one all_gather serves ragged last ranks
    pad = torch.zeros((per,) + tail, dtype=torch.from_numpy(local[:0].copy()).dtype, device=dev)
    pad[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != dst:
        return None
    out = np.empty((ncol,) + tail, dtype=local.dtype, order="F")
    for r in range(world):
        c0, c1 = slab_bounds(ncol, world, r)
        out[c0:c1] = parts[r][:c1 - c0].cpu().numpy()
    return out


def max_over_ranks(value, dist=None):
    """Max of a per-rank scalar (device time) over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- daytime columns (SW only) ---------------------------------------------------------------------------------
# The Solar driver runs the soundings with ZTH > 0 only and balances THEM across its MPI ranks
# (GEOS_SolarGridComp.F90:3686-3712: daytime mask, MAPL_BalanceCreate on NumLit; PackIt / UnPackIt :7753-7799).
# With one process per GPU the same intent is a split of the LIT-column list, not of the grid: a contiguous slab of a
# grid that is half in the dark would leave some GPUs idle.

def lit_columns(zth):
    """Indices of the daytime columns in ascending order: `daytime = ZTH > 0.` (SOL:3686) in PackIt's order."""
    return np.nonzero(np.asarray(zth) > 0.0)[0]


def lit_slab(zth, world_size, rank):
    """The daytime columns `rank` runs: a contiguous run of the lit-column list, ceil(NumLit/world_size) per rank
    (equal work per GPU wherever the terminator lies)."""
    lit = lit_columns(zth)
    c0, c1 = slab_bounds(len(lit), world_size, rank)
    return lit[c0:c1]


def pack_columns(state, index):
    """PackIt (SOL:7753-7773) of a boundary-array / native-state dict: the columns `index` of every array whose
    leading dimension is the column index; scalars pass through."""
    ncol = state["ncol"]
    out = {}
    for k, v in state.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == ncol and k != "band_output":
            out[k] = np.asfortranarray(v[index])
        else:
            out[k] = v
    out["ncol"] = int(len(index))
    return out


def unpack_columns(packed, index, ncol, default=0.0):
    """UnPackIt (SOL:7776-7797) of one output array: `packed` rows go to columns `index`, the rest get `default`."""
    packed = np.asarray(packed)
    out = np.full((ncol,) + packed.shape[1:], default, dtype=packed.dtype, order="F")
    out[index] = packed
    return out
