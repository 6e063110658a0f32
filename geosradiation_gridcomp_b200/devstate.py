"""Device-resident column state: torch CUDA tensors in the reference layout ((ncol,nlay)
Fortran order == torch [nlay][ncol] row-major) and closures that call the C ABI on them.
torch is used only for device memory and streams."""
import numpy as np
import torch

from . import host

_LW_OUT2 = ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")
_SW_OUT2 = ("swuflx", "swdflx", "swuflxc", "swdflxc")
_SW_OUT1 = ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp",
            "cotnhp", "cotnmp", "cotnlp")


def to_device(s, device="cuda", pinned=False, real4=False):
    """Copy a synthetic.make_columns dict to the device (arrays transposed views of the same bytes);
    real4: as real*4 arrays (host arrays for RRTMGX_F32_ARRAYS)."""
    d = {}
    for k, v in s.items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.size > 16:
            t = torch.from_numpy(np.ascontiguousarray(v.T))   # same memory order as the F array
            if real4:
                t = t.to(torch.float32)
            d[k] = t.pin_memory() if pinned else t.to(device)
        else:
            d[k] = v
    return d


def alloc_outputs(ncol, nlay, device="cuda", pinned=False, real4=False):
    kw = dict(dtype=torch.float32 if real4 else torch.float64, device="cpu" if pinned else device, pin_memory=pinned)
    o = {k: torch.zeros((nlay + 1, ncol), **kw) for k in _LW_OUT2 + _SW_OUT2}
    for k in _SW_OUT1:
        o[k] = torch.zeros(ncol, **kw)
    for k in ("fswband", "drband", "dfband"):
        o[k] = torch.zeros((14, ncol), **kw)
    o["olrb"] = torch.zeros((ncol, 16), **kw)
    o["dolrb_dTs"] = torch.zeros((ncol, 16), **kw)
    ikw = dict(dtype=torch.int32, device="cpu" if pinned else device, pin_memory=pinned)
    o["clearCounts_lw"] = torch.zeros((4, ncol), **ikw)
    o["clearCounts_sw"] = torch.zeros((4, ncol), **ikw)
    return o


def _ptr(t, device):
    return t if device else t.numpy().T if t.dim() == 2 else t.numpy()


def lw_runner(d, o=None, device=True, sync=True, skip_checks=False, dudTs=True, iceflg=3, liqflg=1, stream=None,
              f32=False, reuse_clouds=False):
    ncol, nlay = d["ncol"], d["nlay"]
    o = o if o is not None else alloc_outputs(ncol, nlay, pinned=not device)
    p = (lambda t: t) if device else (lambda t: t.data_ptr())

    def run():
        host.rrtmg_lw(ncol, nlay, 4, dudTs, p(d["play"]), p(d["plev"]), p(d["tlay"]), p(d["tlev"]), p(d["tsfc"]),
                      p(d["emis"]), p(d["h2ovmr"]), p(d["o3vmr"]), p(d["co2vmr"]), p(d["ch4vmr"]), p(d["n2ovmr"]),
                      p(d["o2vmr"]), p(d["cfc11vmr"]), p(d["cfc12vmr"]), p(d["cfc22vmr"]), p(d["ccl4vmr"]),
                      p(d["cldf"]), p(d["ciwp"]), p(d["clwp"]), p(d["rei"]), p(d["rel"]), iceflg, liqflg,
                      p(d["tauaer_lw"]), p(d["zm"]), p(d["alat"]), d["dyofyr"], d["cloudLM"], d["cloudMH"],
                      p(o["clearCounts_lw"]), p(o["uflx"]), p(o["dflx"]), p(o["uflxc"]), p(o["dflxc"]),
                      p(o["duflx_dTs"]), p(o["duflxc_dTs"]), d["band_output"], p(o["olrb"]), p(o["dolrb_dTs"]),
                      device=device, sync=sync, skip_checks=skip_checks, stream=stream, f32=f32, reuse_clouds=reuse_clouds)
        return o
    run.outputs = o
    return run


def sw_runner(d, o=None, device=True, sync=True, skip_checks=False, iceflg=3, liqflg=1, isolvar=0, iaer=10,
              normFlx=1, stream=None, f32=False, reuse_clouds=False, radval=None):
    """radval: a [120][ncol] tensor (device or pinned host like the other arrays) selects the SOLAR_RADVAL build."""
    ncol, nlay = d["ncol"], d["nlay"]
    o = o if o is not None else alloc_outputs(ncol, nlay, pinned=not device)
    p = (lambda t: t) if device else (lambda t: t.data_ptr())

    def run():
        host.rrtmg_sw(0, ncol, nlay, d["scon"], d["adjes"], p(d["coszen"]), isolvar, p(d["play"]), p(d["plev"]),
                      p(d["tlay"]), p(d["h2ovmr"]), p(d["o3vmr"]), p(d["co2vmr"]), p(d["ch4vmr"]), p(d["o2vmr"]),
                      iceflg, liqflg, p(d["cldf"]), p(d["ciwp"]), p(d["clwp"]), p(d["rei"]), p(d["rel"]),
                      d["dyofyr"], p(d["zm"]), p(d["alat"]), iaer, p(d["tauaer_sw"]), p(d["ssaaer"]), p(d["asmaer"]),
                      p(d["asdir"]), p(d["asdif"]), p(d["aldir"]), p(d["aldif"]), d["cloudLM"], d["cloudMH"], normFlx,
                      p(o["clearCounts_sw"]), p(o["swuflx"]), p(o["swdflx"]), p(o["swuflxc"]), p(o["swdflxc"]),
                      p(o["nirr"]), p(o["nirf"]), p(o["parr"]), p(o["parf"]), p(o["uvrr"]), p(o["uvrf"]),
                      p(o["fswband"]), p(o["cotdtp"]), p(o["cotdhp"]), p(o["cotdmp"]), p(o["cotdlp"]), p(o["cotntp"]),
                      p(o["cotnhp"]), p(o["cotnmp"]), p(o["cotnlp"]), False, p(o["drband"]), p(o["dfband"]),
                      device=device, sync=sync, skip_checks=skip_checks, stream=stream, f32=f32, reuse_clouds=reuse_clouds,
                      radval=None if radval is None else p(radval))
        return o
    run.outputs = o
    return run
