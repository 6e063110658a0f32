"""The host-array pipeline of the C ABI (api.cu run_staged: H2D / kernels / D2H over staging chunks) with MANY chunks:
the other parity tests stay below one staging chunk, bench.py runs many but does not look at the numbers."""
import os

import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stages", [2, 3])
def test_host_arrays_in_many_staging_chunks_equal_the_device_run(rx, stages):
    import torch
    from geosradiation_gridcomp_b200 import devstate
    ncol, nlay = 3000, 72                       # 1024 + 1024 + 952: two full staging chunks and a ragged one
    s = make_columns(ncol, nlay, seed=47)
    d = devstate.to_device(s)
    od = devstate.alloc_outputs(ncol, nlay)
    devstate.lw_runner(d, od)()
    devstate.sw_runner(d, od)()
    torch.cuda.synchronize()
    ref = {k: v.cpu().numpy().copy() for k, v in od.items()}
    assert np.abs(ref["uflx"]).min() > 0 and np.abs(ref["swdflx"]).max() > 0
    saved = {k: os.environ.get(k) for k in ("RRTMGX_HOST_CHUNK", "RRTMGX_STAGES")}
    os.environ["RRTMGX_HOST_CHUNK"], os.environ["RRTMGX_STAGES"] = "1024", str(stages)
    rx.finalize()
    rx.init()
    try:
        hp = devstate.to_device(s, pinned=True)
        ho = devstate.alloc_outputs(ncol, nlay, pinned=True)
        devstate.lw_runner(hp, ho, device=False)()
        devstate.sw_runner(hp, ho, device=False)()
        torch.cuda.synchronize()
        for k, v in ref.items():
            np.testing.assert_array_equal(ho[k].numpy(), v, err_msg=k)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        rx.finalize()
        rx.init()


def test_real4_device_arrays_take_the_staging_path(rx):
    """RRTMGX_F32_ARRAYS | RRTMGX_DEVICE_PTRS: a caller that holds its real*4 state ON the device (a GPU-resident GEOS)
    gets the same bits as the real*4 host-array call: the arrays go through the staging pipeline with device-to-device
    copies, are widened exactly, the outputs rounded once (over three staging chunks with a ragged last one)."""
    import torch
    from geosradiation_gridcomp_b200 import devstate
    ncol, nlay = 2500, 72
    s = make_columns(ncol, nlay, seed=53)
    saved = os.environ.get("RRTMGX_HOST_CHUNK")
    os.environ["RRTMGX_HOST_CHUNK"] = "1024"
    rx.finalize()
    rx.init()
    try:
        hp = devstate.to_device(s, pinned=True, real4=True)
        ho = devstate.alloc_outputs(ncol, nlay, pinned=True, real4=True)
        devstate.lw_runner(hp, ho, device=False, f32=True)()
        devstate.sw_runner(hp, ho, device=False, f32=True)()
        dp = {k: (v.cuda() if hasattr(v, "is_pinned") else v) for k, v in hp.items()}
        do = devstate.alloc_outputs(ncol, nlay, real4=True)
        devstate.lw_runner(dp, do, device=True, f32=True)()
        devstate.sw_runner(dp, do, device=True, f32=True)()
        torch.cuda.synchronize()
        assert float(ho["uflx"].abs().min()) > 0 and float(ho["swdflx"].abs().max()) > 0
        for k, v in ho.items():
            np.testing.assert_array_equal(do[k].cpu().numpy(), v.numpy(), err_msg=k)
    finally:
        if saved is None:
            os.environ.pop("RRTMGX_HOST_CHUNK", None)
        else:
            os.environ["RRTMGX_HOST_CHUNK"] = saved
        rx.finalize()
        rx.init()
