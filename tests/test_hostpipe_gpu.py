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
