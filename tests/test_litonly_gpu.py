"""RRTMGX_LIT_ONLY: the Solar driver's daytime packing (GEOS_SolarGridComp.F90:3686-3687 `daytime = ZTH > 0.`,
PackIt / UnPackIt :7753-7799) done by the glue kernels of rrtmgx_solar_refresh.  A refresh over a grid that is partly
in the dark must equal, bit for bit, the refresh of the packed daytime columns scattered back, with UnPackIt's defaults
in the night columns - for host arrays (one and several staging chunks), device pointers and real*4 arrays."""
import os

import numpy as np
import pytest

from geosradiation_gridcomp_b200 import sharding
from geosradiation_gridcomp_b200.synthetic import make_columns, make_native_state

pytestmark = pytest.mark.gpu

FLUX = ("fsw", "fsc", "fswu", "fscu", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband")
UNDEF = ("cldts", "cldhs", "cldms", "cldls", "cottp", "cothp", "cotmp", "cotlp")


def _state(ncol, seed):
    n = make_native_state(ncol, 72, seed=seed)
    n["zt"] = np.asfortranarray(make_columns(ncol, 72, seed=seed, lit=False)["coszen"])   # about half at night
    return n


def _check(got, n, ref_packed, lit):
    ncol = n["ncol"]
    night = np.ones(ncol, dtype=bool)
    night[lit] = False
    assert 0 < night.sum() < ncol
    for k in FLUX + UNDEF:
        want = sharding.unpack_columns(ref_packed[k], lit, ncol, default=0.0 if k in FLUX else n["undef"])
        np.testing.assert_array_equal(np.asarray(got[k]), want, err_msg=k)
    assert np.abs(np.asarray(got["fsw"])[lit]).min() > 0.0


def test_lit_only_equals_the_packed_daytime_columns(rx):
    n = _state(900, 81)
    lit = sharding.lit_columns(n["zt"])
    ref = rx.solar_refresh(sharding.pack_columns(n, lit))
    got = rx.solar_refresh(n, lit_only=True)
    _check(got, n, ref, lit)
    # and the unflagged call on the same state is another computation: it runs the night columns too
    full = rx.solar_refresh(n)
    np.testing.assert_array_equal(full["fsw"][lit], got["fsw"][lit])


def test_lit_only_over_several_staging_chunks_and_real4(rx):
    n = _state(2500, 83)
    lit = sharding.lit_columns(n["zt"])
    ref = rx.solar_refresh(sharding.pack_columns(n, lit))
    n32 = {k: (np.asfortranarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v)
           for k, v in n.items()}   # the production kind: a real*4 native state (float32 keeps the sign of ZTH)
    np.testing.assert_array_equal(sharding.lit_columns(n32["zt"]), lit)
    ref4 = rx.solar_refresh(sharding.pack_columns(n32, lit), f32=True)
    saved = os.environ.get("RRTMGX_HOST_CHUNK")
    os.environ["RRTMGX_HOST_CHUNK"] = "1024"
    rx.finalize()
    rx.init()
    try:
        _check(rx.solar_refresh(n, lit_only=True), n, ref, lit)
        _check(rx.solar_refresh(n32, lit_only=True, f32=True), n32, ref4, lit)
    finally:
        if saved is None:
            os.environ.pop("RRTMGX_HOST_CHUNK", None)
        else:
            os.environ["RRTMGX_HOST_CHUNK"] = saved
        rx.finalize()
        rx.init()


def test_lit_only_with_device_pointers(rx):
    import torch
    n = _state(1500, 85)
    lit = sharding.lit_columns(n["zt"])
    ref = rx.solar_refresh(sharding.pack_columns(n, lit))
    d = {k: (torch.from_numpy(np.ascontiguousarray(v.T)).cuda() if isinstance(v, np.ndarray) and v.dtype == np.float64
             else v) for k, v in n.items()}
    got = rx.solar_refresh(d, device=True, lit_only=True)
    _check({k: v.cpu().numpy().T for k, v in got.items()}, n, ref, lit)


def test_lit_only_all_night_and_all_day(rx):
    n = _state(300, 87)
    dark = dict(n, zt=np.asfortranarray(np.zeros(300)))
    got = rx.solar_refresh(dark, lit_only=True)
    for k in FLUX:
        assert not np.asarray(got[k]).any(), k
    for k in UNDEF:
        assert (np.asarray(got[k]) == n["undef"]).all(), k
    day = make_native_state(300, 72, seed=87)
    a, b = rx.solar_refresh(day, lit_only=True), rx.solar_refresh(day)
    for k in FLUX + UNDEF:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_lit_only_refuses_no_sync(rx):
    import ctypes as C
    n = _state(64, 89)
    keep = []
    a = rx._solar_args(n, 3, 1, 0, False, keep)
    a.flags |= rx.LIT_ONLY | rx.NO_SYNC
    assert rx.lib().rrtmgx_solar_refresh(C.byref(a)) != 0
