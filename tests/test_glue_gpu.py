"""GPU parity of the fused Run-phase glue (rrtmgx_irrad_* / rrtmgx_solar_*, csrc/glue.cuh) against the
oracle's restatement of the two GEOS drivers, through the C ABI."""
import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_native_state

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-9     # north_star: fp64 fluxes


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))


@pytest.mark.parametrize("ncol,lm,seed", [(700, 72, 21), (130, 181, 22)])
def test_irrad_prepare_bit_exact(rx, oracle, ncol, lm, seed):
    n = make_native_state(ncol, lm, seed=seed)
    o = oracle.irrad_prepare(n)
    g = rx.irrad_prepare(n)
    for k in ("play", "plev", "tlay", "tlev", "tsfc", "emis", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr",
              "o2vmr", "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "cldf", "ciwp", "clwp", "rei", "rel", "tauaer",
              "zm", "alat"):
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)
    assert (g["cloudLM"], g["cloudMH"]) == (o["cloudLM"], o["cloudMH"])


@pytest.mark.parametrize("ncol,lm,seed", [(700, 72, 23), (130, 181, 24)])
def test_solar_prepare_bit_exact(rx, oracle, ncol, lm, seed):
    n = make_native_state(ncol, lm, seed=seed)
    o = oracle.solar_prepare(n)
    g = rx.solar_prepare(n)
    for k in ("play", "plev", "tlay", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "o2vmr", "cld", "ciwp", "clwp", "rei",
              "rel", "zm", "tauaer", "ssaaer", "asmaer"):
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)
    np.testing.assert_array_equal(g["coszen"], n["zt"])
    np.testing.assert_array_equal(g["asdir"], n["albvr"])
    np.testing.assert_array_equal(g["aldif"], n["albnf"])
    assert (g["cloudLM"], g["cloudMH"]) == (o["cloudLM"], o["cloudMH"])


def test_irrad_refresh(rx, oracle):
    n = make_native_state(1024, 72, seed=25)
    s = oracle.irrad_prepare(n)
    o = oracle.rrtmg_lw(s)
    assert o["rc"] == 0
    f = oracle.irrad_finish(n, o)
    g = rx.irrad_refresh(n)
    for k in ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc", "sfcem"):
        assert relerr(g[k], f[k]) <= FLUX_RTOL, k
    for k in ("cldtt", "cldhi", "cldmd", "cldlo"):
        np.testing.assert_array_equal(g[k], f[k], err_msg=k)     # integer clear counts behind them
    bo = np.nonzero(n["band_output"])[0]
    assert relerr(g["olrb"][bo], f["olrb"][bo]) <= FLUX_RTOL
    assert relerr(g["dolrb_dts"][bo], f["dolrb_dts"][bo]) <= FLUX_RTOL


def test_solar_refresh(rx, oracle):
    n = make_native_state(1024, 72, seed=26)
    s = oracle.solar_prepare(n)
    o = oracle.rrtmg_sw(s)
    assert o["rc"] == 0
    f = oracle.solar_finish(n, o)
    g = rx.solar_refresh(n)
    for k in ("fsw", "fsc", "fswu", "fscu", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband"):
        assert relerr(g[k], f[k]) <= FLUX_RTOL, k
    for k in ("cldts", "cldhs", "cldms", "cldls"):
        np.testing.assert_array_equal(g[k], f[k], err_msg=k)
    for k in ("cottp", "cothp", "cotmp", "cotlp"):
        undef = f[k] == n["undef"]
        np.testing.assert_array_equal(g[k] == n["undef"], undef, err_msg=k)
        assert relerr(g[k][~undef], f[k][~undef]) <= FLUX_RTOL, k


def test_refresh_device_pointers_equal_host_arrays(rx):
    """Chunked staging of host arrays (3 chunks of 16 384) and the device-pointer path give the same bits."""
    import torch
    n = make_native_state(40000, 72, seed=27)
    h_lw, h_sw = rx.irrad_refresh(n), rx.solar_refresh(n)
    d = {k: (torch.from_numpy(np.ascontiguousarray(v.T)).cuda() if isinstance(v, np.ndarray) and v.dtype == np.float64
             else v) for k, v in n.items()}
    d_lw, d_sw = rx.irrad_refresh(d, device=True), rx.solar_refresh(d, device=True)
    for k in ("flxu", "flxd", "flcu", "flcd", "dfdts", "sfcem", "cldtt"):
        np.testing.assert_array_equal(d_lw[k].cpu().numpy().T, h_lw[k], err_msg=k)
    for k in ("fsw", "fsc", "fswu", "fscu", "nirr", "cottp", "cldts", "fswband"):
        np.testing.assert_array_equal(d_sw[k].cpu().numpy().T, h_sw[k], err_msg=k)


def test_irrad_update_bit_exact(rx, oracle):
    """Between-refresh linear update of the LW exports (IRR Update :3861, :3929-3990): plain IEEE sums and
    products in the reference's order, so every export is bit-identical to the oracle's."""
    n = make_native_state(900, 72, seed=28)
    f = rx.irrad_refresh(n)
    rng = np.random.default_rng(5)
    ts_int = np.asfortranarray(n["ts"])
    tsinst = np.asfortranarray(n["ts"] + rng.normal(0, 1.5, n["ncol"]))
    g = rx.irrad_update(f, ts_int, tsinst)
    o = oracle.irrad_update(f, ts_int, tsinst)
    for k, v in o.items():
        np.testing.assert_array_equal(g[k], v, err_msg=k)
    np.testing.assert_array_equal(g["flxd"], f["flxd"])
    assert (g["olr"] > 0).all()
    only = rx.irrad_update(f, ts_int, tsinst, want=("olr", "flns"))
    np.testing.assert_array_equal(only["olr"], o["olr"])


def test_irrad_refresh_real4_arrays(rx):
    """The production kind: a real*4 native state through the fused glue gives the bits of the fp64 interface
    on the widened state, rounded once."""
    n = make_native_state(800, 72, seed=29)
    n32 = {k: (np.asfortranarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v)
           for k, v in n.items()}
    n64 = {k: (np.asfortranarray(v, dtype=np.float64) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v)
           for k, v in n32.items()}
    ref = rx.irrad_refresh(n64)
    got = rx.irrad_refresh(n32, f32=True)
    for k in ("flxu", "flxd", "flcu", "flcd", "dfdts", "dfdtsc", "sfcem", "cldtt", "cldlo"):
        assert got[k].dtype == np.float32
        np.testing.assert_array_equal(got[k], ref[k].astype(np.float32), err_msg=k)
