"""GPU parity of the SW path against the CPU oracle, through the C ABI (host pointers)."""
import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-9     # north_star: fp64 fluxes, max relative error <= 1e-9
PROFILES = ("swuflx", "swdflx", "swuflxc", "swdflxc")
SCALARS = ("nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband", "cotdtp", "cotdhp", "cotdmp", "cotdlp",
           "cotntp", "cotnhp", "cotnmp", "cotnlp")


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))


def compare_sw(o, g, extra=()):
    np.testing.assert_array_equal(g["clearCounts"], o["clearCounts"])
    for k in PROFILES + SCALARS + tuple(extra):
        assert relerr(g[k], o[k]) <= FLUX_RTOL, (k, relerr(g[k], o[k]))


@pytest.mark.parametrize("ncol,nlay,seed", [(1024, 72, 20260120), (300, 72, 7), (96, 181, 20260123)])
def test_sw_fluxes_and_counts(rx, oracle, ncol, nlay, seed):
    s = make_columns(ncol, nlay, seed=seed)
    o = oracle.rrtmg_sw(s)
    assert o["rc"] == 0
    g = rx.run_sw(s)
    compare_sw(o, g)
    assert (o["clearCounts"][:, 0] < 112).sum() > ncol // 4    # cloudy columns are exercised


def test_sw_unnormalised_and_band_fluxes(rx, oracle):
    s = make_columns(256, 72, seed=41)
    o = oracle.rrtmg_sw(s, normFlx=0, do_drfband=True)
    g = rx.run_sw(s, normFlx=0, do_drfband=True)
    compare_sw(o, g, extra=("drband", "dfband"))
    # energy bookkeeping that does not need the oracle: the surface components add up
    tot = g["nirr"] + g["nirf"] + g["parr"] + g["parf"] + g["uvrr"] + g["uvrf"]
    assert relerr(tot, g["swdflx"][:, 0]) <= 1e-12
    assert relerr(g["fswband"].sum(axis=1), g["swdflx"][:, 0] - g["swuflx"][:, 0]) <= 1e-12
    assert relerr((g["drband"] + g["dfband"]).sum(axis=1), g["swdflx"][:, 0]) <= 1e-12


def test_sw_indices_bit_exact(rx, oracle):
    s = make_columns(512, 72, seed=11)
    names = ("jp", "jt", "jt1", "indfor", "indself", "laytrop", "fac00", "fac01", "fac10", "fac11")
    o = oracle.rrtmg_sw(s, taps=names)
    g = rx.run_sw(s, taps=names)
    for k in ("jp", "jt", "jt1", "indfor", "indself", "laytrop"):
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)
    for k in ("fac00", "fac01", "fac10", "fac11"):
        assert relerr(g[k], o[k]) <= 1e-12, k


def test_sw_mcica_mask_and_cloud_optics(rx, oracle):
    s = make_columns(384, 72, seed=5)
    o = oracle.rrtmg_sw(s, taps=("cldymc", "taucmc"))
    g = rx.run_sw(s, taps=("cldymc", "taucmc"))
    np.testing.assert_array_equal(g["cldymc"], o["cldymc"])
    assert o["cldymc"].sum() > 1000
    assert relerr(g["taucmc"], o["taucmc"]) <= 1e-13


def test_sw_gas_optics_per_band(rx, oracle):
    s = make_columns(256, 72, seed=3)
    names = ("taug", "pfracs", "ssi")      # pfracs carries the Rayleigh optical depth in SW
    o = oracle.rrtmg_sw(s, taps=names)
    g = rx.run_sw(s, taps=names)
    ngs = [6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100, 112]
    lo = 0
    for b, hi in enumerate(ngs):
        for k in ("taug", "pfracs"):
            e = relerr(g[k][:, lo:hi, :], o[k][:, lo:hi, :])
            assert e <= 1e-11, (k, b + 16, e)
        e = relerr(g["ssi"][:, lo:hi], o["ssi"][:, lo:hi])
        assert e <= 1e-13, ("ssi", b + 16, e)
        lo = hi


@pytest.mark.parametrize("kw", [
    dict(isolvar=-1), dict(isolvar=-1, bndscl=np.linspace(0.9, 1.1, 14)),
    dict(isolvar=1, solcycfrac=0.3), dict(isolvar=1, solcycfrac=0.7, indsolvar=[1.2, 0.8]),
    dict(isolvar=2, indsolvar=[0.16, 1000.0]), dict(isolvar=3, bndscl=np.linspace(1.05, 0.95, 14)),
    dict(iaer=0), dict(iceflg=1), dict(iceflg=2), dict(iceflg=4)])
def test_sw_options(rx, oracle, kw):
    s = make_columns(160, 72, seed=53)
    o = oracle.rrtmg_sw(s, normFlx=0, **kw)
    assert o["rc"] == 0
    g = rx.run_sw(s, normFlx=0, **kw)
    compare_sw(o, g)


def test_sw_scon_zero_and_homogeneous(rx, oracle):
    s = make_columns(128, 72, seed=59)
    s = dict(s)
    s["scon"] = 0.0
    try:
        for ih in (0, 2):
            oracle.set_mcica(ih)
            rx.set_inhomogeneity(ih)
            o = oracle.rrtmg_sw(s, normFlx=0)
            g = rx.run_sw(s, normFlx=0)
            compare_sw(o, g)
    finally:
        oracle.set_mcica(1)
        rx.set_inhomogeneity(1)


def test_sw_device_pointers_equal_host(rx):
    """Column chunking / device-pointer mode must not change results."""
    import torch
    from geosradiation_gridcomp_b200 import devstate
    s = make_columns(700, 72, seed=29)
    a = rx.run_sw(s)
    d = devstate.to_device(s)
    run = devstate.sw_runner(d)
    o = run()
    torch.cuda.synchronize()
    for k in PROFILES:
        np.testing.assert_array_equal(o[k].cpu().numpy().T, a[k])
    np.testing.assert_array_equal(o["clearCounts_sw"].cpu().numpy().T, a["clearCounts"])
    np.testing.assert_array_equal(o["fswband"].cpu().numpy().T, a["fswband"])


def test_sw_input_traps(rx, oracle):
    s = make_columns(64, 72, seed=31)
    bad = dict(s)
    bad["tlay"] = s["tlay"].copy(order="F")
    bad["tlay"][5, 7] = -1.0
    assert oracle.rrtmg_sw(bad)["rc"] == -103
    with pytest.raises(rx.RrtmgxError) as e:
        rx.run_sw(bad)
    assert e.value.status == -103
    with pytest.raises(rx.RrtmgxError) as e:
        rx.run_sw(s, isolvar=1)          # isolvar == 1 requires solcycfrac
    assert e.value.status == -61
    assert oracle.rrtmg_sw(s, isolvar=1)["rc"] == -61
    bad = dict(s)
    bad["rei"] = s["rei"].copy(order="F")
    bad["rei"][:, :] = 400.0
    assert oracle.rrtmg_sw(bad)["rc"] == -42
    with pytest.raises(rx.RrtmgxError) as e:
        rx.run_sw(bad)
    assert e.value.status == -42


def test_sw_heating_rate(rx, oracle):
    s = make_columns(128, 72, seed=37)
    o = oracle.rrtmg_sw(s, normFlx=0)
    g = rx.run_sw(s, normFlx=0)
    grav, cp = 9.80665, 1004.68506
    def hr(f):
        net = f["swuflx"] - f["swdflx"]
        return (net[:, :-1] - net[:, 1:]) * (grav / cp) / ((s["plev"][:, :-1] - s["plev"][:, 1:]) * 100.0) * 86400.0
    got = rx.heating_rate(np.asfortranarray(g["swuflx"] - g["swdflx"]), s["plev"], grav, cp)
    assert np.max(np.abs(got - hr(o))) <= 1e-6      # north_star: heating rates within 1e-6 K/day


def test_sw_reuse_clouds_with_and_without_aerosols(rx):
    """SORADCORE is run without and then with aerosols on one cloud state (SOL:3249-3287): the second call
    keeps the first one's McICA subcolumns and cloud optics under RRTMGX_REUSE_CLOUDS."""
    s = make_columns(2048, 72, seed=33)
    clean = rx.run_sw(s, iaer=0)
    n0 = rx.launch_count()
    fresh = rx.run_sw(s, iaer=10)
    n1 = rx.launch_count()
    rx.run_sw(s, iaer=0)
    n2 = rx.launch_count()
    reused = rx.run_sw(s, iaer=10, reuse_clouds=True)
    n3 = rx.launch_count()
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "parf", "fswband", "cotntp", "cotdlp", "clearCounts"):
        np.testing.assert_array_equal(reused[k], fresh[k], err_msg=k)
    assert n3 - n2 <= (n1 - n0) - 5
    assert np.abs(fresh["swdflx"] - clean["swdflx"]).max() > 1e-4


def test_sw_real4_arrays(rx):
    """RRTMGX_F32_ARRAYS on the SW path: same bits as the fp64 interface on the widened inputs, rounded once."""
    s = make_columns(1500, 72, seed=42)
    s32 = {k: (np.asfortranarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v)
           for k, v in s.items()}
    s64 = {k: (np.asfortranarray(v, dtype=np.float64) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v)
           for k, v in s32.items()}
    ref = rx.run_sw(s64)
    out = rx.alloc_sw_outputs(1500, 72)
    out = {k: (np.asfortranarray(v, dtype=np.float32) if v.dtype == np.float64 else v) for k, v in out.items()}
    got = rx.run_sw(s32, out=out, f32=True)
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband",
              "cotntp", "cotdtp"):
        assert got[k].dtype == np.float32
        np.testing.assert_array_equal(got[k], ref[k].astype(np.float32), err_msg=k)
    np.testing.assert_array_equal(got["clearCounts"], ref["clearCounts"])


def test_sw_clean_and_full_in_one_call(rx):
    """rrtmgx_sw_run_with_clean = the no-aerosol and the regular SORADCORE pass (SOL:3249-3287) in one call:
    the bits of the two separate calls (iaer = 0, then 10), over two staged chunks."""
    ncol, nlay = 20000, 72
    s = make_columns(ncol, nlay, seed=52)
    clean = rx.run_sw(s, iaer=0)
    full = rx.run_sw(s, iaer=10)
    na = {k: np.zeros_like(full[k], order="F") for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "fswband")}
    got = rx.run_sw(s, iaer=10, clean=na)
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "parf", "fswband", "cotntp", "clearCounts"):
        np.testing.assert_array_equal(got[k], full[k], err_msg=k)
    for k in na:
        np.testing.assert_array_equal(na[k], clean[k], err_msg="clean " + k)
    assert np.abs(clean["swdflx"] - full["swdflx"]).max() > 1e-4


@pytest.mark.parametrize("split,down", [("1", "0"), ("1", "1"), ("2", "0")])
def test_sw_split_path_streaming_downward_kernel(rx, oracle, split, down):
    """RRTMGX_SW_SPLIT=1: upward kernel + the streaming downward kernel (cp.async.bulk ring, mbarrier per stage;
    RRTMGX_SW_DOWN picks the g-points per pass) instead of the fused band kernel: every SW output against the oracle
    at the contract's tolerance, and against the fused path to the rounding of the band sums' order.  Ragged column
    count (partial last tile), cloudy and cloud-free tiles, and a deep column (six mask words).
    RRTMGX_SW_SPLIT=2: the overlapped schedule (upward kernels back to back, the downward kernels as persistent grids on
    a high-priority stream; with RRTMGX_SW_DOWN_BLOCKS=1 and 5 000 columns every block walks several tiles)."""
    import os
    cases = [(make_columns(5000 if split == "2" else 1000, 72, seed=61), None), (make_columns(70, 181, seed=67), None)]
    fused = [rx.run_sw(s, normFlx=0, do_drfband=True) for s, _ in cases]
    saved = {k: os.environ.get(k) for k in ("RRTMGX_SW_SPLIT", "RRTMGX_SW_DOWN", "RRTMGX_SW_DOWN_BLOCKS")}
    os.environ["RRTMGX_SW_SPLIT"], os.environ["RRTMGX_SW_DOWN"] = split, down
    if split == "2":
        os.environ["RRTMGX_SW_DOWN_BLOCKS"] = "1"
    rx.finalize()
    rx.init()
    try:
        for (s, _), f in zip(cases, fused):
            g = rx.run_sw(s, normFlx=0, do_drfband=True)
            o = oracle.rrtmg_sw(s, normFlx=0, do_drfband=True)
            compare_sw(o, g, extra=("drband", "dfband"))
            for k in PROFILES + SCALARS:
                assert relerr(g[k], f[k]) <= 1e-12, (k, relerr(g[k], f[k]))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        rx.finalize()
        rx.init()
