"""GPU parity of the LW path against the CPU oracle, through the C ABI (host pointers)."""
import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-9     # north_star: fp64 fluxes, max relative error <= 1e-9
HR_ATOL = 1e-6       # K/day
FLUXES = ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs")


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))


def compare_lw(o, g, nlay):
    assert g["rc"] if False else True
    np.testing.assert_array_equal(g["clearCounts"], o["clearCounts"])
    for k in FLUXES:
        assert relerr(g[k], o[k]) <= FLUX_RTOL, k
    bo = np.nonzero(np.abs(o["olrb"]).sum(axis=1))[0]
    assert relerr(g["olrb"][bo], o["olrb"][bo]) <= FLUX_RTOL
    assert relerr(g["dolrb_dTs"][bo], o["dolrb_dTs"][bo]) <= FLUX_RTOL


@pytest.mark.parametrize("ncol,nlay,seed", [(1024, 72, 20260119), (300, 72, 7), (96, 181, 20260123)])
def test_lw_fluxes_and_counts(rx, oracle, ncol, nlay, seed):
    s = make_columns(ncol, nlay, seed=seed)
    o = oracle.rrtmg_lw(s)
    assert o["rc"] == 0
    g = rx.run_lw(s)
    compare_lw(o, g, nlay)


def test_lw_indices_bit_exact(rx, oracle):
    s = make_columns(512, 72, seed=11)
    names = ("jp", "jt", "jt1", "indfor", "indself", "indminor", "laytrop", "fac00", "fac01", "fac10", "fac11",
             "pwvcm")
    o = oracle.rrtmg_lw(s, taps=names)
    g = rx.run_lw(s, taps=names)
    for k in ("jp", "jt", "jt1", "indfor", "indself", "indminor", "laytrop"):
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)
    for k in ("fac00", "fac01", "fac10", "fac11", "pwvcm"):
        assert relerr(g[k], o[k]) <= 1e-12, k


def test_lw_mcica_mask_and_cloud_optics(rx, oracle):
    s = make_columns(384, 72, seed=5)
    o = oracle.rrtmg_lw(s, taps=("cldymc", "taucmc"))
    g = rx.run_lw(s, taps=("cldymc", "taucmc"))
    # oracle taps are [icol][ig][ilay]; device taps index the same way through Fortran order
    optical = (o["taucmc"] > 0).astype(np.uint8)
    np.testing.assert_array_equal(g["cldymc"], optical)
    assert optical.sum() > 1000
    assert relerr(g["taucmc"], o["taucmc"]) <= 1e-13


def test_lw_gas_optics_per_band(rx, oracle):
    s = make_columns(256, 72, seed=3)
    o = oracle.rrtmg_lw(s, taps=("taug", "pfracs"))
    g = rx.run_lw(s, taps=("taug", "pfracs"))
    ngs = [10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140]
    lo = 0
    for b, hi in enumerate(ngs):
        for k in ("taug", "pfracs"):
            e = relerr(g[k][:, lo:hi, :], o[k][:, lo:hi, :])
            assert e <= 1e-11, (k, b + 1, e)
        lo = hi


def test_lw_homogeneous_and_gamma_condensate(rx, oracle):
    s = make_columns(256, 72, seed=17)
    try:
        for ih in (0, 2):
            oracle.set_mcica(ih)
            rx.set_inhomogeneity(ih)
            o = oracle.rrtmg_lw(s)
            g = rx.run_lw(s)
            compare_lw(o, g, 72)
    finally:
        oracle.set_mcica(1)
        rx.set_inhomogeneity(1)


def test_lw_no_derivatives_no_band_output(rx, oracle):
    s = make_columns(200, 72, seed=23)
    s["band_output"] = np.zeros(16, dtype=np.int32)
    o = oracle.rrtmg_lw(s, dudTs=False)
    g = rx.run_lw(s, dudTs=False)
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert relerr(g[k], o[k]) <= FLUX_RTOL
    assert not g["duflx_dTs"].any() and not g["olrb"].any()


def test_lw_chunked_equals_single(rx, monkeypatch):
    """Column chunking must not change results (columns are independent)."""
    s = make_columns(700, 72, seed=29)
    a = rx.run_lw(s)
    import torch
    dev = {k: torch.from_numpy(np.ascontiguousarray(v.T)).cuda() for k, v in s.items() if isinstance(v, np.ndarray) and v.dtype == np.float64}
    out = {k: torch.zeros((73, 700), dtype=torch.float64, device="cuda") for k in FLUXES}
    out["olrb"] = torch.zeros((700, 16), dtype=torch.float64, device="cuda")
    out["dolrb_dTs"] = torch.zeros((700, 16), dtype=torch.float64, device="cuda")
    cc = torch.zeros((4, 700), dtype=torch.int32, device="cuda")
    rx.rrtmg_lw(700, 72, 4, True, dev["play"], dev["plev"], dev["tlay"], dev["tlev"], dev["tsfc"], dev["emis"],
                dev["h2ovmr"], dev["o3vmr"], dev["co2vmr"], dev["ch4vmr"], dev["n2ovmr"], dev["o2vmr"],
                dev["cfc11vmr"], dev["cfc12vmr"], dev["cfc22vmr"], dev["ccl4vmr"], dev["cldf"], dev["ciwp"],
                dev["clwp"], dev["rei"], dev["rel"], 3, 1, dev["tauaer_lw"], dev["zm"], dev["alat"], s["dyofyr"],
                s["cloudLM"], s["cloudMH"], cc, out["uflx"], out["dflx"], out["uflxc"], out["dflxc"],
                out["duflx_dTs"], out["duflxc_dTs"], s["band_output"], out["olrb"], out["dolrb_dTs"], device=True)
    torch.cuda.synchronize()
    for k in FLUXES:
        np.testing.assert_array_equal(out[k].cpu().numpy().T, a[k])
    np.testing.assert_array_equal(cc.cpu().numpy().T, a["clearCounts"])
    np.testing.assert_array_equal(out["olrb"].cpu().numpy().T, a["olrb"])


def test_lw_input_traps(rx, oracle):
    s = make_columns(64, 72, seed=31)
    bad = dict(s)
    bad["tlay"] = s["tlay"].copy(order="F")
    bad["tlay"][5, 7] = -1.0
    o = oracle.rrtmg_lw(bad)
    assert o["rc"] == -103
    with pytest.raises(rx.RrtmgxError) as e:
        rx.run_lw(bad)
    assert e.value.status == -103
    bad = dict(s)
    bad["rei"] = s["rei"].copy(order="F")
    bad["rei"][:, :] = 400.0      # outside the Fu table: 'ice radius extrapolation forbidden'
    o = oracle.rrtmg_lw(bad)
    assert o["rc"] in (-46, -47)
    with pytest.raises(rx.RrtmgxError) as e:
        rx.run_lw(bad)
    assert e.value.status == -42


def test_heating_rate(rx, oracle):
    s = make_columns(128, 72, seed=37)
    o = oracle.rrtmg_lw(s)
    g = rx.run_lw(s)
    grav, cp = 9.80665, 1004.68506
    def hr(f):
        net = f["uflx"] - f["dflx"]
        return (net[:, :-1] - net[:, 1:]) * (grav / cp) / ((s["plev"][:, :-1] - s["plev"][:, 1:]) * 100.0) * 86400.0
    got = rx.heating_rate(np.asfortranarray(g["uflx"] - g["dflx"]), s["plev"], grav, cp)
    assert np.max(np.abs(got - hr(o))) <= HR_ATOL
    assert np.max(np.abs(got - hr(g))) <= 1e-9


def test_branch_free_division_equals_ieee(rx):
    """The band kernels' division (csrc/common.cuh ddiv/drcp: the compiler's fast-path sequence without
    its range test) against IEEE a/b and 1/b on the device, bit for bit, over the operand ranges of the
    band kernels: optical depths and column amounts over ~40 decades, denominators 1 - R*R' close to 1,
    and the zero numerators of aerosol-free layers."""
    rng = np.random.default_rng(20260118)
    n = 1 << 21
    a = np.concatenate([
        10.0 ** rng.uniform(-25, 15, n) * rng.choice([-1.0, 1.0], n),
        rng.uniform(0, 1, n),
        np.zeros(1 << 16),
        10.0 ** rng.uniform(-200, -100, 1 << 16),      # below the compiler's 2^-120 fast-path bound
    ])
    b = np.concatenate([
        10.0 ** rng.uniform(-25, 15, n),
        1.0 - rng.uniform(0, 1, n) * rng.uniform(0, 1, n) * (1 - 1e-9),
        10.0 ** rng.uniform(-10, 10, 1 << 16),
        10.0 ** rng.uniform(-10, 10, 1 << 16),
    ])
    qf, qi, rf, ri = rx.debug_divide(a, b)
    np.testing.assert_array_equal(qi, a / b)            # the device's IEEE division is numpy's
    np.testing.assert_array_equal(qf, qi)
    np.testing.assert_array_equal(rf, ri)


def test_lw_reuse_clouds_for_removed_gas_calls(rx):
    """GEOS calls rrtmg_lw once more per removed gas on the same cloud state (IRR:3405-3468):
    RRTMGX_REUSE_CLOUDS keeps the McICA subcolumns, cloud optics and clear counts of the previous call and
    gives the bits of a full call with fewer launches."""
    s = make_columns(2048, 72, seed=31)
    rx.run_lw(s)
    s2 = dict(s)
    s2["ch4vmr"] = np.zeros_like(s["ch4vmr"], order="F")
    n0 = rx.launch_count()
    fresh = rx.run_lw(s2)
    n1 = rx.launch_count()
    reused = rx.run_lw(s2, reuse_clouds=True)
    n2 = rx.launch_count()
    for k in FLUXES + ("olrb", "dolrb_dTs", "clearCounts"):
        np.testing.assert_array_equal(reused[k], fresh[k], err_msg=k)
    assert n2 - n1 <= (n1 - n0) - 5          # partition, prep, thresholds, cloud coefficients, McICA skipped
    assert np.abs(fresh["uflx"] - rx.run_lw(s)["uflx"]).max() > 1e-3    # and the gas did matter
    # a different shape invalidates the kept clouds: the flag is then ignored, not trusted
    s3 = make_columns(1000, 72, seed=32)
    np.testing.assert_array_equal(rx.run_lw(s3, reuse_clouds=True)["uflx"], rx.run_lw(s3)["uflx"])


def test_lw_real4_arrays(rx):
    """RRTMGX_F32_ARRAYS: real*4 boundary arrays (the production kind of GEOS) are widened exactly while staged,
    the arithmetic stays fp64, the outputs are rounded once: the same bits as the fp64 interface run on the
    widened inputs and then rounded."""
    s = make_columns(1500, 72, seed=41)
    s32 = {k: (np.asfortranarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v)
           for k, v in s.items()}
    s64 = {k: (np.asfortranarray(v, dtype=np.float64) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v)
           for k, v in s32.items()}
    ref = rx.run_lw(s64)
    out = rx.alloc_lw_outputs(1500, 72)
    out = {k: (np.asfortranarray(v, dtype=np.float32) if v.dtype == np.float64 else v) for k, v in out.items()}
    got = rx.run_lw(s32, out=out, f32=True)
    for k in FLUXES + ("olrb", "dolrb_dTs"):
        assert got[k].dtype == np.float32
        np.testing.assert_array_equal(got[k], ref[k].astype(np.float32), err_msg=k)
    np.testing.assert_array_equal(got["clearCounts"], ref["clearCounts"])


def test_lw_removed_gas_loop_in_one_call(rx):
    """rrtmgx_lw_run_variants = the removed-gas loop of LW_Driver (IRR:3405-3468) plus the main call: the same
    bits as the separate calls, over several chunks of host arrays and with device pointers."""
    import torch
    names = ("CH4", "H2O", "CFC12")
    key = {"CH4": "ch4vmr", "H2O": "h2ovmr", "CFC12": "cfc12vmr"}
    ncol, nlay = 20000, 72            # host arrays: two staged chunks
    s = make_columns(ncol, nlay, seed=51)
    sep = []
    for n in names:
        s2 = dict(s)
        s2[key[n]] = np.zeros_like(s[key[n]], order="F")
        sep.append(rx.run_lw(s2))
    main = rx.run_lw(s)
    rat = [np.zeros((ncol, nlay + 1, len(names)), order="F") for _ in range(3)]
    n0 = rx.launch_count()
    got = rx.run_lw(s, rats=(names, *rat))
    launches = rx.launch_count() - n0
    for k in FLUXES + ("olrb", "dolrb_dTs", "clearCounts"):
        np.testing.assert_array_equal(got[k], main[k], err_msg=k)
    for i, n in enumerate(names):
        np.testing.assert_array_equal(rat[0][:, :, i], sep[i]["uflx"], err_msg=n)
        np.testing.assert_array_equal(rat[1][:, :, i], sep[i]["dflx"], err_msg=n)
        np.testing.assert_array_equal(rat[2][:, :, i], sep[i]["duflx_dTs"], err_msg=n)
    assert np.abs(sep[1]["uflx"] - main["uflx"]).max() > 1.0      # removing water vapour matters
    # device pointers
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    d = {k: (dev(v) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v) for k, v in s.items()}
    o = {k: dev(v) for k, v in rx.alloc_lw_outputs(ncol, nlay).items()}
    drat = [torch.zeros((len(names), nlay + 1, ncol), dtype=torch.float64, device="cuda") for _ in range(3)]
    rx.run_lw(d, out=o, device=True, rats=(names, *drat))
    np.testing.assert_array_equal(o["uflx"].cpu().numpy().T, main["uflx"])
    for i in range(len(names)):
        np.testing.assert_array_equal(drat[0][i].cpu().numpy().T, sep[i]["uflx"])
        np.testing.assert_array_equal(drat[2][i].cpu().numpy().T, sep[i]["duflx_dTs"])
    assert launches > 0
