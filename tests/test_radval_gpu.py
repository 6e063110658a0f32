"""The SOLAR_RADVAL build of rrtmg_sw on the device (RrtmgxSwArgs::radval; the reference's compile-time option
GEOSsolar_GridComp/CMakeLists.txt:18-20: SW/src/rrtmg_sw_rad.F90:85-122, rrtmg_sw_cldprmc.F90:38-47, 321-351,
rrtmg_sw_spcvmc.F90:681-1105) against the oracle (itself pinned to the executed reference text by the radval* cases of
tests/test_refexec_pin_cpu.py; the CUDA path meets those golden vectors directly in tests/test_refexec_pin_gpu.py).

Bar: the 120 diagnostics within 1e-12 relative (sums of products of the cloud optical properties: no recurrences, so
far tighter than the flux tolerance); every regular output of the call bit-identical to a call without `radval`."""
import os

import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu
TOL = 1e-12
REGULAR = ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband", "cotdtp",
           "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp", "clearCounts")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))


@pytest.mark.parametrize("nlay,iceflg,isolvar", [(72, 3, 0), (72, 1, -1), (72, 2, 0), (72, 4, 0), (181, 3, 0)])
def test_radval_equals_oracle(rx, oracle, nlay, iceflg, isolvar):
    s = make_columns(256 if nlay == 72 else 64, nlay, seed=83)
    o = oracle.rrtmg_sw(s, iceflg=iceflg, isolvar=isolvar, radval=True)
    assert o["rc"] == 0
    g = rx.run_sw(s, iceflg=iceflg, isolvar=isolvar, radval=True)
    assert g["radval"].shape == (s["ncol"], rx.NRADVAL)
    used = np.abs(o["radval"]).max(axis=0) > 0
    assert used.all(), [n for n, u in zip(rx.RADVAL_NAMES, used) if not u]
    for q, name in enumerate(rx.RADVAL_NAMES):
        assert rel(g["radval"][:, q], o["radval"][:, q]) <= TOL, name
    # cloud-free columns hold zeros (rrtmg_sw_rad.F90:1539-1603), a "d" member is never below its "n" member's test
    clear = ~(s["cldf"] > 0).any(axis=1)
    assert clear.any() and not g["radval"][clear].any()
    # the regular outputs are those of the default build, bit for bit
    d = rx.run_sw(s, iceflg=iceflg, isolvar=isolvar)
    for k in REGULAR:
        np.testing.assert_array_equal(g[k], d[k], err_msg=k)


def test_radval_is_consistent_with_the_regular_cot_diagnostics(rx):
    """cot{l,i}* split the regular cot* by phase: where a subcolumn holds only one phase the sums coincide; always
    cotl_n + coti_n = cot_n for the whole column (same weights, tau = tau_liquid + tau_ice) up to rounding."""
    s = make_columns(512, 72, seed=5)
    g = rx.run_sw(s, radval=True)
    rv = {n: g["radval"][:, q] for q, n in enumerate(rx.RADVAL_NAMES)}
    for lev in "thml":
        tot = g["cotn" + lev + "p"]
        assert np.abs(rv["cotln" + lev + "p"] + rv["cotin" + lev + "p"] - tot).max() <= 1e-12 * max(1.0, tot.max()), lev
        assert (rv["cotld" + lev + "p"] <= g["cotd" + lev + "p"] * (1 + 1e-15)).all()
        # single-scattering albedo and asymmetry means lie in their physical range
        for ph in "li":
            d, n = rv["ssa" + ph + "d" + lev + "p"], rv["ssa" + ph + "n" + lev + "p"]
            ok = d > 0
            assert ok.any() and (n[ok] / d[ok] > 0.3).all() and (n[ok] / d[ok] <= 1.0 + 1e-12).all()
            d, n = rv["asm" + ph + "d" + lev + "p"], rv["asm" + ph + "n" + lev + "p"]
            ok = d > 0
            assert (n[ok] / d[ok] > 0.5).all() and (n[ok] / d[ok] < 1.0).all()


def test_radval_host_chunks_device_pointers_real4_and_both_passes(rx, oracle):
    """The same numbers whichever way the arrays travel: host arrays in three staging chunks, device pointers, real*4
    host arrays (outputs rounded once), and the two SORADCORE passes in one call (the no-aerosol pass generates the
    clouds and the layer sums, the regular pass reuses them)."""
    import torch
    s = make_columns(2560, 72, seed=29)
    base = rx.run_sw(s, radval=True)
    o = oracle.rrtmg_sw(s, radval=True)
    assert rel(base["radval"], o["radval"]) <= TOL
    old = os.environ.get("RRTMGX_HOST_CHUNK")
    os.environ["RRTMGX_HOST_CHUNK"] = "1024"
    try:
        rx.finalize(); rx.init()
        chunked = rx.run_sw(s, radval=True)
    finally:
        if old is None:
            os.environ.pop("RRTMGX_HOST_CHUNK")
        else:
            os.environ["RRTMGX_HOST_CHUNK"] = old
        rx.finalize(); rx.init()
    np.testing.assert_array_equal(chunked["radval"], base["radval"])
    for k in REGULAR:
        np.testing.assert_array_equal(chunked[k], base[k], err_msg=k)
    # device pointers
    ncol, nlay = s["ncol"], s["nlay"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    ds = {k: (dev(v) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v) for k, v in s.items()}
    dout = {k: dev(v) for k, v in rx.alloc_sw_outputs(ncol, nlay).items()}
    dout["radval"] = torch.zeros((rx.NRADVAL, ncol), dtype=torch.float64, device="cuda")
    rx.run_sw(ds, out=dout, radval=True, device=True)
    np.testing.assert_array_equal(dout["radval"].cpu().numpy().T, base["radval"])
    np.testing.assert_array_equal(dout["swdflx"].cpu().numpy().T, base["swdflx"])
    # the two passes of one solar refresh in one call
    clean = {k: np.zeros_like(base[k]) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc", "fswband")}
    both = rx.run_sw(s, radval=True, clean=clean)
    np.testing.assert_array_equal(both["radval"], base["radval"])
    np.testing.assert_array_equal(both["swdflx"], base["swdflx"])
    assert np.abs(clean["swdflx"] - base["swdflx"]).max() > 1e-6   # aerosols matter
    # real*4 arrays: fp64 arithmetic on the widened inputs, outputs rounded once
    s4 = {k: (np.asfortranarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v)
          for k, v in s.items()}
    s8 = {k: (np.asfortranarray(v, dtype=np.float64) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v)
          for k, v in s4.items()}
    want = rx.run_sw(s8, radval=True)
    out4 = {k: (np.asfortranarray(v, dtype=np.float32) if v.dtype == np.float64 else v)
            for k, v in rx.alloc_sw_outputs(ncol, nlay).items()}
    out4["radval"] = np.zeros((ncol, rx.NRADVAL), dtype=np.float32, order="F")
    got4 = rx.run_sw(s4, out=out4, radval=True, f32=True)
    np.testing.assert_array_equal(got4["radval"], want["radval"].astype(np.float32))
    np.testing.assert_array_equal(got4["swdflx"], want["swdflx"].astype(np.float32))


def test_radval_on_the_full_c180_grid(rx):
    """BASELINE.json's full size through size-independent properties: a slab of the 194 400-column device-resident run
    (three chunks) equals a small run of the same columns bit for bit, and the identities hold on every column."""
    import torch
    import bench
    from geosradiation_gridcomp_b200 import devstate
    ncol, nlay, seed = 194400, 72, 20260121
    s = bench.make_state(ncol, nlay, seed, 0, 16)
    d = devstate.to_device(s)
    o = devstate.alloc_outputs(ncol, nlay)
    rv = torch.zeros((rx.NRADVAL, ncol), dtype=torch.float64, device="cuda")
    devstate.sw_runner(d, o, radval=rv)()
    torch.cuda.synchronize()
    big = rv.cpu().numpy().T
    c0, n = (ncol * 3) // 8 + 11, 96
    sub = rx.run_sw(make_columns(n, nlay, seed=seed, col0=c0), radval=True)
    np.testing.assert_array_equal(big[c0:c0 + n], sub["radval"])
    np.testing.assert_array_equal(o["swdflx"].cpu().numpy().T[c0:c0 + n], sub["swdflx"])
    q = {name: big[:, i] for i, name in enumerate(rx.RADVAL_NAMES)}
    clear = ~(s["cldf"] > 0).any(axis=1)
    assert clear.sum() > ncol // 4 and not big[clear].any()
    assert np.isfinite(big).all() and (big >= 0).all()
    cot = {k: o[k].cpu().numpy() for k in ("cotntp", "cotdtp", "cotnlp", "cotdlp")}
    for lev in "tl":
        tot = cot["cotn" + lev + "p"]
        assert np.abs(q["cotln" + lev + "p"] + q["cotin" + lev + "p"] - tot).max() <= 1e-12 * tot.max()
        # a subcolumn with liquid (or ice) cloud in the super-layer is a cloudy subcolumn of it
        assert (q["cotld" + lev + "p"] <= cot["cotd" + lev + "p"] * (1 + 1e-14)).all()
        assert (q["cotid" + lev + "p"] <= cot["cotd" + lev + "p"] * (1 + 1e-14)).all()
        # delta scaling lowers the optical thickness, never the count of cloudy subcolumns
        assert (q["cdsn" + lev + "p"] <= tot * (1 + 1e-14)).all()
        np.testing.assert_array_equal(q["cdsd" + lev + "p"] > 0, cot["cotd" + lev + "p"] > 0)
