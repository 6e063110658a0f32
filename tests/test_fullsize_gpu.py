"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would need
minutes there): a column's result does not depend on which call, chunk or slab it sits in, so a
sample of columns from the full C180 run must equal, bit for bit, a small run of the same columns,
which in turn is checked against the oracle; plus the physical identities of the path."""
import numpy as np
import pytest

from geosradiation_gridcomp_b200.synthetic import make_columns

pytestmark = pytest.mark.gpu

SEED = 20260121


def _big_state(ncol, nlay):
    import bench
    return bench.make_state(ncol, nlay, SEED, 0, 16)


@pytest.mark.parametrize("ncol,nlay", [(194400, 72), (20000, 181)])
def test_full_grid_equals_small_runs_and_oracle(rx, oracle, ncol, nlay):
    import torch
    from geosradiation_gridcomp_b200 import devstate
    s = _big_state(ncol, nlay)
    d = devstate.to_device(s)
    o = devstate.alloc_outputs(ncol, nlay)
    devstate.lw_runner(d, o)()
    devstate.sw_runner(d, o)()
    torch.cuda.synchronize()
    big = {k: v.cpu().numpy().T for k, v in o.items() if v.dim() == 2}
    # a slab from the middle of the grid, generated on its own (pure function of the column index)
    c0, n = (ncol * 5) // 8 + 37, 96
    sub = make_columns(n, nlay, seed=SEED, col0=c0)
    g_lw, g_sw = rx.run_lw(sub), rx.run_sw(sub)
    for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs"):
        np.testing.assert_array_equal(big[k][c0:c0 + n], g_lw[k], err_msg=k)
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc"):
        np.testing.assert_array_equal(big[k][c0:c0 + n], g_sw[k], err_msg=k)
    np.testing.assert_array_equal(big["clearCounts_lw"][c0:c0 + n], g_lw["clearCounts"])
    np.testing.assert_array_equal(big["clearCounts_sw"][c0:c0 + n], g_sw["clearCounts"])
    o_lw, o_sw = oracle.rrtmg_lw(sub), oracle.rrtmg_sw(sub)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30 + 1e-12 * np.max(np.abs(b)))))
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert rel(g_lw[k], o_lw[k]) <= 1e-9, k
    for k in ("swuflx", "swdflx", "swuflxc", "swdflxc"):
        assert rel(g_sw[k], o_sw[k]) <= 1e-9, k
    np.testing.assert_array_equal(g_lw["clearCounts"], o_lw["clearCounts"])
    np.testing.assert_array_equal(g_sw["clearCounts"], o_sw["clearCounts"])

    # identities over the whole grid
    clear = ~(s["cldf"] > 0).any(axis=1)
    assert clear.sum() > ncol // 4
    np.testing.assert_array_equal(big["uflx"][clear], big["uflxc"][clear])
    np.testing.assert_array_equal(big["swdflx"][clear], big["swdflxc"][clear])
    np.testing.assert_array_equal(big["clearCounts_lw"][clear], 140)
    np.testing.assert_array_equal(big["clearCounts_sw"][clear], 112)
    assert big["clearCounts_lw"].min() >= 0 and big["clearCounts_lw"].max() <= 140
    assert np.all(np.isfinite(big["uflx"])) and np.all(np.isfinite(big["swuflx"]))
    assert np.all(big["dflx"][:, -1] == 0.0)
    np.testing.assert_allclose(big["swdflx"][:, -1], 1.0, rtol=1e-14)          # normFlx
    assert np.all(big["swuflx"] >= 0) and np.all(big["swdflx"] <= 1.0 + 1e-12)
    net_toa = big["swdflx"][:, -1] - big["swuflx"][:, -1]
    net_sfc = big["swdflx"][:, 0] - big["swuflx"][:, 0]
    assert np.all(net_toa - net_sfc > 0)                                        # the atmosphere absorbs


def test_calls_from_a_second_host_thread(rx):
    """The CUDA current device is per thread; the library must bind its own (threaded LW/SW callers)."""
    import threading
    s = make_columns(128, 72, seed=3)
    ref = rx.run_sw(s)
    out = {}
    t = threading.Thread(target=lambda: out.update(rx.run_sw(s)))
    t.start(); t.join()
    np.testing.assert_array_equal(out["swuflx"], ref["swuflx"])
