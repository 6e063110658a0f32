"""The CUDA path, through the C ABI, against numbers produced by the REFERENCE'S OWN SOURCE TEXT
(tests/golden/rrtmg_refexec_golden.npz, see tests/test_refexec_pin_cpu.py for where it comes from).  The oracle is not
involved: this is product against reference output, on the cases of tests/golden/refexec_cases.py.

Bar (north_star): indices, McICA masks and clear counts bit for bit; fluxes within 1e-9 relative; intermediates
(optical depths, Planck fractions, interpolation factors) within 1e-11."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refexec_cases as rc   # noqa: E402

pytestmark = pytest.mark.gpu
TOL_FLUX, TOL_TAPS = 1e-9, 1e-11


@pytest.fixture(scope="module")
def golden():
    return np.load(rc.GOLDEN)


@pytest.mark.parametrize("name", list(rc.CASES))
def test_cuda_equals_reference_source_output(rx, golden, name):
    def set_mcica(ih, corr):
        rx._mcica["ih"], rx._mcica["corr"] = int(ih), (list(corr) if corr is not None else None)
        rx._apply_mcica()

    def reset():
        rx._mcica["ih"], rx._mcica["corr"] = 1, None
        rx._apply_mcica()
    # the device never materialises the sub-column condensate paths (the McICA kernel turns them into optical depths
    # in registers); `taucmc`, which is compared, is their product with the absorption coefficients
    got = rc.run_case(rx, name, set_mcica, reset, skip=("ciwpmc", "clwpmc"))
    n, same, worst = rc.check_case(got, golden, name, TOL_FLUX, TOL_TAPS)
    assert n >= 9
    print(f"{name}: {n} arrays, {same} bit-identical, worst relative difference {worst:.2e}")


def test_cuda_refresh_equals_the_reference_text_refresh(rx, golden):
    """rrtmgx_irrad_refresh / rrtmgx_solar_refresh (the fused driver glue around the device RRTMG path) against a whole
    refresh computed by the reference's text: its glue lines (GEOS_IrradGridComp.F90:3238-3371, 3487-3533,
    GEOS_SolarGridComp.F90:6116-6219, 6395-6454) around its RRTMG sources, golden keys refresh/*.  Prepared arguments bit
    for bit, exports within 1e-9, cloud fractions and the MAPL_UNDEF pattern exactly."""
    import make_golden_from_refexec as gen
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(gen.REFRESH_NCOL, 72, seed=gen.REFRESH_SEED)
    s = rx.irrad_prepare(n)
    for k in gen.IRR_PREPARED:
        np.testing.assert_array_equal(s["tauaer" if k == "tauaer_lw" else k], golden[f"refresh/irr_prepared/{k}"], err_msg=k)
    s = rx.solar_prepare(n)
    for k in gen.SOL_PREPARED:
        np.testing.assert_array_equal(s[k], golden[f"refresh/sol_prepared/{k}"], err_msg=k)
    f = rx.irrad_refresh(n)
    for k in gen.IRR_EXPORTS:
        ref = golden[f"refresh/irr/{k}"]
        if k.startswith("cld"):
            np.testing.assert_array_equal(f[k], ref, err_msg=k)
        else:
            assert rc.rel_err(f[k], ref) <= TOL_FLUX, k
    f = rx.solar_refresh(n)
    for k in gen.SOL_EXPORTS:
        ref = golden[f"refresh/sol/{k}"]
        if k.startswith("cld"):
            np.testing.assert_array_equal(f[k], ref, err_msg=k)
        elif k.startswith("cot"):
            np.testing.assert_array_equal(f[k] == n["undef"], ref == n["undef"], err_msg=k)
            assert rc.rel_err(np.where(ref == n["undef"], 0.0, f[k]), np.where(ref == n["undef"], 0.0, ref)) <= TOL_FLUX, k
        else:
            assert rc.rel_err(f[k], ref) <= TOL_FLUX, k


def test_cuda_heating_rates_equal_the_reference_lines(rx, golden):
    """rrtmgx_heating_rate against RADLW / RADSW computed by the parent component's own lines
    (GEOS_RadiationGridComp.F90:811, 813-814, executed from the file; golden keys refresh/rad/*) on the fluxes of the
    golden refresh: within 1e-6 K/day (north_star); host arrays and device pointers."""
    import torch
    import make_golden_from_refexec as gen
    from geosradiation_gridcomp_b200.synthetic import make_native_state
    n = make_native_state(gen.REFRESH_NCOL, 72, seed=gen.REFRESH_SEED)
    fnet, plev = rc.heating_inputs(golden, n)
    for k in ("radlw", "radsw"):
        ref = golden[f"refresh/rad/{k}"][:, ::-1] * 86400.0      # K/day, surface first
        got = rx.heating_rate(fnet[k], plev, gen.RAD_GRAV, gen.RAD_CP)
        assert np.abs(got - ref).max() <= 1e-6, k
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
        out = torch.zeros((72, gen.REFRESH_NCOL), dtype=torch.float64, device="cuda")
        rx.heating_rate(dev(fnet[k]), dev(plev), gen.RAD_GRAV, gen.RAD_CP, device=True, out=out)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy().T, got)
