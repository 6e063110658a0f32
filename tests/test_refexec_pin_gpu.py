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
