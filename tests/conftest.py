import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def rx():
    """The CUDA library through its host mirror; fails loudly when it is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import geosradiation_gridcomp_b200 as pkg
    pkg.init()
    return pkg.host
