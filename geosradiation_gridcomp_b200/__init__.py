"""B200-native RRTMG LW + SW + McICA column radiation (drop-in for the GEOS RRTMG drivers).

The compute path is the CUDA library `librrtmgx.so` (csrc/, C ABI in include/rrtmgx.h);
`host` mirrors the reference driver interfaces on top of it.
"""
from . import host, synthetic  # noqa: F401
from .host import (RrtmgxError, finalize, heating_rate, init, initialize_cloud_subcol_gen, rrtmg_lw,  # noqa: F401
                   rrtmg_lw_ini, rrtmg_sw, rrtmg_sw_ini, set_inhomogeneity)
