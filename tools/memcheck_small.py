"""Small end-to-end pass over every entry point for compute-sanitizer (ragged column counts, both layer counts)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import geosradiation_gridcomp_b200 as pkg
from geosradiation_gridcomp_b200 import host as rx
from geosradiation_gridcomp_b200.synthetic import make_columns, make_native_state

pkg.init()
for ncol, nlay in ((203, 72), (37, 181)):
    s = make_columns(ncol, nlay, seed=77)
    lw = rx.run_lw(s)
    sw = rx.run_sw(s)
    rx.run_lw(s, reuse_clouds=True)
    rx.run_sw(s, iaer=0, reuse_clouds=True)
    rv = rx.run_sw(s, radval=True)                                  # the SOLAR_RADVAL build
    rx.run_sw(s, radval=True, iaer=0, reuse_clouds=True)
    rx.run_sw(s, reuse_clouds=True)                                 # default build after a RADVAL call
    assert np.isfinite(rv["radval"]).all()
    n = make_native_state(ncol, nlay, seed=78)
    f = rx.irrad_refresh(n)
    rx.solar_refresh(n)
    rx.irrad_update(f, n["ts"], np.asfortranarray(n["ts"] + 1.0))
    rx.irrad_prepare(n); rx.solar_prepare(n)
    print(ncol, nlay, float(lw["uflx"].sum()), float(sw["swdflx"].sum()), float(f["flxu"].sum()))
print("memcheck pass done")
