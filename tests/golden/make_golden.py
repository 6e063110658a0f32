"""Regenerates tests/golden/rrtmg_golden_L72.npz: inputs are the deterministic synthetic columns
(seed, ncol below); outputs are what the CPU oracle (oracle/, the fp64 restatement of the reference
Fortran) returns for them.  The reference itself cannot be run here (no Fortran compiler, see
DESIGN.md), so these vectors pin the ORACLE against drift and give the GPU tests a fixture that
does not need the oracle to be rebuilt; they are not reference output.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SEED, NCOL, NLAY = 20260118, 48, 72
LW_KEYS = ("uflx", "dflx", "uflxc", "dflxc", "duflx_dTs", "duflxc_dTs", "olrb", "dolrb_dTs", "clearCounts")
SW_KEYS = ("swuflx", "swdflx", "swuflxc", "swdflxc", "nirr", "nirf", "parr", "parf", "uvrr", "uvrf", "fswband",
           "cotdtp", "cotdhp", "cotdmp", "cotdlp", "cotntp", "cotnhp", "cotnmp", "cotnlp", "clearCounts")


def main():
    from geosradiation_gridcomp_b200.synthetic import make_columns
    from oracle import binding as oracle
    s = make_columns(NCOL, NLAY, seed=SEED)
    lw = oracle.rrtmg_lw(s, taps=("jp", "jt", "jt1", "laytrop"))
    sw = oracle.rrtmg_sw(s, taps=("jp", "jt", "jt1", "laytrop"))
    assert lw["rc"] == 0 and sw["rc"] == 0
    out = {"seed": SEED, "ncol": NCOL, "nlay": NLAY}
    for k in LW_KEYS + ("jp", "jt", "jt1", "laytrop"):
        out["lw_" + k] = lw[k]
    for k in SW_KEYS + ("jp", "jt", "jt1", "laytrop"):
        out["sw_" + k] = sw[k]
    np.savez_compressed(os.path.join(HERE, "rrtmg_golden_L72.npz"), **out)
    print("wrote", os.path.join(HERE, "rrtmg_golden_L72.npz"))


if __name__ == "__main__":
    main()
