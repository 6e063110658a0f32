"""Reader for the extracted reference data blob (see tools/extract_tables.py for the format).

Used by tests and tooling; the CUDA library and the C oracle each carry their own C reader.
"""
import os
import struct

import numpy as np

DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "rrtmg_tables.bin")


def load_tables(path=DEFAULT_BLOB):
    """Return {name: ndarray (Fortran order)} for every table in the blob."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"RRTMGTB1":
        raise ValueError("bad table blob magic")
    (n,) = struct.unpack_from("<i", buf, 8)
    out = {}
    pos = 12
    for _ in range(n):
        name = buf[pos:pos + 48].split(b"\0")[0].decode()
        dt, nd, *rest = struct.unpack_from("<ii6iqq", buf, pos + 48)
        dims, off, nb = rest[:6], rest[6], rest[7]
        pos += 48 + 4 + 4 + 24 + 16
        dtype = np.int32 if dt == 1 else np.float64
        a = np.frombuffer(buf, dtype=dtype, count=nb // np.dtype(dtype).itemsize, offset=off)
        out[name] = a.reshape(dims[:nd], order="F")
    return out
