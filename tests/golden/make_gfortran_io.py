#!/usr/bin/env python
"""Golden vectors of the reference's formatted / list-directed I/O, produced by libgfortran.so.5 ITSELF.

The reference's Fortran cannot be compiled here, but its runtime library is in the image (numpy.libs / scipy.libs);
oracle/gfortran_rt.py issues the calls a gfortran build generates for
    write(65,"(3(f16.4,1x))") ...        (lib/output.f90:421-537, :968-1088)
    write(*,*) '...', value, ...         (lib/global.f90:66-90, lib/grid.f90:309-321, the time loop, the force log)
    read(52,*) x, y, z, poro_val         (lib/grid.f90:288, :42)
This script stores what the runtime answers in tests/golden/gfortran_io.npz, so that the tests still have the
reference's own outputs on a machine without the library.

    python tests/golden/make_gfortran_io.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gfortran_rt as g  # noqa: E402
from tests.gfortran_cases import F16_VALUES, LIST_RECORDS, READ_RECORDS  # noqa: E402


def main():
    assert g.available(), "no libgfortran.so.5 found"
    out = {"libgfortran": os.path.basename(g.find_libgfortran())}
    vals = F16_VALUES()
    one = g.formatted_write("(3(f16.4,1x))", vals, 1).decode().split("\n")[:-1]
    assert len(one) == len(vals)
    out["f16_4"] = {"values": [float(v).hex() for v in vals], "records": one}
    three = vals[:len(vals) // 3 * 3]
    out["f16_4_three"] = g.formatted_write("(3(f16.4,1x))", three, 3).decode().split("\n")[:-1]
    out["list_write"] = [g.list_write(*items).decode() for items in LIST_RECORDS()]
    rd = []
    for line in READ_RECORDS():
        ios, x, y, z, v = g.list_read_record(line)
        rd.append({"line": line, "iostat": ios, "xyz": [x, y, z], "value": float(v).hex() if ios == 0 else None})
    out["list_read"] = rd
    import numpy as np
    small = {k: out[k] for k in ("libgfortran", "list_write", "list_read")}
    np.savez_compressed(os.path.join(HERE, "gfortran_io.npz"),
                        f16_values=np.array(vals, dtype=np.float64),
                        f16_records=np.frombuffer("\n".join(one).encode(), dtype=np.uint8),
                        f16_three=np.frombuffer("\n".join(out["f16_4_three"]).encode(), dtype=np.uint8),
                        meta=np.frombuffer(json.dumps(small).encode(), dtype=np.uint8))
    print({k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
