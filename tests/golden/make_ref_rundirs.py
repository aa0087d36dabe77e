"""Golden RUN DIRECTORIES of the reference: tests/golden/ref_rundirs.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_ref_rundirs.py

Each case is a small project directory run by the translated reference program in its "gf" flavour
(oracle/build_ref.py): lib/output.f90 translated too, every WRITE executed by libgfortran.so.5 — so the directory
afterwards holds exactly what a gfortran build of the reference leaves behind: stdout.log (unit *), etc/*.dat,
<output_folder>/*.vtk.  Stored: the deck (settings, raw porosity) and every file's bytes.  The GPU test
tests/test_gpu_zz_driver_rundirs.py runs the drop-in driver on the same decks and compares byte for byte.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_translated as rt  # noqa: E402
from tests.test_ref_output_files import _deck  # noqa: E402

CASES = {
    "u3": ("ibm3_uniform", (12, 10, 8), dict(istep_out=2, AoA=4.0)),
    "a3": ("ibm3_air_condition", (9, 8, 7), dict(istep_out=1)),
    "u2": ("ibm2_uniform", (20, 12, 1), dict(istep_out=2)),
    "b2": ("ibm2_backstep", (19, 11, 1), dict(istep_out=3)),
    # the force log's serial sums: compared byte for byte only where the sums are serial too (the Fortran driver on the
    # ABI test double); the GPU's two-stage sums differ in the last bits (tests/test_gpu_decks.py holds them to 1e-12)
    "d2": ("ibm2_drag", (20, 11, 1), dict(istep_out=100, radius=0.05)),
}


def main():
    out = {}
    for name, (case, dims, extra) in CASES.items():
        eps, st = _deck(case, dims, extra)
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "etc"))
            os.makedirs(os.path.join(d, st["output_folder"]))
            rt.write_deck(d, eps, **st)
            R = rt.RefProgram(case, "gf", "s")
            perr = R.run(d)
            files = {}
            for root, _, fs in os.walk(d):
                for f in fs:
                    rel = os.path.relpath(os.path.join(root, f), d)
                    if rel.startswith(("config", "data")):
                        continue
                    files[rel] = open(os.path.join(root, f), "rb").read()
        out[f"{name}/case"] = np.array(case)
        out[f"{name}/dims"] = np.array(dims)
        out[f"{name}/settings"] = np.array(json.dumps(st))
        out[f"{name}/porosity_in"] = eps
        out[f"{name}/perr"] = perr
        out[f"{name}/files"] = np.array(json.dumps(sorted(files)))
        for rel, data in files.items():
            out[f"{name}/file/{rel}"] = np.frombuffer(data, dtype=np.uint8)
        print(name, case, dims, {k: len(v) for k, v in files.items()})
    path = os.path.join(HERE, "ref_rundirs.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
