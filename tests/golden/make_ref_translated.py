"""Golden vectors from the REFERENCE ITSELF: tests/golden/ref_translated.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_ref_translated.py

The reference is Fortran and there is no Fortran compiler here, so its five programs are turned into C by the
mechanical translator oracle/f90toc.py (oracle/build_ref.py -> oracle/_ref/*.so) and RUN: every case below is a
project directory (config/controlDict.txt + porosity CSV) handed to the translated `program main`, exactly as a
user runs the reference.  Stored per case: the inputs (settings, raw porosity) and what the program left in its
arrays after the last step (u, v, [w,] p, porosity incl. halos), dx/dy/dz/dt as it computed them, the 'p error' it
logged per step and, for ibm2_drag, the Fp/Fv/F/Cd/Cl lines of output_force_log_2d.

The three shipped decks (test/*.zip) are run UNMODIFIED for their first steps (step limit of the harness; the
files are too big to commit, so fields are stored as SHA-256 of their bytes plus the logged p errors).
"""
import hashlib
import json
import os
import re
import shutil
import sys
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_translated as rt  # noqa: E402

REF_TEST = "/root/reference/test"

# name -> (case, dims, settings)
SMALL = {
    "u3_even": ("ibm3_uniform", (12, 10, 8), dict()),
    "u3_odd": ("ibm3_uniform", (11, 9, 7), dict(AoA=12.0)),
    "u3_mixed": ("ibm3_uniform", (12, 9, 8), dict(nonslip=False, xlambda=2e-4, threshold=0.05)),
    "u3_pout": ("ibm3_uniform", (10, 10, 7), dict(outlet_pressure=0.3, density=1.2, thickness=2.0)),
    "a3_even": ("ibm3_air_condition", (10, 8, 6), dict()),
    "a3_odd": ("ibm3_air_condition", (9, 7, 11), dict(outlet_pressure=0.1, threshold=0.05)),
    "u2_even": ("ibm2_uniform", (14, 10, 1), dict(AoA=5.0)),
    "u2_odd": ("ibm2_uniform", (13, 9, 1), dict(nonslip=False, xlambda=1e-4, threshold=0.05)),
    "b2_mixed": ("ibm2_backstep", (14, 9, 1), dict()),
    "d2_mixed": ("ibm2_drag", (13, 10, 1), dict(AoA=3.0, radius=0.07)),
}
STEPS, ITER_MAX = 3, 8


def small_settings(dims, extra):
    m, n, l = dims
    s = dict(xnue=1e-3, xlambda=0.0, density=1.0, width=0.1 * (m - 1) / 16, height=0.1 * (n - 1) / 16,
             depth=0.1 * max(l - 1, 1) / 16, time=0.0005 * STEPS, istep_max=STEPS, iter_max=ITER_MAX, relux_factor=1.7,
             inlet_velocity=0.8, outlet_pressure=0.0, AoA=0.0, thickness=1.5, threshold=1e-6, nonslip=True, radius=0.1)
    s.update(extra)
    return s


def small_porosity(name, dims):
    """a solid blob (tanh profile, >= 0.02) in a fluid box plus a little noise; the faces hold both >= 0.9 and < 0.9
    values (the air-condition walls switch on that).  The `max(poro, threshold)` clamp of lib/grid.f90 is exercised by
    the cases that raise `threshold` to 0.05 and by the cylinder deck (values down to 1.4e-9)."""
    m, n, l = dims
    rng = np.random.default_rng(abs(hash_name(name)) % (2 ** 32))
    k, j, i = np.meshgrid(np.arange(l), np.arange(n), np.arange(m), indexing="ij")
    r = np.sqrt(((i - 0.4 * m) / (0.25 * m)) ** 2 + ((j - 0.5 * n) / (0.3 * n)) ** 2 + (((k - 0.5 * l) / (0.3 * l)) ** 2 if l > 1 else 0))
    e = 0.5 * np.tanh((r - 1.0) * 2.0) + 0.5
    e = np.clip(e + 0.06 * (rng.random((l, n, m)) - 0.5), 0.02, 1.0)
    return e if l > 1 else e[0]


def hash_name(name):
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")


def force_lines(log: str):
    vals = []
    num = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|[-+]?(?:inf|nan)"
    for line in log.splitlines():
        t = line.strip()
        if t.startswith(("Fp =", "Fv =", "F  =", "Cd =")):
            vals += [float(x) for x in re.findall(num, t.split("=", 1)[1].replace("Cl =", " "))]
    return np.array(vals).reshape(-1, 8)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_small(out):
    for name, (case, dims, extra) in SMALL.items():
        st = small_settings(dims, extra)
        eps = small_porosity(name, dims)
        R = rt.RefProgram(case, "serial", "s")
        with tempfile.TemporaryDirectory() as d:
            rt.write_deck(d, eps, **st)
            perr = R.run(d)
        f = R.fields()
        out[f"{name}/case"] = np.array(case)
        out[f"{name}/dims"] = np.array(dims)
        out[f"{name}/settings"] = np.array(json.dumps(st))
        out[f"{name}/porosity_in"] = eps
        for k, a in f.items():
            out[f"{name}/{k}"] = a
        out[f"{name}/perr"] = perr
        out[f"{name}/spacing"] = np.array([R.scalar(k) for k in (("dx", "dy", "dz", "dt") if R.d3 else ("dx", "dy", "dt"))])
        if case == "ibm2_drag":
            out[f"{name}/force"] = force_lines(R.log())
        print(name, case, dims, "p error", perr)


DECKS = {
    "room": ("room.zip", "ibm3_air_condition", "s", 3),
    "cylinder": ("cylinder-2d.zip", "ibm2_uniform", "b", 3),
    "cylinder_drag": ("cylinder-2d.zip", "ibm2_drag", "b", 3),
    "backstep": ("backstep.zip", "ibm2_backstep", "b", 3),
    # a longer run of the unmodified room deck: 100 steps x 100 SOR iterations
    "room_long": ("room.zip", "ibm3_air_condition", "s", 100),
}


def run_decks(out):
    from pixelflow_b200.controldict import parse_controldict
    for name, (zf, case, size, nsteps) in DECKS.items():
        with tempfile.TemporaryDirectory() as d:
            zipfile.ZipFile(os.path.join(REF_TEST, zf)).extractall(d)
            root = os.path.join(d, zf[:-4])
            cd = parse_controldict(open(os.path.join(root, "config", "controlDict.txt")).read())
            if not os.path.exists(os.path.join(root, cd.csv_file)):
                # backstep's controlDict names data/porosity_1.5_300.csv, the zip ships data/backstep.csv
                csvs = [f for f in os.listdir(os.path.join(root, "data")) if f.endswith(".csv")]
                shutil.copy(os.path.join(root, "data", csvs[0]), os.path.join(root, cd.csv_file))
            R = rt.RefProgram(case, "serial", size)
            perr = R.run(root, step_limit=nsteps)
        f = R.fields()
        out[f"deck_{name}/case"] = np.array(case)
        out[f"deck_{name}/steps"] = np.array(nsteps)
        out[f"deck_{name}/perr"] = perr
        out[f"deck_{name}/spacing"] = np.array([R.scalar(k) for k in (("dx", "dy", "dz", "dt") if R.d3 else ("dx", "dy", "dt"))])
        out[f"deck_{name}/sha"] = np.array(json.dumps({k: sha(a) for k, a in f.items()}))
        # a thin sample of p for a readable failure message (every 37th value)
        out[f"deck_{name}/p_sample"] = f["p"].ravel()[::37].copy()
        if case == "ibm2_drag":
            out[f"deck_{name}/force"] = force_lines(R.log())
        print("deck", name, case, "p error", perr)


def main():
    out = {}
    run_small(out)
    run_decks(out)
    path = os.path.join(HERE, "ref_translated.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
