"""The drop-in driver, run for real on the GPU, against the reference's own run directories — byte for byte.

tests/golden/ref_rundirs.npz (made by tests/golden/make_ref_rundirs.py) holds, for four small decks, every file the
reference leaves in its project directory — stdout (unit *), etc/grid.dat, etc/solution_uvp.dat, etc/divergent.dat,
etc/surface_profile.dat, <output_folder>/output_NNNNN.vtk, output_paraview.vtk — produced by the reference's own
programs and output routines (machine-translated, WRITE statements executed by libgfortran).  Here the C++ twin
driver runs the same decks under the reference's executable names: porosity CSV parsed on the GPU, the time steps
on the GPU, the VTK bodies formatted on the GPU.  Everything must be identical except the wall-clock TIME stamps.

(File name sorts last on purpose: it is the end-to-end check of everything the other GPU tests check in parts.)
"""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

EXE = {"ibm3_uniform": "ibm3_uniform_omp", "ibm3_air_condition": "ibm3_air_condition_omp",
       "ibm2_uniform": "ibm2_uniform_omp", "ibm2_backstep": "ibm2_backstep_omp"}


def _no_time(text):
    return [ln for ln in text.splitlines() if not ln.startswith(" # --- TIME:")]


def _write_deck(d, eps, st):
    """config/controlDict.txt + the porosity CSV (template/data/.porosity format), product-side code only"""
    groups = {
        "physical": ("xnue", "xlambda", "density", "width", "height", "depth", "time", "inlet_velocity",
                     "outlet_pressure", "AoA"),
        "file_control": ("istep_out",), "grid_control": ("istep_max",),
        "porosity_control": ("thickness", "threshold", "radius", "center_x", "center_y", "center_z"),
        "calculation_method": ("nonslip",), "directory_control": ("output_folder", "csv_file"),
        "solver_control": ("iter_max", "relux_factor"),
    }
    lines = []
    for g, keys in groups.items():
        lines.append(f"&{g}")
        for k in keys:
            v = st[k]
            if isinstance(v, bool):
                lines.append(f"{k} = {'.true.' if v else '.false.'}")
            elif isinstance(v, str):
                lines.append(f'{k} = "{v}"')
            elif k in ("istep_out", "istep_max", "iter_max"):
                lines.append(f"{k} = {int(v)}")
            else:
                lines.append(f"{k} = {float(v)!r}")
        lines.append("/")
    os.makedirs(os.path.join(d, "config"))
    os.makedirs(os.path.dirname(os.path.join(d, st["csv_file"])), exist_ok=True)
    with open(os.path.join(d, "config", "controlDict.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    e = eps if eps.ndim == 3 else eps[None]
    l, n, m = e.shape
    with open(os.path.join(d, st["csv_file"]), "w") as f:
        f.write(f"{m},{n},{l}\n")
        for k in range(l):
            for j in range(n):
                f.write("".join(f"{i + 1}, {j + 1}, {k + 1}, {float(e[k, j, i])!r}\n" for i in range(m)))


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# gpus > 1: the same driver, `--gpus N` -- one forked process per GPU (pf_ranks_launch), z-slabs of the grid, every
# rank writing its own planes of the VTK snapshots; u3 has 8 planes (4 + 4: the fused SOR kernel with peer stores;
# 2 x 4: half-sweeps with overlapped exchanges), a3 has 7 (4 + 3, open chain).  Skipped below N GPUs.
@pytest.mark.parametrize("name,gpus", [("u3", 1), ("a3", 1), ("u2", 1), ("b2", 1), ("u3", 2), ("a3", 2), ("u3", 4)])
def test_driver_run_directory_equals_the_reference(name, gpus, tmp_path):
    from pixelflow_b200 import build
    build.build_drivers()
    if gpus > 1 and _gpu_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    case = str(gold[f"{name}/case"])
    st = json.loads(str(gold[f"{name}/settings"]))
    _write_deck(str(tmp_path), gold[f"{name}/porosity_in"], st)
    exe = os.path.join(ROOT, "pixelflow_b200", "driver", "bin", EXE[case])
    r = subprocess.run([exe] + (["--gpus", str(gpus)] if gpus > 1 else []), cwd=tmp_path, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    files = json.loads(str(gold[f"{name}/files"]))
    ref_log = bytes(gold[f"{name}/file/stdout.log"]).decode()
    assert _no_time(r.stdout) == _no_time(ref_log)
    produced = sorted(os.path.relpath(os.path.join(root, f), tmp_path) for root, _, fs in os.walk(tmp_path) for f in fs
                      if not root.endswith(("config", "data")))
    assert produced == sorted(f for f in files if f != "stdout.log")
    for rel in files:
        if rel == "stdout.log":
            continue
        a, b = (tmp_path / rel).read_bytes(), bytes(gold[f"{name}/file/{rel}"])
        if a != b:
            la, lb = a.decode().splitlines(), b.decode().splitlines()
            first = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), min(len(la), len(lb)))
            raise AssertionError(f"{rel}: line {first + 1} of {len(la)}/{len(lb)}:\n driver    "
                                 f"{la[first][:160] if first < len(la) else '<eof>'}\n reference "
                                 f"{lb[first][:160] if first < len(lb) else '<eof>'}")
