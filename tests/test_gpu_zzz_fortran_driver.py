"""The product's FORTRAN driver, executed on the GPU.

pixelflow_b200/fortran/ibm3_uniform_gpu.f90 + pixelflow_gpu_mod.f90, translated to C like the reference (no Fortran
compiler in the image; oracle/f90toc.py, oracle/f90_cmodule.py — see tests/test_fortran_driver.py, which runs the
same translation against a CPU test double of the ABI) and linked against the PRODUCT library
pixelflow_b200/libpixelflow_gpu.so (oracle/build_ref.py:build_fortran_driver("gpu"), prebuilt where /root/reference
exists).  Every pf_* call of the Fortran text reaches the CUDA path; the reference's own grid and output routines do
the rest; the run directory must equal the reference's byte for byte (tests/golden/ref_rundirs.npz).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


RUNDIRS = {"u3": "ibm3_uniform", "a3": "ibm3_air_condition", "u2": "ibm2_uniform", "b2": "ibm2_backstep"}


@pytest.mark.parametrize("name", list(RUNDIRS))
def test_fortran_driver_on_the_gpu_leaves_the_reference_run_directory(name, tmp_path):
    from oracle import build_ref, gfortran_rt
    from oracle import ref_translated as rt
    from tests.test_gpu_zz_driver_rundirs import _write_deck
    case = RUNDIRS[name]
    lib = os.path.join(build_ref.OUT, f"fdriver_{case}_gpu.so")
    if not os.path.exists(lib) and not build_ref.available():
        pytest.skip(f"oracle/_ref/fdriver_{case}_gpu.so was not prebuilt (it needs /root/reference to build)")
    if gfortran_rt.find_libgfortran() is None:
        pytest.skip("libgfortran.so.5 not found")
    lib = build_ref.build_fortran_driver("gpu", case)
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    st = json.loads(str(gold[f"{name}/settings"]))
    _write_deck(str(tmp_path), gold[f"{name}/porosity_in"], st)
    (tmp_path / "etc").mkdir()
    (tmp_path / st["output_folder"]).mkdir()
    R = rt.RefProgram(case, "gf", lib=lib)
    perr = R.run(str(tmp_path))
    assert np.array_equal(perr, gold[f"{name}/perr"])
    for rel in json.loads(str(gold[f"{name}/files"])):
        assert (tmp_path / rel).read_bytes() == bytes(gold[f"{name}/file/{rel}"]), rel


def test_fortran_drag_driver_on_the_gpu(tmp_path):
    """ibm2_drag: everything byte for byte except the eight force-log numbers per step, which are two-stage sums on
    the GPU (equal to the reference's serial sums to rounding)"""
    import re
    from oracle import build_ref, gfortran_rt
    from oracle import ref_translated as rt
    from tests.test_gpu_zz_driver_rundirs import _write_deck
    lib = os.path.join(build_ref.OUT, "fdriver_ibm2_drag_gpu.so")
    if not os.path.exists(lib) and not build_ref.available():
        pytest.skip("oracle/_ref/fdriver_ibm2_drag_gpu.so was not prebuilt")
    if gfortran_rt.find_libgfortran() is None:
        pytest.skip("libgfortran.so.5 not found")
    lib = build_ref.build_fortran_driver("gpu", "ibm2_drag")
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    st = json.loads(str(gold["d2/settings"]))
    _write_deck(str(tmp_path), gold["d2/porosity_in"], st)
    (tmp_path / "etc").mkdir()
    (tmp_path / st["output_folder"]).mkdir()
    R = rt.RefProgram("ibm2_drag", "gf", lib=lib)
    perr = R.run(str(tmp_path))
    assert np.array_equal(perr, gold["d2/perr"])
    force = ("Fp =", "Fv =", "F  =", "Cd =")
    num = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"
    for rel in json.loads(str(gold["d2/files"])):
        a, b = (tmp_path / rel).read_bytes(), bytes(gold[f"d2/file/{rel}"])
        if rel != "stdout.log":
            assert a == b, rel
            continue
        la, lb = a.decode().splitlines(), b.decode().splitlines()
        assert len(la) == len(lb)
        for x, y in zip(la, lb):
            if x.strip().startswith(force):
                vx = np.array([float(t) for t in re.findall(num, x.split("=", 1)[1].replace("Cl =", " "))])
                vy = np.array([float(t) for t in re.findall(num, y.split("=", 1)[1].replace("Cl =", " "))])
                assert np.allclose(vx, vy, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(vy).max())), (x, y)
            else:
                assert x == y


@pytest.mark.parametrize("name,gpus", [("u3", 2), ("a3", 2)])
def test_fortran_driver_on_several_gpus(name, gpus, tmp_path):
    """PIXELFLOW_GPUS=N: the Fortran text forks into N ranks (pf_ranks_launch), gathers the fields on rank 0
    (pf_gather) and lets the reference's own output routines write the files there.  Run in a fresh process (the
    program forks).  Every file equals the reference's; the log is compared as a set of lines, because in this test
    harness unit * is a file all ranks share (a compiled Fortran program writes it to stdout, which
    pf_ranks_launch silences in the children)."""
    import subprocess
    import sys
    import torch
    from oracle import build_ref, gfortran_rt
    from tests.test_gpu_zz_driver_rundirs import _write_deck
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    case = RUNDIRS[name]
    lib = os.path.join(build_ref.OUT, f"fdriver_{case}_gpu.so")
    if not os.path.exists(lib) and not build_ref.available():
        pytest.skip(f"oracle/_ref/fdriver_{case}_gpu.so was not prebuilt (it needs /root/reference to build)")
    if gfortran_rt.find_libgfortran() is None:
        pytest.skip("libgfortran.so.5 not found")
    lib = build_ref.build_fortran_driver("gpu", case)
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    st = json.loads(str(gold[f"{name}/settings"]))
    _write_deck(str(tmp_path), gold[f"{name}/porosity_in"], st)
    (tmp_path / "etc").mkdir()
    (tmp_path / st["output_folder"]).mkdir()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import ref_translated as rt; "
            "R = rt.RefProgram(%r, 'gf', lib=%r); perr = R.run(%r); np.save(%r, perr)"
            % (os.path.dirname(HERE), case, lib, str(tmp_path), str(tmp_path / "perr.npy")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PIXELFLOW_GPUS=str(gpus)), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert np.array_equal(np.load(tmp_path / "perr.npy"), gold[f"{name}/perr"])
    for rel in json.loads(str(gold[f"{name}/files"])):
        a, b = (tmp_path / rel).read_bytes(), bytes(gold[f"{name}/file/{rel}"])
        if rel == "stdout.log":
            assert set(a.decode().splitlines()) == set(b.decode().splitlines())
        else:
            assert a == b, rel
