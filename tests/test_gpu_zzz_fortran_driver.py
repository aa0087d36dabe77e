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


def test_fortran_driver_on_the_gpu_leaves_the_reference_run_directory(tmp_path):
    from oracle import build_ref, gfortran_rt
    from oracle import ref_translated as rt
    from tests.test_gpu_zz_driver_rundirs import _write_deck
    lib = os.path.join(build_ref.OUT, "fdriver_ibm3_uniform_gpu.so")
    if not os.path.exists(lib) and not build_ref.available():
        pytest.skip("oracle/_ref/fdriver_ibm3_uniform_gpu.so was not prebuilt (it needs /root/reference to build)")
    if gfortran_rt.find_libgfortran() is None:
        pytest.skip("libgfortran.so.5 not found")
    lib = build_ref.build_fortran_driver("gpu")
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    st = json.loads(str(gold["u3/settings"]))
    _write_deck(str(tmp_path), gold["u3/porosity_in"], st)
    (tmp_path / "etc").mkdir()
    (tmp_path / st["output_folder"]).mkdir()
    R = rt.RefProgram("fortran_driver", "gf", lib=lib)
    perr = R.run(str(tmp_path))
    assert np.array_equal(perr, gold["u3/perr"])
    for rel in json.loads(str(gold["u3/files"])):
        assert (tmp_path / rel).read_bytes() == bytes(gold[f"u3/file/{rel}"]), rel
