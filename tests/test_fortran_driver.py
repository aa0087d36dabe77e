"""The product's FORTRAN driver, executed — on CPU.

pixelflow_b200/fortran/ibm{2_uniform,2_backstep,2_drag,3_uniform,3_air_condition}_gpu.f90 (`use pixelflow_gpu`) are the
reference's five `program main`s with the body of the time loop replaced by calls into the C ABI through the iso_c_binding module
pixelflow_gpu_mod.f90.  No Fortran compiler exists in the image, so it is translated to C like the reference itself
(oracle/f90toc.py; the module's bind(C) types and interfaces are taken from numpy's Fortran parser,
oracle/f90_cmodule.py), together with the reference's own lib/global.f90, lib/grid.f90 and lib/output.f90, and its
WRITE statements are executed by libgfortran.  In the build container there is no GPU either, so here the ABI calls
are answered by a TEST DOUBLE of the C ABI backed by the oracle (oracle/abi_double.c, compiled against the real
header; it is not the product and cannot be loaded by it).  What this pins: the driver's statement order, the
marshalling of the namelist values into `pf_config` through the bind(C) type (leading dimensions md+1 / nd+1 of the
static arrays; the type's layout itself is held to the header by tests/test_fortran_binding.py), and that driver + reference output routines leave the reference's run directory
byte for byte.  The same translated driver linked against the product library runs in
tests/test_gpu_zzz_fortran_driver.py.
"""
import json
import os

import numpy as np
import pytest

from oracle import build_ref, gfortran_rt
from oracle import ref_translated as rt

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = [
    pytest.mark.skipif(not build_ref.available(), reason="the Fortran driver is translated together with /root/reference's lib/*.f90"),
    pytest.mark.skipif(gfortran_rt.find_libgfortran() is None, reason="libgfortran.so.5 not found"),
]


RUNDIRS = {"u3": "ibm3_uniform", "a3": "ibm3_air_condition", "u2": "ibm2_uniform", "b2": "ibm2_backstep",
           "d2": "ibm2_drag"}


@pytest.mark.parametrize("name", list(RUNDIRS))
def test_fortran_driver_leaves_the_reference_run_directory(name, tmp_path):
    """all five drivers (the reference's five programs): log, etc/*.dat and every VTK file, byte for byte.  ibm2_drag
    prints its force log from pf_force_log_2d; on the test double the sums are the oracle's serial ones, so those
    lines are identical too (on the GPU they agree to rounding, tests/test_gpu_decks.py)"""
    from tests.test_gpu_zz_driver_rundirs import _write_deck
    case = RUNDIRS[name]
    driver = build_ref.build_fortran_driver("double", case)
    gold = np.load(os.path.join(HERE, "golden", "ref_rundirs.npz"))
    st = json.loads(str(gold[f"{name}/settings"]))
    _write_deck(str(tmp_path), gold[f"{name}/porosity_in"], st)
    (tmp_path / "etc").mkdir()                      # `call system('mkdir -p ...')` is a stub
    (tmp_path / st["output_folder"]).mkdir()
    R = rt.RefProgram("fortran_driver", "gf", lib=driver)
    perr = R.run(str(tmp_path))
    assert np.array_equal(perr, gold[f"{name}/perr"])
    assert R.stub_count("get_now_time") == 4 and R.stub_count("system") == 2
    for rel in json.loads(str(gold[f"{name}/files"])):
        assert (tmp_path / rel).read_bytes() == bytes(gold[f"{name}/file/{rel}"]), rel


@pytest.mark.parametrize("name", ["u3_even", "u3_odd", "u3_mixed", "u3_pout", "a3_even", "a3_odd", "u2_even", "u2_odd",
                                  "b2_mixed", "d2_mixed"])
def test_fortran_driver_fields_equal_the_reference(name, tmp_path):
    """the fields the driver holds after its last pf_download == what the reference program leaves in its arrays
    (golden vectors of tests/golden/ref_translated.npz): even / odd sizes, nonslip off, xlambda, outlet pressure"""
    gold = np.load(os.path.join(HERE, "golden", "ref_translated.npz"))
    case = str(gold[f"{name}/case"])
    driver = build_ref.build_fortran_driver("double", case)
    st = dict(rt.DEFAULTS)
    st.update(json.loads(str(gold[f"{name}/settings"])))
    rt.write_deck(str(tmp_path), gold[f"{name}/porosity_in"], **st)
    (tmp_path / "etc").mkdir()
    (tmp_path / st["output_folder"]).mkdir()
    R = rt.RefProgram(case, "gf", lib=driver)
    perr = R.run(str(tmp_path))
    assert np.array_equal(perr, gold[f"{name}/perr"])
    for k in (("u", "v", "w", "p", "porosity") if case.startswith("ibm3") else ("u", "v", "p", "porosity")):
        assert np.array_equal(R.array(k), gold[f"{name}/{k}"]), k
