"""The two SOR kernels written for the reference's own (L2-resident, latency-bound) decks, against the reference's outputs:

  variant 7 (pixelflow_b200/csrc/pf_sor_persistent.cu): the half-sweeps of a whole solve in one launch, hand-rolled
            grid barrier between them -- 2D cases and 3D air-condition;
  variant 8 (pixelflow_b200/csrc/pf_sor_tb2d.cu): temporally blocked -- four red-black iterations per launch on
            shared-memory tiles with a recomputed ring -- 2D cases.

Golden vectors: tests/golden/ref_translated.npz (the reference's own programs, machine-translated and run here).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_translated.npz"))


@pytest.mark.parametrize("variant", [7, 8])
@pytest.mark.parametrize("name", ["u2_even", "u2_odd", "b2_mixed", "d2_mixed", "a3_even", "a3_odd"])
def test_deck_kernels_equal_reference_outputs(gold, name, variant):
    from tests.test_gpu_z_ref_golden import _run, _same, _solver
    case = str(gold[f"{name}/case"])
    if variant == 8 and case.startswith("ibm3"):
        pytest.skip("variant 8 is the 2D kernel")
    st = json.loads(str(gold[f"{name}/settings"]))
    s = _solver(case, gold[f"{name}/dims"], st, gold[f"{name}/spacing"], sor_variant=variant)
    assert s.sor_variant == variant
    s.set_porosity(np.ascontiguousarray(gold[f"{name}/porosity"]))
    s.initial_conditions()
    errs, _ = _run(s, case, int(st["istep_max"]), st["radius"])
    u, v, w, p = s.download()
    _same(u, gold[f"{name}/u"], f"{name} u")
    _same(v, gold[f"{name}/v"], f"{name} v")
    if case.startswith("ibm3"):
        _same(w, gold[f"{name}/w"], f"{name} w")
    _same(p, gold[f"{name}/p"], f"{name} p")
    assert np.array_equal(errs, gold[f"{name}/perr"])
    s.close()


@pytest.mark.parametrize("variant", [7, 8])
@pytest.mark.parametrize("deck,golden_deck,case", [("cylinder", "cylinder", "ibm2_uniform"),
                                                   ("backstep", "backstep", "ibm2_backstep"),
                                                   ("room", "room_long", "ibm3_air_condition")])
def test_deck_kernels_on_the_shipped_decks(gold, deck, golden_deck, case, variant):
    if variant == 8 and case.startswith("ibm3"):
        pytest.skip("variant 8 is the 2D kernel")
    import hashlib
    from pixelflow_b200 import workloads as wl
    from pixelflow_b200.controldict import parse_controldict
    from tests.test_gpu_z_ref_golden import _run, _solver
    z = np.load(os.path.join(HERE, "golden", "decks", deck + ".npz"))
    cd = parse_controldict(str(z["controldict"]))
    steps = int(gold[f"deck_{golden_deck}/steps"])
    m, n, l = (int(x) for x in z["dims"])
    st = dict(xnue=cd.xnue, xlambda=cd.xlambda, density=cd.density, thickness=cd.thickness, nonslip=cd.nonslip,
              iter_max=cd.iter_max, relux_factor=cd.relux_factor, inlet_velocity=cd.inlet_velocity,
              outlet_pressure=cd.outlet_pressure, AoA=cd.AoA)
    s = _solver(case, (m, n, l), st, gold[f"deck_{golden_deck}/spacing"], sor_variant=variant)
    eps = np.maximum(z["porosity"] if case.startswith("ibm3") else z["porosity"][0], cd.threshold)
    s.set_porosity(wl.with_halos(eps, case))
    s.initial_conditions()
    errs, _ = _run(s, case, steps, cd.radius)
    u, v, w, p = s.download()
    sha = json.loads(str(gold[f"deck_{golden_deck}/sha"]))
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert h(u) == sha["u"] and h(v) == sha["v"] and h(p) == sha["p"]
    assert np.array_equal(errs, gold[f"deck_{golden_deck}/perr"])
    s.close()
