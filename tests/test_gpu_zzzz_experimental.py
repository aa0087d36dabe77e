"""EXPERIMENTAL kernels — run only with PF_TEST_EXPERIMENTAL=1.

SOR variant 7 (pixelflow_b200/csrc/pf_sor_persistent.cu: the half-sweeps of a whole solve in one cooperative launch)
was written after the round's GPU budget was spent: it compiles for sm_100a, its SASS shows coherent loads for the
pressure and grid.sync()'s L1 invalidation, but it has not run on a GPU yet.  It is opt-in (`sor_variant=7`, never
auto-selected), and these parity tests are the first thing to run when a GPU is available:

    PF_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zzzz_experimental.py -q
    python tools/bench_decks.py --sor-variant 7      # against: python tools/bench_decks.py
"""
import json
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PF_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernel, not yet verified on a GPU: set PF_TEST_EXPERIMENTAL=1")]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_translated.npz"))


@pytest.mark.parametrize("name", ["u2_even", "u2_odd", "b2_mixed", "d2_mixed", "a3_even", "a3_odd"])
def test_persistent_half_sweeps_equal_reference_outputs(gold, name):
    from tests.test_gpu_z_ref_golden import _run, _same, _solver
    case = str(gold[f"{name}/case"])
    st = json.loads(str(gold[f"{name}/settings"]))
    s = _solver(case, gold[f"{name}/dims"], st, gold[f"{name}/spacing"], sor_variant=7)
    assert s.sor_variant == 7
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


@pytest.mark.parametrize("deck,golden_deck,case", [("cylinder", "cylinder", "ibm2_uniform"),
                                                   ("backstep", "backstep", "ibm2_backstep"),
                                                   ("room", "room_long", "ibm3_air_condition")])
def test_persistent_half_sweeps_on_the_shipped_decks(gold, deck, golden_deck, case):
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
    s = _solver(case, (m, n, l), st, gold[f"deck_{golden_deck}/spacing"], sor_variant=7)
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
