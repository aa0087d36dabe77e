"""Longer runs: parity must hold step after step (no drift), not just after a handful of steps."""
import numpy as np
import pytest

from tests.test_gpu_decks import _same, deck_kwargs, load_deck

pytestmark = pytest.mark.gpu


def test_room_deck_30_steps(oracle):
    """configs[2] for 30 time steps (3,000 SOR iterations): still bit-identical, p error equal every step"""
    from pixelflow_b200 import Solver
    cd, (m, n, l), eps = load_deck("room")
    kw = deck_kwargs(cd, (m, n, l), True)
    P = oracle.make_params(m=m, n=n, l=l, wall=(1, 0, 0, 0, 2, 0), **kw)
    oc = oracle.Oracle3D(P, True, eps)
    oc.initialise()
    s = Solver("ibm3_air_condition", m, n, l, wall=(1, 0, 0, 0, 2, 0), **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    err_o, err_g = oc.step(30), s.step(30)
    assert np.array_equal(err_o, err_g)
    u, v, w, p = s.download()
    for nm, a, b in (("u", u, oc.u), ("v", v, oc.v), ("w", w, oc.w), ("p", p, oc.p)):
        _same(a, b, nm)
    s.close()


def test_porous_channel_64_cubed_40_steps_all_kernels(oracle):
    """S3 (64^3 porous channel, iter_max=50), 40 steps, through the half-sweep, the fused and the TMA kernels:
    every variant reproduces the oracle's fields and its per-step p error"""
    from pixelflow_b200 import Solver, workloads as wl
    m = n = l = 64
    dx, dy, dz, dt = wl.grid_spacing(0.063, 0.063, 0.063, 0.02, 400, m, n, l)   # dt = 5e-5: stable (see workloads.py)
    kw = dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=50, inlet_velocity=1.0, outlet_pressure=0.0, AoA=0.0)
    eps = wl.porous_channel(m, n, l, pitch=32)
    P = oracle.make_params(m=m, n=n, l=l, **kw)
    oc = oracle.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    oc.initialise()
    err_o = oc.step(40)
    assert err_o[-1] > 0 and np.isfinite(oc.u).all()
    for variant in (1, 3, 6):
        s = Solver("ibm3_uniform", m, n, l, sor_variant=variant, **kw)
        assert s.sor_variant == variant
        s.set_porosity(eps)
        s.initial_conditions()
        err_g = s.step(40)
        u, v, w, p = s.download()
        s.close()
        assert np.array_equal(err_o, err_g), variant
        for nm, a, b in (("u", u, oc.u), ("v", v, oc.v), ("w", w, oc.w), ("p", p, oc.p)):
            _same(a, b, f"variant {variant}: {nm}")
