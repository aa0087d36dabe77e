"""The C oracle (oracle/pf_oracle.c) against the independent numpy restatement (oracle/oracle_np.py).

The reference ships no golden vectors for this path and cannot be compiled here (no Fortran
compiler): two separately written transcriptions of the source agreeing BIT FOR
BIT is what pins the oracle.  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_np as onp
from tests.conftest import rand_field, rand_porosity


def _np_params(P, **extra):
    kw = {k: getattr(P, k) for k in ("m", "n", "l", "dx", "dy", "dz", "dt", "xnue", "xlambda", "density",
                                     "thickness", "iter_max", "relux_factor", "inlet_velocity",
                                     "outlet_pressure", "AoA")}
    kw["nonslip"] = bool(P.nonslip)
    kw["wall"] = tuple(P.wall)
    kw.update(extra)
    return onp.Params(**kw)


CASES_3D = [
    # m, n, l, xlambda, nonslip, AoA, outlet_pressure
    (6, 6, 6, 0.0, 1, 0.0, 0.0),
    (5, 7, 5, 0.3, 1, 10.0, 0.0),
    (6, 5, 7, 0.0, 0, 0.0, 0.25),
    (7, 6, 5, 0.1, 1, -5.0, 0.0),
    (8, 4, 6, 0.0, 1, 0.0, 0.0),
    (16, 12, 10, 0.0, 1, 3.0, 0.0),
]


@pytest.mark.parametrize("m,n,l,xlambda,nonslip,AoA,pout", CASES_3D)
def test_ibm3_uniform_steps_bitwise(oracle, m, n, l, xlambda, nonslip, AoA, pout):
    rng = np.random.default_rng(1234 + m * 100 + n * 10 + l)
    P = oracle.make_params(m=m, n=n, l=l, dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=xlambda,
                           nonslip=nonslip, iter_max=7, relux_factor=1.7, inlet_velocity=1.0,
                           outlet_pressure=pout, AoA=AoA)
    eps_in = rand_porosity(rng, (l, n, m))
    oc = oracle.Oracle3D(P, False, eps_in)
    # numpy side
    Pn = _np_params(P)
    e = np.zeros(oc.shape)
    e[1:-1, 1:-1, 1:-1] = eps_in
    onp.porosity_halo_3d_uniform(Pn, e)
    assert np.array_equal(e, oc.e)
    st = onp.State3D(Pn, False, e)
    # start from a perturbed state so that every term is exercised
    for name in ("u", "v", "w", "p"):
        a = rand_field(rng, oc.shape, 0.1, 1.0 if name == "u" else 0.0)
        getattr(oc, name)[...] = a
        getattr(st, name)[...] = a
    oc.boundary()
    onp.boundary_3d_uniform(Pn, st.p, st.u, st.v, st.w)
    for name in ("u", "v", "w", "p"):
        assert np.array_equal(getattr(oc, name), getattr(st, name)), name
    for step in range(3):
        err_c = oc.step(1)[0]
        err_n = st.step()
        assert err_c == err_n, (step, err_c, err_n)
        for name in ("u", "v", "w", "p"):
            assert np.array_equal(getattr(oc, name), getattr(st, name)), (step, name)
        for name in ("ap", "ae", "aw", "an", "as", "at", "ab", "bb", "div"):
            assert np.array_equal(oc.ws.array(name), st.c[name]), (step, name)
    assert np.isfinite(oc.u).all()


WALLS = [
    (1, 0, 0, 0, 2, 0),   # shipped: top inlet, south outlet
    (0, 0, 0, 0, 0, 0),
    (2, 1, 2, 1, 1, 2),
    (1, 2, 1, 2, 2, 1),
    (0, 2, 2, 0, 1, 1),
]


@pytest.mark.parametrize("wall", WALLS)
@pytest.mark.parametrize("m,n,l", [(6, 7, 8), (9, 8, 7)])
def test_ibm3_air_condition_steps_bitwise(oracle, wall, m, n, l):
    rng = np.random.default_rng(99 + sum(wall) + m)
    P = oracle.make_params(m=m, n=n, l=l, dx=0.01, dy=0.01, dz=0.01, dt=5e-4, xnue=0.025, xlambda=0.0,
                           iter_max=6, inlet_velocity=1.5, outlet_pressure=0.1, wall=wall)
    eps_in = rand_porosity(rng, (l, n, m))
    # make sure some face cells are "fluid" (>= 0.9) and some are not
    eps_in[:, :, 0][::2] = 1.0
    eps_in[-1, ::2, :] = 0.95
    eps_in[:, 0, ::2] = 1.0
    oc = oracle.Oracle3D(P, True, eps_in)
    Pn = _np_params(P)
    e = np.zeros(oc.shape)
    e[1:-1, 1:-1, 1:-1] = eps_in
    onp.porosity_halo_3d_wall(Pn, e)
    assert np.array_equal(e, oc.e)
    st = onp.State3D(Pn, True, e)
    for name in ("u", "v", "w", "p"):
        a = rand_field(rng, oc.shape, 0.1)
        getattr(oc, name)[...] = a
        getattr(st, name)[...] = a
    oc.boundary()
    onp.boundary_3d_air(Pn, st.e, st.p, st.u, st.v, st.w)
    for name in ("u", "v", "w", "p"):
        assert np.array_equal(getattr(oc, name), getattr(st, name)), ("bc", name)
    for step in range(3):
        err_c = oc.step(1)[0]
        err_n = st.step()
        assert err_c == err_n
        for name in ("u", "v", "w", "p"):
            assert np.array_equal(getattr(oc, name), getattr(st, name)), (step, name)
        for name in ("ap", "ae", "aw", "an", "as", "at", "ab", "bb"):
            a, b = oc.ws.array(name)[1:-1, 1:-1, 1:-1], st.c[name][1:-1, 1:-1, 1:-1]
            assert np.array_equal(a, b), (step, name)


@pytest.mark.parametrize("backstep", [False, True])
@pytest.mark.parametrize("m,n", [(8, 6), (7, 6), (8, 5), (9, 7), (24, 16)])
def test_ibm2_steps_bitwise(oracle, backstep, m, n):
    rng = np.random.default_rng(7 + m * 31 + n)
    P = oracle.make_params(m=m, n=n, dx=1e-3, dy=1.1e-3, dt=2e-4, xnue=1e-3, xlambda=0.05, iter_max=9,
                           inlet_velocity=1.0, outlet_pressure=0.0, AoA=4.0)
    eps_in = rand_porosity(rng, (n, m))
    oc = oracle.Oracle2D(P, backstep, eps_in)
    Pn = _np_params(P)
    e = np.zeros(oc.shape)
    e[1:-1, 1:-1] = eps_in
    onp.porosity_halo_2d(Pn, e)
    assert np.array_equal(e, oc.e)
    p, u, v = (rand_field(rng, oc.shape, 0.1) for _ in range(3))
    oc.p[...], oc.u[...], oc.v[...] = p, u, v
    c = {nm: np.zeros(oc.shape) for nm in ("ap", "ae", "aw", "an", "as", "bb", "div")}
    for step in range(3):
        err_c = oc.step(1)[0]
        err_n = onp.step_2d(Pn, backstep, e, p, u, v, c)
        assert err_c == err_n
        assert np.array_equal(oc.p, p) and np.array_equal(oc.u, u) and np.array_equal(oc.v, v), step
        for nm in c:
            assert np.array_equal(oc.ws.array(nm), c[nm]), (step, nm)


def test_initial_conditions(oracle):
    P = oracle.make_params(m=5, n=4, l=3, dx=0.1, dy=0.1, dz=0.1, dt=0.01, xnue=1e-3, AoA=30.0,
                           inlet_velocity=2.0, outlet_pressure=0.5)
    oc = oracle.Oracle3D(P, False, np.ones((3, 4, 5)))
    oc.initialise()
    Pn = _np_params(P)
    z = lambda: np.zeros(oc.shape)
    p, u, v, w = z(), z(), z(), z()
    onp.initial_3d(Pn, False, p, u, v, w)
    onp.boundary_3d_uniform(Pn, p, u, v, w)
    for a, b in ((oc.p, p), (oc.u, u), (oc.v, v), (oc.w, w)):
        assert np.array_equal(a, b)


def test_force_log_2d_against_numpy(oracle):
    """output_force_log_2d (lib/output.f90:244-305): C oracle vs a vectorised numpy evaluation (sums to rounding)"""
    rng = np.random.default_rng(3)
    m, n = 24, 16
    P = oracle.make_params(m=m, n=n, dx=1e-3, dy=1.1e-3, dt=2e-4, xnue=1e-3, inlet_velocity=1.3, density=1.2)
    oc = oracle.Oracle2D(P, False, rand_porosity(rng, (n, m)))
    for a in (oc.p, oc.u, oc.v):
        a[...] = rand_field(rng, oc.shape, 0.3)
    out = oc.force_log(0.016)
    e, p, u, v = oc.e, oc.p, oc.u, oc.v
    c = (slice(1, -1), slice(1, -1))
    gx = (e[1:-1, 2:] - e[1:-1, :-2]) * 0.5
    gy = (e[2:, 1:-1] - e[:-2, 1:-1]) * 0.5
    na = np.sqrt(gx * gx + gy * gy)
    nx, ny = gx / np.maximum(na, 1e-6), gy / np.maximum(na, 1e-6)
    ec = e[c]
    fpx = np.sum(-P.dx * P.dy * p[c] * 2 * ec * (1.0 - ec) / (P.thickness * P.dx) * nx)
    fpy = np.sum(-P.dx * P.dy * p[c] * 2 * ec * (1.0 - ec) / (P.thickness * P.dy) * ny)
    fvx = np.sum(P.dx * P.dy * 32.0 * P.density * P.xnue * ((ec * (1.0 - ec)) / (P.thickness * P.dx)) ** 2 * u[c])
    fvy = np.sum(P.dx * P.dy * 32.0 * P.density * P.xnue * ((ec * (1.0 - ec)) / (P.thickness * P.dy)) ** 2 * v[c])
    den = P.density * P.inlet_velocity ** 2 * 0.016
    ref = np.array([fpx, fpy, fvx, fvy, fpx + fvx, fpy + fvy, (fpx + fvx) / den, (fpy + fvy) / den])
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-20)


def test_force_log_3d_against_numpy(oracle):
    """output_force_log_3d (lib/output.f90:1090-1165): C oracle vs a vectorised numpy evaluation (sums to rounding)"""
    rng = np.random.default_rng(4)
    m, n, l = 12, 9, 7
    P = oracle.make_params(m=m, n=n, l=l, dx=1e-3, dy=1.1e-3, dz=0.9e-3, dt=2e-4, xnue=1e-3, inlet_velocity=1.3,
                           density=1.2)
    oc = oracle.Oracle3D(P, False, rand_porosity(rng, (l, n, m)))
    for a in (oc.p, oc.u, oc.v, oc.w):
        a[...] = rand_field(rng, oc.shape, 0.3)
    out = oc.force_log(0.016)
    e, p = oc.e, oc.p
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    gx = (e[1:-1, 1:-1, 2:] - e[1:-1, 1:-1, :-2]) * 0.5
    gy = (e[1:-1, 2:, 1:-1] - e[1:-1, :-2, 1:-1]) * 0.5
    gz = (e[2:, 1:-1, 1:-1] - e[:-2, 1:-1, 1:-1]) * 0.5
    den = np.maximum(np.sqrt(gx * gx + gy * gy + gz * gz), 1e-6)
    ec, vol = e[c], P.dx * P.dy * P.dz
    fp = [np.sum(-vol * p[c] * 2 * ec * (1.0 - ec) / (P.thickness * d) * (g / den))
          for d, g in ((P.dx, gx), (P.dy, gy), (P.dz, gz))]
    fv = [np.sum(vol * 32.0 * P.density * P.xnue * ((ec * (1.0 - ec)) / (P.thickness * d)) ** 2 * q[c])
          for d, q in ((P.dx, oc.u), (P.dy, oc.v), (P.dz, oc.w))]
    f = [a + b for a, b in zip(fp, fv)]
    coef = P.density * P.inlet_velocity ** 2 * 0.016
    ref = np.array(fp + fv + f + [x / coef for x in f])
    assert np.allclose(out, ref, rtol=1e-11, atol=1e-20)
    assert np.abs(out[:6]).min() > 0
