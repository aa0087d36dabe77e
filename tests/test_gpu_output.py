"""ASCII VTK snapshot bodies formatted on the GPU (pf_vtk_section, csrc/pf_output.cu) against the numpy/Python
restatement of lib/output.f90:968-1088 / :421-537 (oracle/oracle_np.py: vtk_section, f16_4): byte for byte."""
import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _coords(m, n, l, dx=0.01, dy=0.011, dz=0.009):
    return np.arange(m + 2) * dx, np.arange(n + 2) * dy, np.arange(l + 2) * dz


def test_f16_4_torture():
    """exact decimal rounding incl. ties, signs, tiny and huge magnitudes, overflow asterisks, NaN / Infinity"""
    from pixelflow_b200 import Solver
    m, n, l = 64, 16, 16
    rng = np.random.default_rng(11)
    vals = rng.standard_normal(m * n * l) * 10.0 ** rng.integers(-9, 13, m * n * l)
    special = [0.0, -0.0, 0.03125, -0.03125, 0.09375, 0.00005, 0.00015, 0.00025, -0.00005, 1e-300, -1e-300, 5e-324,
               0.99995, 0.999949999, 9.99995, 99999999999.0, 99999999999.99994, 99999999999.99996, -9999999999.99995,
               -99999999999.0, 1e11, 1.0e15, 1e300, -1e300, np.inf, -np.inf, np.nan, 2.0 ** -20, 12345.67895, 0.5, 1.5,
               2.5e-5, 7.5e-5, 123456789.12345, -0.00004999, 4.9999999e-5, 5.0000001e-5]
    ties = (np.arange(1, 4001) * 2 + 1) / 32.0 * 1e-3          # near-ties: decided by the exact binary expansion
    ties2 = np.arange(-2000, 2000) / 8.0 + 0.03125               # k/8 + 1/32: exactly representable, 5th decimal is 5
    vals[:len(special)] = special
    vals[100:100 + len(ties)] = ties
    vals[5000:5000 + len(ties2)] = ties2
    s = Solver("ibm3_uniform", m, n, l, dx=0.01, dy=0.011, dz=0.009, dt=1e-4, xnue=1e-3)
    s.set_porosity(np.ones(s.shape))
    p = np.zeros(s.shape)
    p[1:-1, 1:-1, 1:-1] = vals.reshape(l, n, m)
    s.set_field("p", p)
    xp, yp, zp = _coords(m, n, l)
    got = s.vtk_section("pressure", xp, yp, zp)
    s.close()
    want = onp.vtk_section("pressure", 3, p, p, p, p, p, xp, yp, zp)
    assert len(got) == m * n * l * 17
    if got != want:
        g, w = got.split(b"\n"), want.split(b"\n")
        bad = [(float(vals[q]).hex(), g[q], w[q]) for q in range(len(w)) if g[q] != w[q]][:10]
        raise AssertionError(bad)
    lines = got.split(b"\n")
    assert lines[0] == b"          0.0000" and lines[1] == b"         -0.0000" and lines[2] == b"          0.0312"
    assert lines[15] == b"99999999999.0000" and lines[19] == b"****************" and lines[24] == b"        Infinity"
    assert lines[25] == b"       -Infinity" and lines[26] == b"             NaN"


@pytest.mark.parametrize("case,m,n,l", [("ibm3_uniform", 20, 12, 8), ("ibm3_air_condition", 9, 7, 5)])
def test_all_3d_sections_after_a_few_steps(oracle, case, m, n, l):
    from pixelflow_b200 import Solver
    rng = np.random.default_rng(m)
    kw = dict(dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, iter_max=8, inlet_velocity=1.0)
    P = oracle.make_params(m=m, n=n, l=l, wall=(1, 0, 0, 0, 2, 0), **kw)
    air = case == "ibm3_air_condition"
    oc = oracle.Oracle3D(P, air, np.clip(rng.random((l, n, m)), 1e-6, 1.0))
    oc.initialise()
    s = Solver(case, m, n, l, wall=(1, 0, 0, 0, 2, 0), **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    oc.step(3)
    s.step(3)
    xp, yp, zp = _coords(m, n, l)
    for sec in ("points", "velocity", "velocityInFluid", "porosity", "pressure", "VelocityDivergent"):
        got = s.vtk_section(sec, xp, yp, zp)
        want = onp.vtk_section(sec, 3, oc.u, oc.v, oc.w, oc.p, oc.e, xp, yp, zp)
        assert got == want, sec
        # plane chunks concatenate to the whole body (how the driver bounds its buffers)
        parts = b"".join(s.vtk_section(sec, xp, yp, zp, k_local0=k0, nplanes=min(3, l - k0 + 1)) for k0 in range(1, l + 1, 3))
        assert parts == want, sec
    with pytest.raises(Exception, match="section"):
        s.vtk_section("dimless_v", xp, yp, zp)
    s.close()


def test_all_2d_sections(oracle):
    from pixelflow_b200 import Solver
    m, n = 300, 37          # more than one 256-record block, ragged tail
    rng = np.random.default_rng(2)
    kw = dict(dx=0.01, dy=0.011, dt=2e-4, xnue=1e-3, iter_max=8, inlet_velocity=0.7)
    P = oracle.make_params(m=m, n=n, **kw)
    oc = oracle.Oracle2D(P, False, np.clip(rng.random((n, m)), 1e-6, 1.0))
    oc.initialise()
    s = Solver("ibm2_uniform", m, n, **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    oc.step(2)
    s.step(2)
    xp, yp, _ = _coords(m, n, 1)
    for sec in ("points", "velocity", "velocityInFluid", "dimless_v", "porosity", "pressure", "VelocityDivergent",
                "abs_dimless_v"):
        got = s.vtk_section(sec, xp, yp)
        want = onp.vtk_section(sec, 2, oc.u, oc.v, None, oc.p, oc.e, xp, yp, None, inlet_velocity=0.7)
        assert got == want, sec
    s.close()
