"""The restated f16.4 editing and record layout of the reference's VTK snapshots (oracle/oracle_np.py), on CPU."""
import numpy as np

from oracle import oracle_np as onp


def test_f16_4_known_answers():
    assert onp.f16_4(1.23456) == "          1.2346"
    assert onp.f16_4(-0.00001) == "         -0.0000"
    assert onp.f16_4(0.03125) == "          0.0312"          # an exact binary tie rounds to even, like printf
    assert onp.f16_4(0.09375) == "          0.0938"
    assert onp.f16_4(99999999999.0) == "99999999999.0000"
    assert onp.f16_4(-99999999999.0) == "*" * 16              # 17 columns needed
    assert onp.f16_4(float("nan")) == "             NaN" and onp.f16_4(float("-inf")) == "       -Infinity"


def test_record_layout():
    """(3(f16.4,1x)): three items -> 50 columns + newline (gfortran drops the trailing 1x), one item -> 16 + newline"""
    u = np.arange(5 * 4 * 6, dtype=float).reshape(5, 4, 6) / 7
    xp, yp, zp = np.arange(6) * 0.1, np.arange(4) * 0.2, np.arange(5) * 0.3
    pts = onp.vtk_section("points", 3, u, u, u, u, u, xp, yp, zp)
    assert len(pts) == 3 * 2 * 4 * 51
    first = pts.split(b"\n")[0]
    assert first == b"          0.1000           0.2000           0.3000"
    sc = onp.vtk_section("pressure", 3, u, u, u, u, u, xp, yp, zp)
    assert len(sc) == 24 * 17 and sc.split(b"\n")[0] == ("%16.4f" % u[1, 1, 1]).encode()
    two = onp.vtk_section("velocity", 2, u[0], u[1], None, u[0], u[0], xp, yp)
    assert two.split(b"\n")[0].endswith(b"          0.0000")   # the 2D files carry 0.0d0 as third component
