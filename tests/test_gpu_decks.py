"""The reference's shipped input decks (test/*.zip -> tests/golden/decks/*.npz, see make_decks.py) and
full-size property checks, on the GPU, through the C ABI and through the C++ twin driver."""
import os
import re
import subprocess

import numpy as np
import pytest

from pixelflow_b200 import workloads as wl
from pixelflow_b200.controldict import parse_controldict
from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TOL = 1e-10


def load_deck(name):
    z = np.load(os.path.join(HERE, "golden", "decks", name + ".npz"))
    cd = parse_controldict(str(z["controldict"]))
    m, n, l = (int(x) for x in z["dims"])
    eps = np.maximum(z["porosity"], cd.threshold)  # lib/grid.f90:50 / :289
    return cd, (m, n, l), eps


def deck_kwargs(cd, dims, d3):
    m, n, l = dims
    dx, dy, dz, dt = wl.grid_spacing(cd.width, cd.height, cd.depth, cd.time, cd.istep_max, m, n, l if d3 else 1)
    return dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=cd.xnue, xlambda=cd.xlambda, density=cd.density,
                thickness=cd.thickness, nonslip=cd.nonslip, iter_max=cd.iter_max, relux_factor=cd.relux_factor,
                inlet_velocity=cd.inlet_velocity, outlet_pressure=cd.outlet_pressure, AoA=cd.AoA)


def _same(a, b, what):
    assert np.array_equal(a, b), f"{what}: {int((a != b).sum())} values differ, relL2={rel_l2(a, b):.3e}"
    assert rel_l2(a, b) <= TOL


@pytest.mark.parametrize("deck,case,backstep", [("cylinder", "ibm2_uniform", False), ("cylinder", "ibm2_drag", False),
                                                ("backstep", "ibm2_backstep", True)])
def test_2d_decks(oracle, deck, case, backstep):
    """configs[0] cylinder-2d (default controlDict) and configs[1] backstep (iter_max=100, w=1.7)"""
    from pixelflow_b200 import Solver
    cd, (m, n, _), eps = load_deck(deck)
    assert (cd.iter_max, cd.relux_factor) == (100, 1.7)
    kw = deck_kwargs(cd, (m, n, 1), False)
    P = oracle.make_params(m=m, n=n, **kw)
    oc = oracle.Oracle2D(P, backstep, eps[0])
    oc.initialise()
    s = Solver(case, m, n, **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    nsteps = 3
    err_o, err_g = oc.step(nsteps), s.step(nsteps)
    u, v, _, p = s.download()
    _same(u, oc.u, "u"); _same(v, oc.v, "v"); _same(p, oc.p, "p")
    assert np.array_equal(err_o, err_g)
    assert np.isfinite(p).all() and err_g[-1] > 0
    s.close()


def test_room_deck(oracle):
    """configs[2]: room air-conditioning 3D, ibm3_air_condition, shipped wall_conditions"""
    from pixelflow_b200 import Solver
    cd, (m, n, l), eps = load_deck("room")
    kw = deck_kwargs(cd, (m, n, l), True)
    P = oracle.make_params(m=m, n=n, l=l, wall=(1, 0, 0, 0, 2, 0), **kw)
    oc = oracle.Oracle3D(P, True, eps)
    oc.initialise()
    s = Solver("ibm3_air_condition", m, n, l, wall=(1, 0, 0, 0, 2, 0), **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    err_o, err_g = oc.step(4), s.step(4)
    u, v, w, p = s.download()
    for nm, a, b in (("u", u, oc.u), ("v", v, oc.v), ("w", w, oc.w), ("p", p, oc.p)):
        _same(a, b, nm)
    assert np.array_equal(err_o, err_g)
    assert np.abs(w).max() > 0  # the inlet drives a flow
    s.close()


def test_uniform_flow_is_a_fixed_point_at_256_cubed():
    """SURVEY.md 8c invariant at BASELINE size: eps == 1, outlet_pressure == 0 -> u=Uin, v=w=p=0 stays put"""
    from pixelflow_b200 import Solver
    m = n = l = 256
    dx, dy, dz, dt = wl.grid_spacing(0.255, 0.255, 0.255, 0.02, 100, m, n, l)
    s = Solver("ibm3_uniform", m, n, l, dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=20)
    s.set_porosity(np.ones(s.shape))
    s.initial_conditions()
    err = s.step(2)
    u, v, w, p = s.download()
    assert (u == 1.0).all() and (v == 0).all() and (w == 0).all() and (p == 0).all()
    assert (err == 0).all()
    s.close()


def test_full_size_256_cubed_against_the_oracle_and_across_kernel_variants(oracle):
    """BASELINE size (256^3 porous channel): two steps with iter_max=10 bit-identical to the CPU oracle,
    and identical across the SOR kernel variants / CUDA-graph replay (implementation independence).
    (Translation invariance is NOT a property of this path: the reference's stale periodic halos,
    SURVEY.md H2, make the seam a special place -- verified, see DESIGN.md.)"""
    from pixelflow_b200 import Solver
    m = n = l = 256
    dx, dy, dz, dt = wl.grid_spacing(0.255, 0.255, 0.255, 0.02, 100, m, n, l)
    kw = dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=10, inlet_velocity=1.0, outlet_pressure=0.0, AoA=0.0)
    eps = wl.porous_channel(m, n, l)
    P = oracle.make_params(m=m, n=n, l=l, **kw)
    oc = oracle.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    assert np.array_equal(oc.e, eps)
    oc.initialise()
    err_o = oc.step(2)
    for variant, graph in ((0, 1), (1, 1), (1, 0), (2, 1), (3, 1), (4, 0), (6, 1), (6, 0)):
        s = Solver("ibm3_uniform", m, n, l, sor_variant=variant, use_graph=graph, **kw)
        assert s.sor_variant == (variant or 6)
        s.set_porosity(eps)
        s.initial_conditions()
        err_g = s.step(2)
        u, v, w, p = s.download()
        s.close()
        for nm, a, b in (("u", u, oc.u), ("v", v, oc.v), ("w", w, oc.w), ("p", p, oc.p)):
            _same(a, b, f"256^3 variant {variant} graph {graph}: {nm}")
        assert np.array_equal(err_o, err_g)
    assert err_o[-1] > 0


def test_cpp_twin_driver_runs_the_room_deck(oracle, tmp_path):
    """the drop-in driver: controlDict + CSV in, the reference's log lines out, same p error as the oracle"""
    from pixelflow_b200 import build
    build.build_drivers()
    exe = os.path.join(ROOT, "pixelflow_b200", "driver", "bin", "ibm3_air_condition_omp")
    z = np.load(os.path.join(HERE, "golden", "decks", "room.npz"))
    m, n, l = (int(x) for x in z["dims"])
    (tmp_path / "config").mkdir()
    (tmp_path / "data").mkdir()
    (tmp_path / "config" / "controlDict.txt").write_text(str(z["controldict"]))
    e = z["porosity"]
    with open(tmp_path / "data" / "room.csv", "w") as f:
        f.write(f"{m},{n},{l}\n")
        for k in range(l):
            for j in range(n):
                f.write("".join(f"{i + 1},{j + 1},{k + 1},{e[k, j, i]:.6E}\n" for i in range(m)))
    r = subprocess.run([exe, "--steps", "3", "--cache"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "# --- MAC algorithm start" in out and "program finished" in out
    assert len(re.findall(r"--- time_steps=", out)) == 3
    errs = [float(x) for x in re.findall(r"SOR iteration no\.\s+100 -- p error:\s*([-+0-9.E]+)", out)]
    assert len(errs) == 3
    for f in ("etc/grid.dat", "etc/solution_uvp.dat", "etc/divergent.dat", "etc/surface_profile.dat",
              "room/output_00000.vtk", "room/output_paraview.vtk"):
        assert (tmp_path / f).exists(), f
    # same numbers as the oracle run of the same deck (the CSV round trip .6E is exact for this deck)
    cd, dims, eps = load_deck("room")
    kw = deck_kwargs(cd, dims, True)
    P = oracle.make_params(m=m, n=n, l=l, wall=(1, 0, 0, 0, 2, 0), **kw)
    oc = oracle.Oracle3D(P, True, eps)
    oc.initialise()
    err_o = oc.step(3)
    assert np.allclose(errs, err_o, rtol=1e-15, atol=0), (errs, err_o)
    vtk = (tmp_path / "room" / "output_paraview.vtk").read_text().splitlines()
    assert vtk[0] == "# vtk DataFile Version 3.0" and vtk[4].split() == ["DIMENSIONS", "64", "64", "64"]
    # bodies formatted on the GPU (pf_vtk_section): 3 vector sections of 51-byte records, 3 scalar ones of 17
    assert len(vtk) == 5 + 1 + 1 + 1 + 1 + 2 * 3 + 6 * 64 ** 3
    assert len(vtk[6]) == 50 and len(vtk[6].split()) == 3 and len(vtk[-1]) == 16
    size = (tmp_path / "room" / "output_paraview.vtk").stat().st_size
    assert 64 ** 3 * 204 < size < 64 ** 3 * 204 + 400
    assert "s in VTK snapshots" in r.stderr
    # the final file orders its scalars pressure, VelocityDivergent, porosity (lib/output.f90:859-907); snapshots
    # write porosity first (:1029-1075)
    heads = [ln for ln in vtk if ln.startswith(("SCALARS", "VECTORS"))]
    assert [h.split()[1] for h in heads] == ["velocity", "velocityInFluid", "pressure", "VelocityDivergent", "porosity"]
    snap = [ln.split()[1] for ln in (tmp_path / "room" / "output_00000.vtk").read_text().splitlines()
            if ln.startswith(("SCALARS", "VECTORS"))]
    assert snap == ["velocity", "velocityInFluid", "porosity", "pressure", "VelocityDivergent"]
    sol = (tmp_path / "etc" / "solution_uvp.dat").read_text().splitlines()
    assert sol[0].split()[:4] == ["m,", "n,", "l", "="] and sol[1].strip() == "velocity u_bulk"
    assert len(sol) == 1 + 9 * (1 + 64) and len(sol[2].split()) == 64 * 64      # one record per k-plane
    dv = (tmp_path / "etc" / "divergent.dat").read_text().splitlines()
    assert dv[0] == "" and dv[1].strip() == "porosity" and dv[2 + 64 * 64 + 1].strip() == "divergent velocity"
    # second run: the parsed CSV comes from the (opt-in: --cache) binary copy data/room.csv.pfbin: the same numbers
    assert (tmp_path / "data" / "room.csv.pfbin").exists()
    r2 = subprocess.run([exe, "--steps", "3", "--no-output", "--cache"], cwd=tmp_path, capture_output=True, text=True,
                        timeout=600)
    assert r2.returncode == 0, r2.stderr[-2000:]
    errs2 = [float(x) for x in re.findall(r"SOR iteration no\.\s+100 -- p error:\s*([-+0-9.E]+)", r2.stdout)]
    assert errs2 == errs


def test_drag_force_log_on_the_cylinder_deck(oracle):
    """ibm2_drag calls output_force_log_2d after every step (ibm_2d_drag_omp_cpu.f90:121): pf_force_log_2d
    against the oracle's serial sums.  Per-cell terms are exact; the summation order differs -> 1e-12."""
    from pixelflow_b200 import Solver
    cd, (m, n, _), eps = load_deck("cylinder")
    kw = deck_kwargs(cd, (m, n, 1), False)
    P = oracle.make_params(m=m, n=n, **kw)
    oc = oracle.Oracle2D(P, False, eps[0])
    oc.initialise()
    s = Solver("ibm2_drag", m, n, **kw)
    s.set_porosity(oc.e)
    s.initial_conditions()
    for _ in range(2):
        oc.step(1)
        s.step(1)
        fo = oc.force_log(cd.radius)
        fg = s.force_log_2d(cd.radius)["raw"]
        # the lift components cancel to ~1e-6 of the summed magnitudes: tolerance relative to the force scale
        scale = np.abs(fo[:4]).max()
        assert np.allclose(fg[:6], fo[:6], rtol=1e-12, atol=1e-13 * scale), (fg, fo)
        assert np.allclose(fg[6:], fo[6:], rtol=1e-12, atol=1e-13 * abs(fo[6])), (fg, fo)
        assert abs(fo[6]) > 0  # a drag coefficient comes out
    s.close()
