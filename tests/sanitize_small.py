"""Small end-to-end runs of every SOR variant and every solver case, meant to be run under
`compute-sanitizer --tool memcheck` (and racecheck) on the GPU box:

    compute-sanitizer --tool memcheck python tests/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pixelflow_b200 import Solver, workloads as wl  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    for (m, n, l) in ((20, 12, 8), (130, 16, 12), (7, 6, 5), (70, 32, 44)):   # the last: unrolled steady state of the TMA kernel
        dx, dy, dz, dt = wl.grid_spacing(0.1, 0.1, 0.1, 0.02, 100, m, n, l)
        for variant in (1, 2, 3, 4, 6):
            s = Solver("ibm3_uniform", m, n, l, dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=4, sor_variant=variant)
            e = np.zeros(s.shape)
            e[1:-1, 1:-1, 1:-1] = np.clip(rng.random((l, n, m)), 1e-6, 1.0)
            wl.porosity_halo_3d_periodic(e)
            s.set_porosity(e)
            s.initial_conditions()
            err = s.step(2)
            assert np.isfinite(err).all()
            s.close()
    m, n, l = 12, 10, 8
    s = Solver("ibm3_air_condition", m, n, l, dx=0.01, dy=0.01, dz=0.01, dt=5e-4, xnue=0.025, iter_max=4,
               inlet_velocity=1.5)
    s.set_porosity(wl.room_like(m, n, l))
    s.initial_conditions()
    s.step(2)
    s.close()
    s = Solver("ibm3_air_condition", m, n, l, dx=0.01, dy=0.01, dz=0.01, dt=5e-4, xnue=0.025, iter_max=4,
               inlet_velocity=1.5, sor_variant=7)
    s.set_porosity(wl.room_like(m, n, l))
    s.initial_conditions()
    s.step(2)
    s.close()
    for case in ("ibm2_uniform", "ibm2_backstep", "ibm2_drag"):
        for (m, n) in ((40, 18), (150, 37)):
            for variant in (1, 7, 8):            # half-sweep chain, persistent half-sweeps, temporally blocked tiles
                s = Solver(case, m, n, dx=1e-3, dy=1e-3, dt=2e-4, xnue=1e-3, iter_max=6, sor_variant=variant)
                s.set_porosity(wl.cylinder_2d(m, n))
                s.initial_conditions()
                s.step(2)
                s.force_log_2d(0.01)
                s.close()
    # input preparation kernels
    from pixelflow_b200 import stl2poro
    tri = rng.normal(0, 1, (200, 3, 3)).astype(np.float32)
    stl2poro.calculate_sdf(tri, rng.normal(0, 1.5, (3000, 3)))
    print("sanitize_small: done")


if __name__ == "__main__":
    main()
