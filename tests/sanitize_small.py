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
    for (m, n, l) in ((20, 12, 8), (130, 16, 12), (7, 6, 5)):
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
    for case in ("ibm2_uniform", "ibm2_backstep", "ibm2_drag"):
        m, n = 40, 18
        s = Solver(case, m, n, dx=1e-3, dy=1e-3, dt=2e-4, xnue=1e-3, iter_max=4)
        s.set_porosity(wl.cylinder_2d(m, n))
        s.initial_conditions()
        s.step(2)
        s.force_log_2d(0.01)
        s.close()
    print("sanitize_small: done")


if __name__ == "__main__":
    main()
