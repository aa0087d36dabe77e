#!/usr/bin/env python
"""Time of one ASCII VTK snapshot body (all sections of output_paraview_temp_3d) formatted on the GPU, against
formatting the same records on the host with printf-style conversion (what a CPU driver does), 256^3 dragon."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelflow_b200 import Solver, workloads as wl  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    occ = wl.load_occupancy(os.path.join(ROOT, "tests", "golden", f"dragon_voxels_{n}.npz"))
    eps = wl.porosity_from_occupancy(occ)
    width = {64: 0.063, 256: 0.255}[n]
    dx, dy, dz, dt = wl.grid_spacing(width, width, width, 0.02, 400, n, n, n)
    s = Solver("ibm3_uniform", n, n, n, dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=20)
    s.set_porosity(eps)
    s.initial_conditions()
    s.step(2)
    xp, yp, zp = (np.arange(n + 2) * d for d in (dx, dy, dz))
    secs = ("points", "velocity", "velocityInFluid", "porosity", "pressure", "VelocityDivergent")
    s.vtk_section("pressure", xp, yp, zp, 1, 8)   # warm-up (allocations)
    t0 = time.perf_counter()
    total = 0
    chunk = 64
    for sec in secs:
        for k0 in range(1, n + 1, chunk):
            total += len(s.vtk_section(sec, xp, yp, zp, k0, min(chunk, n - k0 + 1)))
    t_gpu = time.perf_counter() - t0
    # host formatting of a sample: 2^20 values through the same conversion (numpy's %-formatting = C printf)
    u, v, w, p = s.download()
    sample = p[1:-1, 1:-1, 1:-1].ravel()[:1 << 20]
    t0 = time.perf_counter()
    np.char.mod("%16.4f", sample)
    t_host = (time.perf_counter() - t0) / sample.size * (n ** 3 * 12)     # 12 values per cell and snapshot
    print(f"snapshot {n}^3: {total / 1e9:.3f} GB of text; GPU format + copy to host {t_gpu:.3f} s "
          f"({total / t_gpu / 1e9:.2f} GB/s); host formatting of the same 12 values/cell at the sampled rate: {t_host:.1f} s")
    s.close()


if __name__ == "__main__":
    main()
