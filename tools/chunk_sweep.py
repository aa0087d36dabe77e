#!/usr/bin/env python
"""Tuning experiment: SOR iteration time of the TMA kernel against the z-chunk size (PF_TMA_CHUNK), to fit the
per-block start-up cost of the chunk model (PF_TMA_CHUNK = uniform chunks; unset = pf_tma_schedule).  Prints one line per chunk size."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelflow_b200 import Solver, workloads as wl  # noqa: E402


def main():
    m, n, l = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 256, 256)))
    chunks = [int(x) for x in sys.argv[4:]] or [256, 128, 86, 64, 43, 32, 16]
    dx, dy, dz, dt = wl.grid_spacing(0.255, 0.255, 0.255, 0.02, 400, m, n, l)
    eps = wl.porous_channel(m, n, l)
    for cz in chunks:
        if cz > 0:
            os.environ["PF_TMA_CHUNK"] = str(cz)
        else:
            os.environ.pop("PF_TMA_CHUNK", None)   # 0 = the library's own choice
        s = Solver("ibm3_uniform", m, n, l, dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=100, sor_variant=6)
        s.set_porosity(eps)
        s.initial_conditions()
        s.step(2)
        s.step(5)
        t = s.last_timing()
        xt = -(-(((m + 1) // 2) + 2) // 30)
        yt = -(-n // 14)
        nz = -(-l // cz) if cz > 0 else 0
        blocks = xt * yt * nz
        print(f"grid {m}x{n}x{l} chunk {cz:4d}: blocks {blocks:5d} waves {blocks / 148:6.2f}  "
              f"sor iteration {t['ms_sor'] / 500 * 1e3:8.1f} us", flush=True)
        s.close()


if __name__ == "__main__":
    main()
