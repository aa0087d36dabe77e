#!/usr/bin/env python
"""voxel -> porosity timing / profiling driver: python tools/voxel_run.py [N=256] [repeats=2]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelflow_b200 import workloads as wl  # noqa: E402
from pixelflow_b200.voxel2poro import create_tanh_kernel, convolve_nearest  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    occ = wl.load_occupancy(os.path.join(ROOT, "tests", "golden", f"dragon_voxels_{n}.npz"))
    k = create_tanh_kernel(1.5)
    for r in range(reps):
        t0 = time.perf_counter()
        por = convolve_nearest(occ, k)
        dt = time.perf_counter() - t0
        taps = occ.size * k.size
        print(f"voxel2poro {n}^3 x {k.shape[0]}^3 taps: {dt:.3f} s wall (host arrays in/out), {taps / dt / 1e12:.3f} T taps/s, "
              f"porosity range [{por.min():.3e}, {por.max():.3f}]", flush=True)


if __name__ == "__main__":
    main()
