#!/usr/bin/env python
"""Timings of the reference's shipped decks (BASELINE.json configs[0..2]) on the GPU path and, beside
them, of the restated CPU reference on the same box -- parity-test configurations, reported for
completeness (they are launch-latency bound: 0.26-0.93 M cells).  Prints one JSON line per deck.

    python tools/bench_decks.py [--steps 20] [--cpu-steps 3] [--sor-variant 7]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pixelflow_b200 import Solver, workloads as wl  # noqa: E402
from pixelflow_b200.controldict import parse_controldict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--sor-variant", type=int, default=0, help="0 auto; 7 = persistent half-sweeps, 8 = temporally blocked tiles (2D)")
    ap.add_argument("--use-graph", type=int, default=1)
    args = ap.parse_args()
    from oracle import oracle_c  # CPU baseline only
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    decks = [("cylinder", "ibm2_uniform", False), ("backstep", "ibm2_backstep", False), ("room", "ibm3_air_condition", True)]
    for name, case, d3 in decks:
        z = np.load(os.path.join(ROOT, "tests", "golden", "decks", name + ".npz"))
        cd = parse_controldict(str(z["controldict"]))
        m, n, l = (int(x) for x in z["dims"])
        eps = np.maximum(z["porosity"], cd.threshold)
        dx, dy, dz, dt = wl.grid_spacing(cd.width, cd.height, cd.depth, cd.time, cd.istep_max, m, n, l if d3 else 1)
        kw = dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=cd.xnue, xlambda=cd.xlambda, density=cd.density,
                  thickness=cd.thickness, nonslip=cd.nonslip, iter_max=cd.iter_max, relux_factor=cd.relux_factor,
                  inlet_velocity=cd.inlet_velocity, outlet_pressure=cd.outlet_pressure, AoA=cd.AoA)
        if d3:
            P = oracle_c.make_params(m=m, n=n, l=l, wall=(1, 0, 0, 0, 2, 0), **kw)
            oc = oracle_c.Oracle3D(P, True, eps)
            s = Solver(case, m, n, l, wall=(1, 0, 0, 0, 2, 0), sor_variant=args.sor_variant, use_graph=args.use_graph, **kw)
        else:
            P = oracle_c.make_params(m=m, n=n, **kw)
            oc = oracle_c.Oracle2D(P, case == "ibm2_backstep", eps[0])
            s = Solver(case, m, n, sor_variant=args.sor_variant, use_graph=args.use_graph, **kw)
        oc.initialise()
        s.set_porosity(oc.e)
        s.initial_conditions()
        s.step(3)
        s.step(args.steps)
        t = s.last_timing()
        t0 = time.perf_counter()
        oc.step(args.cpu_steps)
        cpu_ms = (time.perf_counter() - t0) / args.cpu_steps * 1e3
        cells = m * n * (l if d3 else 1)
        print(json.dumps({"deck": name, "solver": case, "grid": [m, n, l], "iter_max": cd.iter_max,
                          "sor_variant": s.sor_variant,
                          "gpu_ms_per_step": t["ms_total"] / args.steps, "gpu_ms_sor_per_step": t["ms_sor"] / args.steps,
                          "gpu_cell_updates_per_s": cells * args.steps / (t["ms_total"] * 1e-3),
                          "gpu_sor_sweeps_per_s": args.steps * cd.iter_max / (t["ms_sor"] * 1e-3),
                          "cpu_ms_per_step": cpu_ms, "cpu_cores": int(os.environ["OMP_NUM_THREADS"]),
                          "cpu_cell_updates_per_s": cells / (cpu_ms * 1e-3), "speedup": cpu_ms / (t["ms_total"] / args.steps)}))
        s.close()


if __name__ == "__main__":
    main()
