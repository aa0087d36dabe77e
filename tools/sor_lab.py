#!/usr/bin/env python
"""Kernel lab: SOR iteration time of one library build / SOR variant on the porous channel, as one JSON line.

    python tools/sor_lab.py --variant 8 --grid 256 256 256 [--lib exp/libpf_x.so] [--iters 100] [--steps 5] [--check]

Times `steps` whole time steps (CUDA events inside the library) and reports the SOR phase per iteration, the DRAM
bytes the kernel's layout has to move (48 B/cell/sweep for the fused kernels, 88 for the half-sweeps) as GB/s, and --
with --check -- whether u, v, w, p and the p errors after the steps equal those of the half-sweep kernel (variant 1)
of the PRODUCT library bit for bit.  An alternative build is chosen with --lib (PIXELFLOW_GPU_LIB); experiment builds
with PF_TMA_NOCOMPUTE / PF_TMA_NOLOAD compute garbage by design (never --check them).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args):
    sys.path.insert(0, ROOT)
    import numpy as np
    from pixelflow_b200 import Solver, workloads as wl
    m, n, l = args.grid
    w = 0.001 * (m - 1)
    dx, dy, dz, dt = wl.grid_spacing(w, 0.001 * (n - 1), 0.001 * (l - 1), 0.02, 400, m, n, l)
    eps = wl.porous_channel(m, n, l)
    s = Solver("ibm3_uniform", m, n, l, dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=args.iters,
               sor_variant=args.variant, use_graph=args.graph)
    s.set_porosity(eps)
    s.initial_conditions()
    s.step(args.warmup)
    errs = s.step(args.steps)
    t = s.last_timing()
    v = s.sor_variant
    per_iter_us = t["ms_sor"] / (args.steps * args.iters) * 1e3
    bytes_cell = {1: 88.0, 5: 88.0, 7: 88.0}.get(v, 48.0)
    out = {"lib": os.environ.get("PIXELFLOW_GPU_LIB", "product"), "variant": v, "grid": [m, n, l],
           "sor_us_per_iteration": per_iter_us, "sweeps_per_s": 1e6 / per_iter_us,
           "layout_gbs": bytes_cell * m * n * l / (per_iter_us * 1e-6) / 1e9,
           "ms_per_step": t["ms_total"] / args.steps, "p_error_last": float(errs[-1])}
    if args.check:
        h = hashlib.sha256()
        for a in s.download():
            h.update(np.ascontiguousarray(a).data)
        h.update(np.ascontiguousarray(errs).data)
        out["sha"] = h.hexdigest()[:16]
    s.close()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--grid", type=int, nargs=3, default=(256, 256, 256))
    ap.add_argument("--lib", default=None)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child or not args.lib:
        return run(args)
    env = dict(os.environ, PIXELFLOW_GPU_LIB=os.path.join(ROOT, args.lib) if not os.path.isabs(args.lib) else args.lib)
    cmd = [sys.executable, os.path.abspath(__file__), "--child"] + [a for a in sys.argv[1:]]
    return subprocess.call(cmd, env=env)


if __name__ == "__main__":
    sys.exit(main())
