#!/usr/bin/env python
"""The reference's three shipped decks on the GPU path: parity against the reference's own outputs AND timing, as JSON.

    python tools/decks_probe.py [--sor-variant 0|7] [--steps 20]

For each deck (cylinder 1024x512 ibm2, backstep 2251x411 ibm2_backstep, room 64^3 ibm3_air_condition; inputs from
tests/golden/decks/*.npz): run the first 3 steps of the unmodified deck and compare SHA-256 of u, v, [w,] p and the
logged p errors with tests/golden/ref_translated.npz (what the translated reference program left behind), then time
`--steps` more steps.  One JSON object per deck on stdout; nothing from oracle/ is used.
bench.py calls this in a subprocess (with a timeout) so that its JSON line also carries the decks, and — labelled
--sor-variant 7 / 8 time the two deck kernels (persistent half-sweeps, temporally blocked tiles) the same way.
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sor-variant", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from pixelflow_b200 import Solver, workloads as wl
    from pixelflow_b200.controldict import parse_controldict
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_translated.npz"))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for deck, case in (("cylinder", "ibm2_uniform"), ("backstep", "ibm2_backstep"), ("room", "ibm3_air_condition")):
        z = np.load(os.path.join(ROOT, "tests", "golden", "decks", deck + ".npz"))
        cd = parse_controldict(str(z["controldict"]))
        m, n, l = (int(x) for x in z["dims"])
        d3 = case.startswith("ibm3")
        sp = [float(x) for x in gold[f"deck_{deck}/spacing"]]
        kw = dict(xnue=cd.xnue, xlambda=cd.xlambda, density=cd.density, thickness=cd.thickness, nonslip=cd.nonslip,
                  iter_max=cd.iter_max, relux_factor=cd.relux_factor, inlet_velocity=cd.inlet_velocity,
                  outlet_pressure=cd.outlet_pressure, AoA=cd.AoA, sor_variant=args.sor_variant)
        if d3:
            s = Solver(case, m, n, l, dx=sp[0], dy=sp[1], dz=sp[2], dt=sp[3], **kw)
        else:
            s = Solver(case, m, n, dx=sp[0], dy=sp[1], dt=sp[2], **kw)
        eps = np.maximum(z["porosity"] if d3 else z["porosity"][0], cd.threshold)
        s.set_porosity(wl.with_halos(eps, case))
        s.initial_conditions()
        nref = int(gold[f"deck_{deck}/steps"])
        errs = np.array([s.step(1)[0] for _ in range(nref)])
        u, v, w, p = s.download()
        want = json.loads(str(gold[f"deck_{deck}/sha"]))
        same = sha(u) == want["u"] and sha(v) == want["v"] and sha(p) == want["p"] and (not d3 or sha(w) == want["w"])
        same = bool(same and np.array_equal(errs, gold[f"deck_{deck}/perr"]))
        s.step(3)
        s.step(args.steps)
        t = s.last_timing()
        print(json.dumps({"deck": deck, "solver": case, "grid": [m, n, l], "iter_max": cd.iter_max,
                          "sor_variant": s.sor_variant, "bit_identical_to_reference_after_3_steps": same,
                          "ms_per_step": t["ms_total"] / args.steps, "ms_sor_per_step": t["ms_sor"] / args.steps,
                          "cell_updates_per_s": m * n * (l if d3 else 1) * args.steps / (t["ms_total"] * 1e-3)}), flush=True)
        s.close()


if __name__ == "__main__":
    main()
