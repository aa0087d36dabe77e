#!/usr/bin/env python
"""print a compact summary of bench.py JSON lines given as files"""
import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, "unreadable", e); continue
    r = d.get("roofline") or {}
    a = d.get("also") or {}
    e = d.get("e2e") or {}
    print(f"{fn}: {d['config']['workload']} N={d['n_gpus']} ms/step={d['ms_per_step']:.2f} sor_ms={d.get('ms_sor_per_step',0):.2f} "
          f"value={d['value']/1e6:.1f}M sweeps/s={d.get('sor_sweeps_per_s') or 0:.1f} frac={r.get('frac') or 0:.3f} "
          f"e2e={(e.get('value') or 0)/1e6:.1f}M launches={d.get('gpu_launches')} clocks={d.get('clocks')}"
          + (f" | also {a['workload']} ms={a['ms_per_step']:.2f} sweeps/s={a['sor_sweeps_per_s']:.1f} frac={(a.get('roofline') or {}).get('frac') or 0:.3f}" if a else ""))
