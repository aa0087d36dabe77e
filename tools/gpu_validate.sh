#!/usr/bin/env bash
# One gpurun call that answers "is the tree healthy on a B200?": GPU tests, smoke, the default bench line, the
# reference arm and the ncu launch list of the default command.  Everything lands in gpurun_out/<tag>_*.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh r02a'
# Optional second argument: extra pytest arguments (e.g. "-k ref_golden").
set -u
tag=${1:-check}
extra=${2:-}
out=gpurun_out
mkdir -p "$out"
echo "== pytest -m gpu" | tee "$out/${tag}_summary.txt"
timeout 1500 python -m pytest tests -m gpu -q -x $extra > "$out/${tag}_pytest.log" 2>&1
echo "pytest rc=$? : $(tail -1 "$out/${tag}_pytest.log")" | tee -a "$out/${tag}_summary.txt"
echo "== smoke" | tee -a "$out/${tag}_summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke rc=$? : $(tail -1 "$out/${tag}_smoke.log")" | tee -a "$out/${tag}_summary.txt"
echo "== bench (default)" | tee -a "$out/${tag}_summary.txt"
timeout 900 python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
echo "bench rc=$?" | tee -a "$out/${tag}_summary.txt"
python - "$out/${tag}_bench.json" <<'PY' | tee -a "$out/${tag}_summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r, c = d["roofline"], d.get("cpu_baseline") or {}
    print(f"  {d['config']['workload']}: {d['ms_per_step']:.1f} ms/step, {d['value']/1e6:.1f} M cell-updates/s, "
          f"{d['sor_sweeps_per_s']:.1f} sweeps/s, roofline {r['frac']:.3f} ({r['kernel']}), e2e "
          f"{(d.get('e2e') or {}).get('value', 0)/1e6:.1f} M, cpu_baseline {c.get('kind')} {c.get('value', 0)/1e6:.2f} M on "
          f"{c.get('cores')} cores, clocks {d.get('clocks')}")
except Exception as e:
    print("  no bench line:", e)
PY
echo "== bench --impl reference" | tee -a "$out/${tag}_summary.txt"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$out/${tag}_bench_reference.json" 2> "$out/${tag}_bench_reference.err"
echo "reference arm rc=$?" | tee -a "$out/${tag}_summary.txt"
echo "== ncu launch list (2+1 steps, never a bench value)" | tee -a "$out/${tag}_summary.txt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$out/${tag}_launches.csv" \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-also --no-decks --use-graph 0 > "$out/${tag}_launches_run.log" 2>&1
echo "ncu rc=$?" | tee -a "$out/${tag}_summary.txt"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> "$out/${tag}_summary.txt" 2>&1
cat "$out/${tag}_summary.txt"
