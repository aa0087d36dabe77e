#!/usr/bin/env bash
set -u
tag=${1:-r02u}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_stl2poro.py -m gpu -q -s > $out/${tag}_pytest_stl.log 2>&1
echo "pytest stl rc=$? : $(tail -1 $out/${tag}_pytest_stl.log)" | tee $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)|dragon 256|disagree" $out/${tag}_pytest_stl.log | head | tee -a $out/${tag}_summary.txt
timeout 300 python tools/decks_probe.py --sor-variant 0 2>&1 | cut -c1-300 | tee -a $out/${tag}_summary.txt
timeout 600 python bench.py --workload dragon_stl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-decks > $out/${tag}_bench_dragon_stl.json 2> $out/${tag}_bench_dragon_stl.err
echo "bench dragon_stl rc=$?" | tee -a $out/${tag}_summary.txt
timeout 600 python bench.py --workload dragon --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-also --no-decks > $out/${tag}_bench_dragon.json 2> $out/${tag}_bench_dragon.err
python - $out/${tag}_bench_dragon_stl.json $out/${tag}_bench_dragon.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("  ", d["config"]["workload"], f"{d['ms_per_step']:.2f} ms/step, {d['sor_sweeps_per_s']:.0f} sweeps/s", d.get("setup"), d.get("parity", {}).get("crosscheck", {}).get("fields_identical"))
    except Exception as e:
        print("   no line", f, e)
PY
tail -3 $out/${tag}_bench_dragon_stl.err
