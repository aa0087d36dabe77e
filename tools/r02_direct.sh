#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02dir}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decks.py tests/test_gpu_long.py tests/test_gpu_z_ref_golden.py tests/test_gpu_voxel2poro.py tests/test_gpu_zz_driver_rundirs.py -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-decks > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - $out/${tag}_bench.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("  ", f"{d['ms_per_step']:.2f} ms/step, SOR {d['ms_sor_per_step']:.2f}", d["parity"]["fields_sha256"][:16], d["parity"]["crosscheck"]["fields_identical"], "also", d["also"]["ms_per_step"], d["also"]["parity"]["fields_sha256"][:16])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-also --no-decks --no-parity --use-graph 0 > $out/${tag}_launches_run.log 2>&1
echo "ncu launch list rc=$?"
