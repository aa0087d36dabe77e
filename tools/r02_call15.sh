#!/usr/bin/env bash
# 2 GPUs: persistent half-sweeps (variant 7, hand-rolled grid barrier) parity + deck timings; drivers on 2 GPUs; N=2 bench
set -u
tag=${1:-r02r}; n=2
out=gpurun_out; mkdir -p $out
echo "== variant 7 parity" | tee $out/${tag}_summary.txt
PF_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zzzz_experimental.py -m gpu -q > $out/${tag}_pytest_v7.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest_v7.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest_v7.log | head | tee -a $out/${tag}_summary.txt
timeout 300 python tools/decks_probe.py --sor-variant 0 > $out/${tag}_decks_v0.jsonl 2> $out/${tag}_decks_v0.err
timeout 300 python tools/decks_probe.py --sor-variant 7 > $out/${tag}_decks_v7.jsonl 2> $out/${tag}_decks_v7.err
cat $out/${tag}_decks_v0.jsonl $out/${tag}_decks_v7.jsonl | cut -c1-330 | tee -a $out/${tag}_summary.txt
tail -3 $out/${tag}_decks_v7.err | tee -a $out/${tag}_summary.txt
echo "== multi_gpu_check world=2" | tee -a $out/${tag}_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_n2.log 2>&1
echo "multi_gpu_check rc=$?" | tee -a $out/${tag}_summary.txt
grep -E "^(FAIL|multi_gpu_check: world)" $out/${tag}_multi_gpu_check_n2.log | cut -c1-300 | tee -a $out/${tag}_summary.txt
echo "== pytest: drivers on several GPUs" | tee -a $out/${tag}_summary.txt
timeout 900 python -m pytest tests/test_gpu_zz_driver_rundirs.py tests/test_gpu_zzz_fortran_driver.py tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest_drivers.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest_drivers.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest_drivers.log | head | tee -a $out/${tag}_summary.txt
for ht in 0 3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$ht bench.py --gpus 2 --steps 5 --warmup 3 --halo-transport $ht --no-cpu-baseline --no-e2e > $out/${tag}_bench_n2_ht$ht.json 2> $out/${tag}_bench_n2_ht$ht.err
  echo "bench ht=$ht rc=$?" | tee -a $out/${tag}_summary.txt
  python - $out/${tag}_bench_n2_ht$ht.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("  ", d["n_gpus"], "GPUs", d["config"].get("halo_transport"), f"{d['ms_per_step']:.2f} ms/step, SOR {d['ms_sor_per_step']:.2f} ms,",
          f"{d['value']/1e6:.1f} M, launches", d.get("gpu_launches"), "parity", d.get("parity", {}).get("fields_sha256", "")[:16])
except Exception as e:
    print("   no line:", e)
PY
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-decks > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python - $out/${tag}_bench_n1.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("   1 GPU", f"{d['ms_per_step']:.2f} ms/step, SOR {d['ms_sor_per_step']:.2f} ms, also", d.get("also", {}).get("ms_per_step"), d.get("also", {}).get("sor_sweeps_per_s"), "parity", d.get("parity", {}).get("fields_sha256", "")[:16])
PY
