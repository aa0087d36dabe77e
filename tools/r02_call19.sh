#!/usr/bin/env bash
set -u
tag=${1:-r02v}
out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_n2.log 2>&1
echo "multi_gpu_check rc=$?" | tee $out/${tag}_summary.txt
grep -E "^(FAIL|multi_gpu_check: world)" $out/${tag}_multi_gpu_check_n2.log | cut -c1-300 | tee -a $out/${tag}_summary.txt
tail -5 $out/${tag}_multi_gpu_check_n2.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head -20 | tee -a $out/${tag}_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python - $out/${tag}_bench_n2.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("  ", d["n_gpus"], "GPUs", f"{d['ms_per_step']:.2f} ms/step, e2e", d["e2e"]["ms_per_step"], d["e2e"].get("host_numa_binding"), d["e2e"]["pcie_gbs_per_gpu_if_serial"])
PY
