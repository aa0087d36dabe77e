#!/usr/bin/env bash
# round 2, 8-GPU call: z-slab parity against the single-domain oracle on 8 ranks, a driver on 4 GPUs, strong scaling
# 1 / 4 / 8 GPUs on ONE box with the same steps (field hashes must agree)
set -u
tag=${1:-r02q}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== multi_gpu_check world=8" | tee $out/${tag}_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_n8.log 2>&1
echo "multi_gpu_check rc=$?" | tee -a $out/${tag}_summary.txt
grep -E "^(FAIL|multi_gpu_check: world)" $out/${tag}_multi_gpu_check_n8.log | cut -c1-400 | tee -a $out/${tag}_summary.txt
echo "== pytest: drivers on 2 and 4 GPUs" | tee -a $out/${tag}_summary.txt
timeout 900 python -m pytest tests/test_gpu_zz_driver_rundirs.py tests/test_gpu_zzz_fortran_driver.py -m gpu -q -k "2 or 4" > $out/${tag}_pytest_drivers.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest_drivers.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest_drivers.log | head | tee -a $out/${tag}_summary.txt
B="bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also --no-decks"
run() { # n ht extra
  local n=$1 ht=$2; shift 2
  if [ $n = 1 ]; then timeout 900 python $B --gpus 1 "$@" > $out/${tag}_bench_n${n}_ht$ht.json 2> $out/${tag}_bench_n${n}_ht$ht.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 297$n$ht $B --gpus $n --halo-transport $ht "$@" > $out/${tag}_bench_n${n}_ht$ht.json 2> $out/${tag}_bench_n${n}_ht$ht.err; fi
  echo "bench N=$n ht=$ht rc=$?" | tee -a $out/${tag}_summary.txt
  python - $out/${tag}_bench_n${n}_ht$ht.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("  ", d["n_gpus"], "GPUs |", d["config"].get("halo_transport"), f"| {d['ms_per_step']:.2f} ms/step, SOR {d['ms_sor_per_step']:.2f} ms,",
          f"{d['value']/1e6:.1f} M, e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "| fields", d.get("parity", {}).get("fields_sha256", "")[:16],
          "p_error", d.get("parity", {}).get("p_error_sha256", "")[:16], "| clocks", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("   no line:", e)
PY
}
run 8 0
run 8 3 --no-e2e
run 4 0 --no-e2e
run 1 0 --no-e2e
