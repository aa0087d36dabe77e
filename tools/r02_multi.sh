#!/usr/bin/env bash
# round 2, multi-GPU call: z-slab parity against the single-domain oracle, the drivers on N GPUs, the bench at N
# usage: gpurun --gpus N -- 'bash tools/r02_multi.sh TAG N'
set -u
tag=${1:-r02m}; n=${2:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
echo "== multi_gpu_check world=$n" | tee $out/${tag}_summary.txt
timeout 900 $TR --master-port 29611 tests/multi_gpu_check.py > $out/${tag}_multi_gpu_check_n$n.log 2>&1
echo "multi_gpu_check rc=$?" | tee -a $out/${tag}_summary.txt
grep -E "^(FAIL|multi_gpu_check)" $out/${tag}_multi_gpu_check_n$n.log | cut -c1-600 | tee -a $out/${tag}_summary.txt
echo "== pytest: drivers on several GPUs" | tee -a $out/${tag}_summary.txt
timeout 900 python -m pytest tests/test_gpu_zz_driver_rundirs.py tests/test_gpu_zzz_fortran_driver.py -m gpu -q -k "2 or 4" > $out/${tag}_pytest_drivers.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest_drivers.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest_drivers.log | head | tee -a $out/${tag}_summary.txt
echo "== bench at N=$n: auto transport (in-kernel handshake), then the barrier kernel" | tee -a $out/${tag}_summary.txt
for ht in 0 3; do
  timeout 900 $TR --master-port 2962$ht bench.py --gpus $n --steps 5 --warmup 3 --halo-transport $ht --no-cpu-baseline $( [ $ht = 3 ] && echo --no-e2e ) \
     > $out/${tag}_bench_n${n}_ht$ht.json 2> $out/${tag}_bench_n${n}_ht$ht.err
  echo "bench ht=$ht rc=$?" | tee -a $out/${tag}_summary.txt
  python - $out/${tag}_bench_n${n}_ht$ht.json <<'PY' | tee -a $out/${tag}_summary.txt
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("  ", d["n_gpus"], "GPUs", d["config"].get("halo_transport"), f"{d['ms_per_step']:.2f} ms/step, SOR {d['ms_sor_per_step']:.2f} ms,",
          f"{d['value']/1e6:.1f} M, e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("parity", {}).get("fields_sha256", "")[:16],
          d.get("parity", {}).get("p_error_sha256", "")[:16])
except Exception as e:
    print("   no line:", e)
PY
done
tail -3 $out/${tag}_bench_n${n}_ht0.err
