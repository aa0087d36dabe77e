#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02h}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_ref_golden.py -m gpu -q -k "9" 2>&1 | tail -3
for v in 9; do timeout 300 python tools/sor_lab.py --variant $v --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
for lib in v9nocompute v9ng5; do timeout 300 python tools/sor_lab.py --variant 9 --grid 256 256 256 --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
timeout 400 python tools/sor_lab.py --variant 9 --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v9nocompute v9ng5; do timeout 400 python tools/sor_lab.py --variant 9 --grid 1024 512 512 --lib exp/libpf_$lib.so --steps 2 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
bash tools/r02_ncu.sh ${tag}_v9_256 9 sor_tma3 - 256 256 256
