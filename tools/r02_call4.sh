#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02d}
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "branch_free or (uniform_steps and 8)" 2>&1 | tail -3
timeout 300 python tools/sor_lab.py --variant 8 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 400 python tools/sor_lab.py --variant 8 --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
bash tools/r02_ncu.sh ${tag}_v8_256 8 sor_tma2 - 256 256 256
