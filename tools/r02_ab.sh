#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02ab}
L="python tools/sor_lab.py"
for rep in 1 2; do
timeout 300 $L --variant 6 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
PF_TMA_CHUNK=256 timeout 300 $L --variant 6 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
done
PF_TMA_CHUNK=128 timeout 300 $L --variant 6 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
cat $out/${tag}_lab.jsonl | cut -c1-200
