#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02m}
mkdir -p $out
L="python tools/sor_lab.py"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
for rep in 1 2; do
timeout 300 $L --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 300 $L --variant 6 --grid 256 256 256 --check --lib exp/libpf_v6committed.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
done
timeout 300 $L --variant 6 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 300 $L --variant 6 --grid 1024 512 512 --lib exp/libpf_v6committed.so --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
