#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02g}
mkdir -p $out
for lib in v6nocompute v6noload; do
  timeout 300 python tools/sor_lab.py --variant 6 --grid 256 256 256 --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
  timeout 400 python tools/sor_lab.py --variant 6 --grid 1024 512 512 --lib exp/libpf_$lib.so --steps 2 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
done
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
