#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02i}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_ref_golden.py tests/test_gpu_long.py -m gpu -q -k "6 or long or every_sor" 2>&1 | tail -3
timeout 300 python tools/sor_lab.py --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6old v6nofence; do timeout 300 python tools/sor_lab.py --variant 6 --grid 256 256 256 --check --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
timeout 400 python tools/sor_lab.py --variant 6 --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6old v6nofence; do timeout 400 python tools/sor_lab.py --variant 6 --grid 1024 512 512 --check --lib exp/libpf_$lib.so --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
timeout 900 python bench.py --steps 3 --warmup 1 --sor-variant 6 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
