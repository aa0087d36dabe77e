#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02j}
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head
timeout 300 python tools/sor_lab.py --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6nocompute v6noload; do timeout 300 python tools/sor_lab.py --variant 6 --grid 256 256 256 --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
timeout 400 python tools/sor_lab.py --variant 6 --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6nocompute v6noload; do timeout 400 python tools/sor_lab.py --variant 6 --grid 1024 512 512 --lib exp/libpf_$lib.so --steps 2 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
bash tools/r02_ncu.sh ${tag}_v6_256 6 sor_tma_kernel - 256 256 256
