#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02sch}
mkdir -p $out
L="python tools/sor_lab.py"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decks.py tests/test_gpu_long.py tests/test_gpu_z_ref_golden.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head
for rep in 1 2; do
timeout 300 $L --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
PF_TMA_CHUNK=86 timeout 300 $L --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
done
timeout 300 $L --variant 6 --grid 1024 512 64 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
PF_TMA_CHUNK=32 timeout 300 $L --variant 6 --grid 1024 512 64 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 300 $L --variant 6 --grid 1024 512 128 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
PF_TMA_CHUNK=64 timeout 300 $L --variant 6 --grid 1024 512 128 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 300 $L --variant 6 --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 300 $L --variant 1 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
cat $out/${tag}_lab.jsonl | cut -c1-250; tail -5 $out/${tag}_lab.err
