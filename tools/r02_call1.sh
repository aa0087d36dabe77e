#!/usr/bin/env bash
# round 2, GPU call 1: everything that has never run on a GPU + the first numbers of SOR variant 8
set -u
out=gpurun_out; tag=r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
echo "== lab (parity vs variant 1 by hash, then timing)" | tee $out/${tag}_summary.txt
for v in 1 6 8; do timeout 300 python tools/sor_lab.py --variant $v --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
for lib in nocompute noload minb1; do timeout 300 python tools/sor_lab.py --variant 8 --grid 256 256 256 --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
for v in 6 8; do timeout 400 python tools/sor_lab.py --variant $v --grid 1024 512 512 --check --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
for lib in nocompute noload; do timeout 400 python tools/sor_lab.py --variant 8 --grid 1024 512 512 --lib exp/libpf_$lib.so --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
cat $out/${tag}_lab.jsonl | tee -a $out/${tag}_summary.txt
tail -5 $out/${tag}_lab.err | tee -a $out/${tag}_summary.txt
echo "== pytest -m gpu (experimental included)" | tee -a $out/${tag}_summary.txt
PF_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head -40 | tee -a $out/${tag}_summary.txt
echo "== decks probe: default and variant 7" | tee -a $out/${tag}_summary.txt
timeout 300 python tools/decks_probe.py --sor-variant 0 > $out/${tag}_decks_v0.jsonl 2> $out/${tag}_decks_v0.err
timeout 300 python tools/decks_probe.py --sor-variant 7 > $out/${tag}_decks_v7.jsonl 2> $out/${tag}_decks_v7.err
cat $out/${tag}_decks_v0.jsonl $out/${tag}_decks_v7.jsonl | tee -a $out/${tag}_summary.txt
tail -3 $out/${tag}_decks_v7.err | tee -a $out/${tag}_summary.txt
echo "== compute-sanitizer memcheck (small cases, every kernel)" | tee -a $out/${tag}_summary.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_small.py > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$? : $(tail -2 $out/${tag}_memcheck.log | tr '\n' ' ')" | tee -a $out/${tag}_summary.txt
