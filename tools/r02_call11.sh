#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02l}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head
L="python tools/sor_lab.py"
timeout 300 $L --variant 6 --grid 256 256 256 --check --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6nocompute v6noload; do timeout 300 $L --variant 6 --grid 256 256 256 --lib exp/libpf_$lib.so --steps 3 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
timeout 400 $L --variant 6 --grid 1024 512 512 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
for lib in v6nocompute v6noload; do timeout 400 $L --variant 6 --grid 1024 512 512 --lib exp/libpf_$lib.so --steps 2 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err; done
# the shape of one rank of the 2-GPU run, on one GPU (periodic in z, no neighbour): is 1.72 ms/iteration the shape or the slab path?
timeout 400 $L --variant 6 --grid 1024 512 256 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 400 $L --variant 6 --grid 1024 512 128 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
timeout 400 $L --variant 6 --grid 1024 512 64 --steps 3 --warmup 1 >> $out/${tag}_lab.jsonl 2>> $out/${tag}_lab.err
cat $out/${tag}_lab.jsonl; tail -5 $out/${tag}_lab.err
timeout 600 python tools/chunk_sweep.py 1024 512 256 256 128 86 64 > $out/${tag}_chunk_1024x512x256.txt 2>&1; cat $out/${tag}_chunk_1024x512x256.txt
bash tools/r02_ncu.sh ${tag}_v6_256 6 sor_tma_kernel - 256 256 256
bash tools/r02_ncu.sh ${tag}_v6_s1 6 sor_tma_kernel - 1024 512 512
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1500 $out/${tag}_bench.json | head -c 600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-also --no-decks --no-parity --use-graph 0 > $out/${tag}_launches_run.log 2>&1
echo "ncu launches rc=$?"
