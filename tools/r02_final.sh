#!/usr/bin/env bash
# end-of-round validation on one GPU: GPU tests, smoke, default bench line, reference arm, ncu launch list + full captures
set -u
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv | tee $out/${tag}_summary.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee -a $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head -20 | tee -a $out/${tag}_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$? : $(tail -1 $out/${tag}_smoke.log)" | tee -a $out/${tag}_summary.txt
timeout 1200 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?" | tee -a $out/${tag}_summary.txt
python tools/summarize.py $out/${tag}_bench.json 2>&1 | cut -c1-600 | tee -a $out/${tag}_summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference arm rc=$? $(cut -c1-300 $out/${tag}_bench_reference.json)" | tee -a $out/${tag}_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-also --no-decks --no-parity --use-graph 0 > $out/${tag}_launches_run.log 2>&1
echo "ncu launch list rc=$?" | tee -a $out/${tag}_summary.txt
bash tools/r02_ncu.sh ${tag}_v6_256 6 sor_tma_kernel - 256 256 256 | tail -2
bash tools/r02_ncu.sh ${tag}_v6_s1 6 sor_tma_kernel - 1024 512 512 | tail -2
