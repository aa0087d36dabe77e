#!/usr/bin/env bash
# programmatic dependent launch of the half-sweep chains: parity, then deck timings with and without it
set -u
tag=${1:-r02t}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head -20 | tee -a $out/${tag}_summary.txt
probe() { timeout 300 python tools/decks_probe.py --sor-variant $1 2>> $out/${tag}_probe.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', d['deck'], 'variant', d['sor_variant'], 'identical', d['bit_identical_to_reference_after_3_steps'], 'ms/step %.3f' % d['ms_per_step'], 'sor %.3f' % d['ms_sor_per_step'])
"; }
for rep in 1 2; do
echo "== variant 1, PF_PDL=0" | tee -a $out/${tag}_summary.txt
PF_PDL=0 probe 1 | tee -a $out/${tag}_summary.txt
echo "== variant 1, programmatic dependent launches" | tee -a $out/${tag}_summary.txt
probe 1 | tee -a $out/${tag}_summary.txt
done
tail -5 $out/${tag}_probe.err
