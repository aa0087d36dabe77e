#!/usr/bin/env bash
# variant 8 (2D temporally blocked SOR): parity, then tile-shape sweep on the decks
set -u
tag=${1:-r02s}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzzz_deck_kernels.py -m gpu -q -k "ibm2 or deck" > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)" $out/${tag}_pytest.log | head -20 | tee -a $out/${tag}_summary.txt
probe() { timeout 300 python tools/decks_probe.py --sor-variant $1 2>> $out/${tag}_probe.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', d['deck'], 'variant', d['sor_variant'], 'identical', d['bit_identical_to_reference_after_3_steps'], 'ms/step %.3f' % d['ms_per_step'], 'sor %.3f' % d['ms_sor_per_step'])
"; }
echo "== variant 1 (half-sweep launches)" | tee -a $out/${tag}_summary.txt
probe 1 | tee -a $out/${tag}_summary.txt
for cfg in "4 64 32" "3 64 32" "2 64 32" "4 128 16" "3 128 16" "2 128 24" "4 48 48" "3 56 40" "5 64 24" "6 64 16" "2 32 32" "1 64 48"; do
  set -- $cfg
  echo "== variant 8  T=$1 owned $2 x $3" | tee -a $out/${tag}_summary.txt
  PF_TB_T=$1 PF_TB_OW=$2 PF_TB_OH=$3 probe 8 | tee -a $out/${tag}_summary.txt
done
tail -5 $out/${tag}_probe.err
