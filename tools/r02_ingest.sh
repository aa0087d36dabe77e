#!/usr/bin/env bash
set -u
out=gpurun_out; tag=${1:-r02ing}
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_zz_driver_rundirs.py tests/test_gpu_decks.py -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$? : $(tail -1 $out/${tag}_pytest.log)" | tee $out/${tag}_summary.txt
grep -E "^(FAILED|ERROR)|^E  " $out/${tag}_pytest.log | head -20 | tee -a $out/${tag}_summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_small.py > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$? : $(tail -2 $out/${tag}_memcheck.log | tr '\n' ' ')" | tee -a $out/${tag}_summary.txt
grep -c "Invalid\|out of bounds" $out/${tag}_memcheck.log | tee -a $out/${tag}_summary.txt
