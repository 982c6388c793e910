#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck) and the smoke sweep (racecheck, initcheck).
# Run on the GPU box:  bash scripts/sanitize.sh [memcheck-seconds] [other-seconds]   -> gpurun_out/sanitize_*.log
# A tool that does not finish inside its time limit is reported as such; nothing here is a timing run.
T1=${1:-480}
T2=${2:-200}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
SMOKE='import __graft_entry__ as g; g.smoke()'
echo "== memcheck: tests/test_gpu_parity.py" > $OUT/sanitize_summary.txt
timeout $T1 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitize_memcheck.log \
  python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider > $OUT/sanitize_memcheck_pytest.log 2>&1
echo "exit $? (124 = time limit, 9 = sanitizer errors)" >> $OUT/sanitize_summary.txt
tail -3 $OUT/sanitize_memcheck_pytest.log >> $OUT/sanitize_summary.txt
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" $OUT/sanitize_memcheck.log | sort | uniq -c | head -20 >> $OUT/sanitize_summary.txt
for tool in racecheck initcheck; do
  echo "== $tool: smoke()" >> $OUT/sanitize_summary.txt
  timeout $T2 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/sanitize_$tool.log python -c "$SMOKE" > $OUT/sanitize_${tool}_run.log 2>&1
  echo "exit $? (124 = time limit, 9 = sanitizer errors)" >> $OUT/sanitize_summary.txt
  tail -2 $OUT/sanitize_${tool}_run.log >> $OUT/sanitize_summary.txt
  grep -E "SUMMARY|hazard|Uninitialized" $OUT/sanitize_$tool.log | sort | uniq -c | head -20 >> $OUT/sanitize_summary.txt
done
cat $OUT/sanitize_summary.txt
