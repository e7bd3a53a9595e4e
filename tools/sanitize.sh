#!/bin/bash
# compute-sanitizer over every kernel of libfutbol_b200.so ("race detection / sanitizers").
# Run on a GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2'
# Writes gpurun_out/<tag>_sanitize_{memcheck,synccheck,racecheck,initcheck}.log and a one-line-per-tool summary.
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python tools/sanitize_exercise.py > $out/${tag}_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 $out/${tag}_sanitize_plain.log; exit 1; }
summary=$out/${tag}_sanitize_summary.txt
: > $summary
echo "compute-sanitizer $(compute-sanitizer --version | tail -1); $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader)" >> $summary
echo "workload: tools/sanitize_exercise.py  ($(tail -1 $out/${tag}_sanitize_plain.log))" >> $summary
for tool in memcheck synccheck racecheck initcheck; do
    log=$out/${tag}_sanitize_${tool}.log
    extra=""
    [ $tool = memcheck ] && extra="--leak-check no"
    [ $tool = racecheck ] && extra="--racecheck-report all"
    timeout 420 compute-sanitizer --tool $tool $extra --print-limit 50 --target-processes all \
        python tools/sanitize_exercise.py > $log 2>&1
    rc=$?
    echo "$tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1); script: $(grep -c 'sanitize_exercise ok' $log) ok line(s)" >> $summary
done
cat $summary
