#!/bin/bash
# compute-sanitizer pass over every kernel (tools/sanitize_driver.py); writes gpurun_out/sanitize_<tool>.log
# usage (GPU box): bash tools/sanitize.sh [tools...]      default: memcheck racecheck synccheck initcheck
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck initcheck}
for t in $TOOLS; do
  extra=""
  [ "$t" = "memcheck" ] && extra="--leak-check no"
  timeout 900 compute-sanitizer --tool $t $extra --error-exitcode 9 --print-limit 30 \
      python tools/sanitize_driver.py > gpurun_out/sanitize_$t.log 2>&1
  echo "$t rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$t.log | tail -1)"
done
