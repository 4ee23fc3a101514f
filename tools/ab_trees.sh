#!/bin/bash
# bench the headline step of several exported commit trees (tools/_ab/tree_<sha>, each with its own library and bench.py) on ONE box
RUNS=$1; shift
for i in $(seq $RUNS); do
  for t in "$@"; do
    (cd $t && python bench.py --steps 60 --warmup 5 --steady 256 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('$t', round(d['value'] / 1e9, 4), 'G env-steps/s', round(d['roofline']['frac'], 4))")
  done
done
