#!/bin/bash
# A/B/C... several builds of libtetris_b200.so on ONE box, alternating runs:  bash tools/ab_multi.sh <runs> <lib1.so> <lib2.so> ...
set -e
RUNS=$1; shift
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_ab_keep.so
for i in $(seq $RUNS); do
  for lib in "$@"; do
    cp "$lib" $L; touch $L
    python bench.py --steps 60 --warmup 5 --steady 256 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('$lib', round(d['value'] / 1e9, 4), 'G env-steps/s', round(d['roofline']['frac'], 4))"
  done
done
cp /tmp/_ab_keep.so $L; touch $L
