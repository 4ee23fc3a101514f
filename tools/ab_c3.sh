#!/bin/bash
# grouped + features (C3, 1 M envs) with several builds of libtetris_b200.so on ONE box:  bash tools/ab_c3.sh <runs> <lib1.so> ...
RUNS=$1; shift
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_ab_keep.so
for i in $(seq $RUNS); do
  for lib in "$@"; do
    cp "$lib" $L; touch $L
    python bench_suite.py --only c3 2>/dev/null | grep 'C3 grouped' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', d['envs'], round(d['ms'] * 1e3, 1), 'us', round(d['placements_per_s'] / 1e9, 2), 'G placements/s')"
  done
done
cp /tmp/_ab_keep.so $L; touch $L
