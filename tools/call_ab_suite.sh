#!/bin/bash
# A/B of several libs on one bench_suite leg:  bash tools/call_ab_suite.sh <leg> <grep pattern> <lib1.so> <lib2.so> ...
LEG=$1; PAT=$2; shift 2
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_keep.so
for i in 1 2; do for lib in "$@"; do
  cp $lib $L; touch $L
  python bench_suite.py --only $LEG 2>/dev/null | grep "$PAT" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', d['envs'], round(d.get('ms', d.get('ms_per_launch', 0)), 4), 'ms', round(d['env_steps_per_s'] / 1e6, 1), 'M env-steps/s')"
done; done
cp /tmp/_keep.so $L
