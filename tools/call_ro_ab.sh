#!/bin/bash
# rollout parity + A/B of several libs on bench_suite C4 (fused rollout)
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_cross_kernel.py -x -q 2>&1 | tail -3
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_keep.so
for i in 1 2; do for lib in "$@"; do
  cp $lib $L; touch $L
  python bench_suite.py --only c4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', d['envs'], round(d.get('ms', d.get('ms_per_launch', 0)), 2), 'ms', round(d['placements_per_s'] / 1e9, 2), 'G placements/s')"
done; done
cp /tmp/_keep.so $L
