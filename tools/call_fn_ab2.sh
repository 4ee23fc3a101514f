#!/bin/bash
python -m pytest tests/test_gpu_fn.py tests/test_gpu_fn_kats.py -x -q 2>&1 | tail -2
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_keep.so
for i in 1 2; do for lib in "$@"; do
  cp $lib $L; touch $L
  python bench_suite.py --only c6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', d['envs'], round(d['ms'] * 1e3, 1), 'us', round(d['env_steps_per_s'] / 1e9, 3), 'G', round(d['frac_of_hbm_peak'], 3))"
done; done
cp /tmp/_keep.so $L
