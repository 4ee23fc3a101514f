#!/bin/bash
python -m pytest tests/test_gpu_cnn_obs.py -x -q 2>&1 | tail -3
TG_CNN_V2=1 python -m pytest tests/test_gpu_cnn_obs.py -x -q 2>&1 | tail -1
for i in 1 2; do for v in 0 1; do
  echo "== TG_CNN_V2=$v"; ( if [ $v = 1 ]; then export TG_CNN_V2=1; fi; python bench_suite.py --only c5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'CNN' in d['config'] or 'cnn' in d['config']: print(d['config'][:60], d['envs'], round(d['ms'], 3), 'ms', round(d['env_steps_per_s'] / 1e6, 1), 'M env-steps/s')" )
done; done
