cd $GRAFT_REPO_ROOT
g() { timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9, d['value']/1e9)"; }
timeout 900 python -m pytest tests/test_gpu_rollout.py -q -m gpu 2>&1 | tail -3
g rollout_x_6cta
