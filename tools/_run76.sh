cd $GRAFT_REPO_ROOT
g() { timeout 300 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9, d['value']/1e9)"; }
timeout 600 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_base.py -x -q -m gpu 2>&1 | tail -3
g whole16
TG_WHOLE=99 g whole_off
TG_WHOLE=8 g whole8
TG_WHOLE=12 g whole12
TG_WHOLE=99 TG_L2HINT=5 g off_hint5
TG_WHOLE=99 TG_L2HINT=0 g off_hint0
TG_WHOLE=99 TG_L2HINT=1 g off_hint1
TG_WHOLE=99 TG_L2HINT=5 g off_hint5
TG_WHOLE=99 TG_L2HINT=0 g off_hint0
TG_WHOLE=99 TG_L2HINT=1 g off_hint1
