cd $GRAFT_REPO_ROOT
g() { timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9, d['value']/1e9)"; }
timeout 600 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_base.py -x -q -m gpu 2>&1 | tail -3
g regs40
g regs40_again
TG_NSX=3 g nsx3
TG_NSX=4 g nsx4
TG_NO_CARVEOUT=1 g nocarve
TG_NVCC_FLAGS="-DGF_REGS=48" python -m tetris_gymnasium_b200._build > /dev/null 2>&1
g regs48
TG_NVCC_FLAGS="-DGF_REGS=56" python -m tetris_gymnasium_b200._build > /dev/null 2>&1
g regs56
