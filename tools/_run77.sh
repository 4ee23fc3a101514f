cd $GRAFT_REPO_ROOT
g() { timeout 300 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9, d['value']/1e9)"; }
timeout 600 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_base.py -x -q -m gpu 2>&1 | tail -3
g info_in_feats
TG_INFO_IN_STEP=1 g info_in_step
g info_in_feats
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grouped3.csv python tools/prof_grouped.py > gpurun_out/ncu_grouped3.log 2>&1
grep -E "k_grouped|k_step" gpurun_out/launches_grouped3.csv | awk -F'","' '{print substr($5,1,40), $(NF)}' | tail -8
