cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_v9.json
python -c "
import json; d=json.load(open('gpurun_out/bench_v9.json')); print(d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e6, d['e2e']['obs_on_device']['value']/1e9, d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9, d['cpu_baseline']['value']/1e6, d['clocks'])"
# ncu: steady-state step kernel at the bench workload (full set, one launch) + launch list of the bench command
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 33 -c 1 -o gpurun_out/prof_step_v6 python bench.py --steps 4 --warmup 30 --no-e2e --no-cpu-baseline --no-extra --no-probe > gpurun_out/ncu33.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu34.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grouped_feats_x -s 8 -c 1 -o gpurun_out/prof_gfeats_x2 python tools/prof_paths.py feats --envs 1048576 > gpurun_out/ncu35.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 8 -c 1 -o gpurun_out/prof_step_mode2_v2 python tools/prof_paths.py feats --envs 1048576 > gpurun_out/ncu36.log 2>&1
timeout 900 python bench_suite.py --out gpurun_out/suite_v5 > gpurun_out/suite_v5.log 2>&1
tail -30 gpurun_out/suite_v5.md
