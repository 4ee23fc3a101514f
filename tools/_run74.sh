cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_grouped.py -x -q -m gpu 2>&1 | tail -3
g() { timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['extra']['grouped']['placements_per_s']/1e9, d['extra']['rollout']['placements_per_s']/1e9)"; }
g base
TG_NL=6 TG_NF=2 g nl6nf2
TG_NL=5 TG_NF=2 g nl5nf2
TG_NL=5 TG_NF=3 g nl5nf3
TG_NL=3 TG_NF=2 g nl3nf2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grouped_feats_x -s 8 -c 1 -o gpurun_out/prof_gfeats_x python tools/prof_paths.py feats --envs 1048576 > gpurun_out/ncu30.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 8 -c 1 -o gpurun_out/prof_step_mode2 python tools/prof_paths.py feats --envs 1048576 > gpurun_out/ncu31.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
