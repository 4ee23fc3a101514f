cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_x -s 2 -c 1 -o gpurun_out/prof_rollout_x python tools/prof_paths.py rollout --envs 524288 > gpurun_out/ncu32.log 2>&1
ls -la gpurun_out/prof_rollout_x.ncu-rep
