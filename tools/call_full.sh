#!/bin/bash
# full GPU suite + bench (both arms) + ncu captures of the kernels changed in this session
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02d_gputests.log 2>&1; tail -3 gpurun_out/r02d_gputests.log
( time python bench.py ) > gpurun_out/r02d_bench.log 2> gpurun_out/r02d_bench.err; tail -c 600 gpurun_out/r02d_bench.log; tail -3 gpurun_out/r02d_bench.err
TG_GROUPED_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grouped_feats_x -s 45 -c 1 -f -o gpurun_out/r02d_gfeats python tools/time_grouped.py > gpurun_out/r02d_gfeats.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fn_step_tile -s 6 -c 1 -f -o gpurun_out/r02d_fn python tools/prof_paths.py fn --envs 1048576 > gpurun_out/r02d_fn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_x -s 2 -c 1 -f -o gpurun_out/r02d_rollout python tools/prof_paths.py rollout --envs 1048576 > gpurun_out/r02d_rollout.log 2>&1
ls -la gpurun_out | tail -12
