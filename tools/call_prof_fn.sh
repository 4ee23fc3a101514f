#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fn_step_tile -s 6 -c 1 -f -o gpurun_out/r02c_fn python tools/prof_paths.py fn --envs 1048576 > gpurun_out/r02c_fn.log 2>&1
tail -2 gpurun_out/r02c_fn.log
