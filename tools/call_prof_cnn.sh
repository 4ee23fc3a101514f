#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cnn_obs3 -s 6 -c 1 -f -o gpurun_out/r02d_cnn python tools/prof_paths.py cnn --envs 65536 > gpurun_out/r02d_cnn.log 2>&1
tail -2 gpurun_out/r02d_cnn.log
