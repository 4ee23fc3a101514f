#!/bin/bash
mkdir -p gpurun_out
TG_GROUPED_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grouped_feats_x -s 45 -c 1 -f -o gpurun_out/r02c_gfeats python tools/time_grouped.py > gpurun_out/r02c_gfeats.log 2>&1
tail -2 gpurun_out/r02c_gfeats.log
