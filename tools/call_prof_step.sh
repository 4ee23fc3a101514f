#!/bin/bash
# ncu capture (with source) of the headline step kernel at the bench workload + the launch list of the same command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 70 -c 1 -f -o gpurun_out/r02g_step python bench.py --steps 10 --warmup 5 --steady 64 --no-e2e --no-cpu-baseline --no-extra --no-probe > gpurun_out/r02g_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_step.csv python bench.py --steps 10 --warmup 5 --steady 64 --no-e2e --no-cpu-baseline --no-extra --no-probe > gpurun_out/r02g_launch.log 2>&1
ls -la gpurun_out | grep r02g
