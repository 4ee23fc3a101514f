#!/bin/bash
# A/B of the functional tile kernel's envs-per-tile (TG_FN_E) + source-level ncu captures of the grouped path (steady state)
mkdir -p gpurun_out
for e in 16 32 8; do
  echo "== TG_FN_E=$e tests"; TG_FN_E=$e python -m pytest tests/test_gpu_fn.py tests/test_gpu_fn_kats.py -x -q 2>&1 | tail -2
done
for i in 1 2; do for e in 16 32 8; do
  echo "== TG_FN_E=$e"; TG_FN_E=$e python bench_suite.py --only c6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['envs'], round(d['ms'] * 1e3, 1), 'us', round(d['env_steps_per_s'] / 1e9, 3), 'G', round(d['frac_of_hbm_peak'], 3))"
done; done
TG_GROUPED_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_ws -s 45 -c 1 -f -o gpurun_out/r02b_gstep python tools/time_grouped.py > gpurun_out/r02b_gstep.log 2>&1
ncu -i gpurun_out/r02b_gstep.ncu-rep --page source --csv > gpurun_out/r02b_gstep_src.csv 2>/dev/null
TG_GROUPED_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grouped_feats_x -s 45 -c 1 -f -o gpurun_out/r02b_gfeats python tools/time_grouped.py > gpurun_out/r02b_gfeats.log 2>&1
ncu -i gpurun_out/r02b_gfeats.ncu-rep --page source --csv > gpurun_out/r02b_gfeats_src.csv 2>/dev/null
ls -la gpurun_out
