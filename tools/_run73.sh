cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -5
for v in 1 0; do
  if [ $v = 1 ]; then export TG_GFEATS_V1=1; else unset TG_GFEATS_V1; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('V1=$v', d['extra']['grouped'])"
done
unset TG_GFEATS_V1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grouped2.csv python tools/prof_grouped.py > gpurun_out/ncu_grouped2.log 2>&1
grep -E "k_grouped|k_step" gpurun_out/launches_grouped2.csv | awk -F'","' '{print $5, $(NF)}' | sort | uniq -c | sort -rn | head -8
