#!/bin/bash
# A/B two builds of libtetris_b200.so on ONE box, alternating runs (the step kernel reacts to code-footprint changes of a few
# hundred instructions by +-10 %, so every change to csrc/tg_step.cuh / tg_device.cuh gets this check).
#   bash tools/ab_lib.sh <libA.so> <libB.so> [runs=3] [bench args...]
# e.g. build the previous commit into tools/_old_lib.so in the build container:
#   git archive HEAD~1 tetris_gymnasium_b200/csrc include | tar -x -C /tmp/old && (cd /tmp/old && nvcc -gencode \
#     arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --use_fast_math \
#     -o $REPO/tools/_old_lib.so tetris_gymnasium_b200/csrc/tg_api.cu -lcudart)
#   gpurun -- 'bash tools/ab_lib.sh tools/_old_lib.so tetris_gymnasium_b200/libtetris_b200.so 3'
set -e
A=$1; B=$2; RUNS=${3:-3}; shift 3 || true
L=tetris_gymnasium_b200/libtetris_b200.so
cp "$A" /tmp/_ab_a.so; cp "$B" /tmp/_ab_b.so
run() { cp "$1" $L; touch $L
  python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-extra "${@:3}" 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('$2', round(d['value'] / 1e9, 4), 'G env-steps/s', round(d['roofline']['frac'], 4))"; }
for i in $(seq $RUNS); do run /tmp/_ab_a.so A "$@"; run /tmp/_ab_b.so B "$@"; done
cp /tmp/_ab_b.so $L; touch $L
