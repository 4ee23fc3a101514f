#!/bin/bash
# parity of the grouped kernels + A/B of two libs on the grouped path (steady state, tools/time_grouped.py) and bench_suite C3
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grouped.py tests/test_gpu_cross_kernel.py tests/test_gpu_streams.py tests/test_gpu_custom_set.py tests/test_gpu_holder.py -x -q 2>&1 | tail -3
L=tetris_gymnasium_b200/libtetris_b200.so
cp $L /tmp/_keep.so
for i in 1 2; do for lib in "$@"; do
  cp $lib $L; touch $L
  echo "== $lib"; TG_GROUPED_SPLIT=1 python tools/time_grouped.py 2>&1 | tail -1
done; done
bash tools/ab_c3.sh 1 "$@" 2>&1 | tail -8
cp /tmp/_keep.so $L
