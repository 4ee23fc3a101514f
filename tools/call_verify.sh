#!/bin/bash
# full GPU suite, smoke, both bench arms, bench_suite and ncu captures of the functional / CNN kernels: bash tools/call_verify.sh (under gpurun)
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/verify_gputests.log 2>&1; tail -3 gpurun_out/verify_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/verify_bench_ref.log 2> gpurun_out/verify_bench_ref.err; tail -c 200 gpurun_out/verify_bench_ref.log
python bench.py > gpurun_out/verify_bench.log 2> gpurun_out/verify_bench.err; tail -c 200 gpurun_out/verify_bench.log; tail -3 gpurun_out/verify_bench.err
python bench_suite.py --out gpurun_out/verify_suite > gpurun_out/verify_suite.log 2>&1; tail -2 gpurun_out/verify_suite.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fn_step_tile -s 6 -c 1 -f -o gpurun_out/verify_fn python tools/prof_paths.py fn --envs 1048576 > gpurun_out/verify_fn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cnn_obs3 -s 6 -c 1 -f -o gpurun_out/verify_cnn python tools/prof_paths.py cnn --envs 65536 > gpurun_out/verify_cnn.log 2>&1
ls gpurun_out | grep verify_ || true
