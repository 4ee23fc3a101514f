#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02e_gputests.log 2>&1; tail -3 gpurun_out/r02e_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --impl reference ) > gpurun_out/r02e_bench_ref.log 2> gpurun_out/r02e_bench_ref.err; tail -c 300 gpurun_out/r02e_bench_ref.log
( time python bench.py ) > gpurun_out/r02e_bench.log 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.log; tail -3 gpurun_out/r02e_bench.err
python bench_suite.py --out gpurun_out/r02e_suite > gpurun_out/r02e_suite.log 2>&1; tail -3 gpurun_out/r02e_suite.log | cut -c1-200
