cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_fn.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench_suite.py --only c6 2>&1 | tail -3
