cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_rollout.py -q -m gpu 2>&1 | tail -15
