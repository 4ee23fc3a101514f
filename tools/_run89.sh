cd $GRAFT_REPO_ROOT
L=tetris_gymnasium_b200/libtetris_b200.so
b() { timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value']/1e9, d['roofline']['frac'])"; }
cp $L /tmp/cold_lib.so
TG_NVCC_FLAGS="-DTG_RESET_INLINE" python -m tetris_gymnasium_b200._build > /dev/null 2>&1; cp $L /tmp/inl_lib.so
for i in 1 2 3; do
cp /tmp/cold_lib.so $L; touch $L; b cold
cp /tmp/inl_lib.so $L; touch $L; b inline
done
cp /tmp/cold_lib.so $L; touch $L
timeout 900 python -m pytest tests/test_gpu_base.py tests/test_gpu_10k_episodes.py -x -q -m gpu 2>&1 | tail -3
