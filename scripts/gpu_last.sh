cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_kernels.py -q -x -k "tma_rows" 2>&1 | tail -2
timeout 45 python bench.py --no-cpu-baseline --profile-steps 0 --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'])"
