set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 7 15 7 15; do
B200_OPTIONS=tma_rows=$v timeout 200 python scripts/profile_step.py > gpurun_out/cupti_tma_$v.txt 2>&1; grep -E "steps, device|void $" gpurun_out/cupti_tma_$v.txt
done
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tma_rows" 2>&1 | tail -2
B200_OPTIONS=tma_rows=15 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "linear" 2>&1 | tail -2
B200_OPTIONS=tma_rows=15 timeout 300 python bench.py --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ3', d['ms_per_step'])"
B200_OPTIONS=tma_rows=7 timeout 300 python bench.py --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ2', d['ms_per_step'])"
