set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python scripts/tma_rows_probe.py 2>&1 | grep -E "^tn|^   gw" | tail -30
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "tma_rows or linear" 2>&1 | tail -4
for v in 3 7; do
B200_OPTIONS=tma_rows=$v timeout 200 python scripts/profile_step.py > gpurun_out/cupti_tn_$v.txt 2>&1; grep -E "steps, device|void $|linear_bwd_weight_kernel|tc_tn_kernel|tc_reduce|tn_skinny|tc_skinny" gpurun_out/cupti_tn_$v.txt
done
timeout 300 python bench.py --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'])"
