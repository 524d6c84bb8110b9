set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "bn_" > gpurun_out/pytest_bn.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_bn.log
for v in 0 1; do
B200_OPTIONS=bn_backward_fused=$v timeout 200 python scripts/profile_step.py --graphed > gpurun_out/cupti_bn_$v.txt 2>&1; grep -E "steps, device|affine_act|bn_finalize" gpurun_out/cupti_bn_$v.txt
B200_OPTIONS=bn_backward_fused=$v timeout 300 python bench.py --no-cpu-baseline --profile-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'])"
done
