set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_step.py > gpurun_out/step_kernels_cupti.txt 2>&1; head -40 gpurun_out/step_kernels_cupti.txt
