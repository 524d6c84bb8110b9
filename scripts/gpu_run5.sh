set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python scripts/microbench.py --json gpurun_out/mb_all_ffma2.json > gpurun_out/mb_all_ffma2.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --kernel-report gpurun_out/bench_kernels_ffma2.json > gpurun_out/bench_ffma2.json 2> gpurun_out/bench_ffma2.err; tail -c 300 gpurun_out/bench_ffma2.json
