set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_graphed.py -q -k "tma_rows or bn_act or linear or scratch" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sub.log
timeout 200 python scripts/microbench.py --only linear --json gpurun_out/mb_linear_v2.json > gpurun_out/mb_linear_v2.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --kernel-report gpurun_out/bench_kernels_v2.json > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 300 gpurun_out/bench_v2.json
