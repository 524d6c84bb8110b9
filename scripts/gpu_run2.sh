set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tma_rows_nn -s 3 -c 1 -o gpurun_out/ncu_tma_nn -f python scripts/microbench.py --only "linear_fwd:32+0->32@L0" --iters 2 --warmup 3 > gpurun_out/ncu_nn.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tma_rows_tn -s 3 -c 1 -o gpurun_out/ncu_tma_tn -f python scripts/microbench.py --only "linear_bwd_weight:32+0->32@L0" --iters 2 --warmup 3 > gpurun_out/ncu_tn.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.json
