set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python scripts/tma_rows_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
tail -40 gpurun_out/probe.log
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "tma_rows" 2>&1 | tail -25
B200_OPTIONS=tma_rows=0 timeout 200 python scripts/microbench.py --only linear --json gpurun_out/mb_linear_off.json > gpurun_out/mb_linear_off.txt 2>&1
B200_OPTIONS=tma_rows=7 timeout 200 python scripts/microbench.py --only linear --json gpurun_out/mb_linear_on.json > gpurun_out/mb_linear_on.txt 2>&1
paste gpurun_out/mb_linear_off.txt gpurun_out/mb_linear_on.txt | awk '{print $1, $2, $10}' 
B200_OPTIONS=tma_rows=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_off.json 2> gpurun_out/bench_off.err; tail -c 600 gpurun_out/bench_off.json
B200_OPTIONS=tma_rows=7 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_on.json 2> gpurun_out/bench_on.err; tail -c 600 gpurun_out/bench_on.json
