# Final evidence of a round: bench lines (configs B, D, E + the CPU reference arm), CUPTI and ncu launch lists, microbench.
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
mkdir -p gpurun_out
timeout 400 python bench.py --kernel-report gpurun_out/bench_kernel_events_${tag}.json > gpurun_out/bench_${tag}_configB_n1.json 2> gpurun_out/bench_B.err; tail -c 400 gpurun_out/bench_${tag}_configB_n1.json
timeout 300 python bench.py --config D > gpurun_out/bench_${tag}_configD_n1.json 2> gpurun_out/bench_D.err; tail -c 200 gpurun_out/bench_${tag}_configD_n1.json
timeout 300 python bench.py --config E > gpurun_out/bench_${tag}_configE_n1.json 2> gpurun_out/bench_E.err; tail -c 200 gpurun_out/bench_${tag}_configE_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_${tag}_reference_arm.json
timeout 200 python scripts/profile_step.py --graphed > gpurun_out/step_kernels_cupti_${tag}.txt 2>&1; head -12 gpurun_out/step_kernels_cupti_${tag}.txt
timeout 300 python scripts/microbench.py > gpurun_out/microbench_${tag}.txt 2>&1
bash scripts/ncu_launches.sh ${tag} | head -30
