set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B200_CLEAN_BUILD_TEST=1 timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_final.log
grep -E "clean build|built|smoke:" gpurun_out/pytest_gpu_final.log | head -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|Error|error" | head
timeout 400 python bench.py --kernel-report gpurun_out/bench_kernel_events_r02.json > gpurun_out/bench_r02_configB_n1.json 2> gpurun_out/bench_B.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_configB_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
