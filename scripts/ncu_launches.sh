#!/bin/bash
# ncu launch list (device time of EVERY kernel, torch's included) of ONE timed bench step (graph replay).
# bench.py brackets the timed region with cudaProfilerStart/Stop when B200_NCU_RANGE=1, so warm-up, graph capture and the
# per-kernel event pass are not replayed under ncu (~0.4 s per profiled launch: ~500 launches -> ~4 min).
# (--eager: since round 2 the captured step has parallel branches and a 200 KB-shared-memory draw kernel that ncu's
# graph-node replay cannot launch; the eager step launches the same kernels one after the other.)
# Usage (on the GPU box, through gpurun): bash scripts/ncu_launches.sh <tag>
tag=${1:-r1}
B200_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --graph-profiling node --metrics gpu__time_duration.sum \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --eager --steps 1 --warmup 3 --profile-steps 0 --no-cpu-baseline > gpurun_out/launches_${tag}.bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv --all > gpurun_out/launches_${tag}.summary.txt
head -50 gpurun_out/launches_${tag}.summary.txt
