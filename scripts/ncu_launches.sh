#!/bin/bash
# ncu launch list (device time of EVERY kernel, torch's included) of one eager bench run.
# Usage (on the GPU box, through gpurun): bash scripts/ncu_launches.sh <tag>
tag=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
    --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --eager --steps 2 --warmup 3 --profile-steps 0 --no-cpu-baseline > gpurun_out/launches_${tag}.bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv > gpurun_out/launches_${tag}.summary.txt
tail -60 gpurun_out/launches_${tag}.summary.txt
