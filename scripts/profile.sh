#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU): launch list + full capture of the frame kernels.
# usage: scripts/profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_(gbuffer|shadow|denoise|resolve)' -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_(gbuffer|shadow|denoise|resolve)' -s 24 -c 4 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
