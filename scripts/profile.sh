#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU): launch list of the timed command + full capture of the frame kernels.
# usage: scripts/profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
K='^k_(gbuffer|shadow|denoise|resolve|chain)'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 30 -c 5 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/prof_${TAG}.log 2>&1
[ "${RT:-0}" = 1 ] && ncu --set full --clock-control none -k regex:'^k_raytrace' -c 1 -o gpurun_out/prof_${TAG}_rt -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/prof_${TAG}_rt.log 2>&1
ls -la gpurun_out | tail -8
