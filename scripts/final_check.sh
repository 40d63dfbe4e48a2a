#!/bin/bash
# Round-end check on one B200 (run under gpurun): the -m gpu suite, smoke(), the default bench line, the launch list of the timed command
# and a full ncu capture of one k_shadow launch (an accumulation frame: the kernel that also applies the filter).  TAG = file prefix.
TAG=${1:-r4i}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 240 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_cfg2.json
K='^k_(gbuffer|shadow|denoise|resolve|chain)'
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 110 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/launches_${TAG}.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'^k_shadow' -s 5 -c 1 \
    -o gpurun_out/prof_${TAG}_shadow -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/prof_${TAG}_shadow.log 2>&1
ls -la gpurun_out | grep ${TAG}
