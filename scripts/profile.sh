#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU; gpurun brings back at most 64 MiB per call, so frame and build are two calls).
#   scripts/profile.sh <tag> frame   launch list of the timed command + full capture of one frame's kernels (chain schedule 7: 11
#                                    kernels matching the regex per frame)
#   scripts/profile.sh <tag> build   launch list + full capture of the acceleration-structure build kernels (build_probe.py)
# then here: python scripts/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <tag> --kernels-json   (and prof_build_<tag> likewise)
TAG=${1:-r2}; WHAT=${2:-frame}
mkdir -p gpurun_out
if [ "$WHAT" = frame ]; then
K='^k_(gbuffer|shadow|denoise|resolve|chain)'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 110 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 55 -c 11 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/prof_${TAG}.log 2>&1
[ "${RT:-0}" = 1 ] && ncu --set full --clock-control none -k regex:'^k_raytrace' -c 1 -o gpurun_out/prof_${TAG}_rt -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}_rt.log 2>&1
else
# the build of the default scene: 5 BLAS, the large one (328 k triangles: the statue) is the 2nd; first context of the process
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_(rs_|scan_|ploc|fit|emit|morton|centroid|triangle_boxes|mark_keep|hierarchy|instance_boxes|build_result|set_int)' -c 400 --csv \
    --log-file gpurun_out/launches_build_${TAG}.csv python scripts/build_probe.py 0 > gpurun_out/launches_build_${TAG}.log 2>&1
# one launch each of the kernels of the large BLAS: -s skips the first (small) BLAS
ncu --set full --clock-control none -k regex:'^k_(rs_scatter|rs_hist|ploc_nn_dev|ploc_merge_dev|ploc_finish|fit|emit_nodes|emit_triangles|morton)' -s 40 -c 28 \
    -o gpurun_out/prof_build_${TAG} -f python scripts/build_probe.py 0 > gpurun_out/prof_build_${TAG}.log 2>&1
fi
ls -la gpurun_out | tail -8
