#!/bin/bash
# GPU check of the accumulation filter applied by k_shadow's threads (VKP_FUSE_ACCUMULATE, default on): the whole -m gpu suite, then the A/B
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r4g_tests.log 2>&1; tail -3 gpurun_out/r4g_tests.log
scripts/ab.sh r4g on "VKP_FUSE_ACCUMULATE=1" off "VKP_FUSE_ACCUMULATE=0"
