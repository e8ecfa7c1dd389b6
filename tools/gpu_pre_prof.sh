#!/bin/bash
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 30 -c 1 \
    -f -o gpurun_out/prof_pre python tools/hbm_kernels_bench.py ) > gpurun_out/ncu_pre.log 2>&1
echo "capture exit $?"
