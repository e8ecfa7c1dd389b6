#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -m gpu ) > gpurun_out/test_gpu_attention.log 2>&1
echo "test_gpu_attention exit $?"; tail -3 gpurun_out/test_gpu_attention.log
for poly in 0 8 4; do
  ( VZ_ATTN_POLY=$poly timeout 300 python tools/attn_bench.py ) > gpurun_out/attn_bench_$poly.log 2>&1
  echo "VZ_ATTN_POLY=$poly: $(grep 'impl=1' gpurun_out/attn_bench_$poly.log)"
done
