#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "k_by_n or stream_k" ) > gpurun_out/test_gpu_kernels.log 2>&1
echo "kernels exit $?"; tail -5 gpurun_out/test_gpu_kernels.log
( timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_attention.py -x -q -m gpu ) > gpurun_out/test_gpu_e2e.log 2>&1
echo "e2e exit $?"; tail -3 gpurun_out/test_gpu_e2e.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-330
