#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu ) > gpurun_out/test_gpu_kernels.log 2>&1
echo "kernels exit $?"; tail -2 gpurun_out/test_gpu_kernels.log
( timeout 300 python tools/gemm_bench.py ) > gpurun_out/gemm_bench.log 2>&1
echo "gemm_bench exit $?"; grep -v "+sk" gpurun_out/gemm_bench.log | tail -9
