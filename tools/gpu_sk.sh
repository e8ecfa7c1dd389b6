#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stream_k or two_cta or plain or batched" ) > gpurun_out/test_gpu_kernels.log 2>&1
echo "test_gpu_kernels exit $?"; tail -5 gpurun_out/test_gpu_kernels.log
( timeout 300 python tools/gemm_bench.py ) > gpurun_out/gemm_bench.log 2>&1
echo "gemm_bench exit $?"; cat gpurun_out/gemm_bench.log | tail -20
