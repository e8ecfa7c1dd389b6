#!/bin/bash
mkdir -p gpurun_out
for bn in 0 256 192; do echo "== VZ_GEMM_BN=$bn"; VZ_GEMM_BN=$bn timeout 300 python tools/gemm_bench.py; done > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log
