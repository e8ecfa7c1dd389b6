#!/bin/bash
mkdir -p gpurun_out
for two in 0 1; do echo "== VZ_GEMM_2CTA=$two"; VZ_GEMM_2CTA=$two timeout 300 python tools/gemm_bench.py; done > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log
