#!/bin/bash
mkdir -p gpurun_out
for mn in 148 1; do echo "== VZ_GEMM_2CTA_MIN=$mn"; VZ_GEMM_2CTA_MIN=$mn timeout 300 python tools/gemm_bench.py; done > gpurun_out/gemm_bench2.log 2>&1
grep -E "==|sa_|ffn|ca_q" gpurun_out/gemm_bench2.log
