#!/bin/bash
mkdir -p gpurun_out
( echo "== BN=256 2CTA_MIN=1"; VZ_GEMM_BN=256 VZ_GEMM_2CTA_MIN=1 timeout 300 python tools/gemm_bench.py; echo "== BN=256 2CTA off"; VZ_GEMM_BN=256 VZ_GEMM_2CTA=0 timeout 300 python tools/gemm_bench.py ) > gpurun_out/gemm_bench3.log 2>&1
grep -E "==|sa_|ffn|ca_q" gpurun_out/gemm_bench3.log
