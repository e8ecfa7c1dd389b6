#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 600 $PT tests/test_gpu_kernels.py -k "two_cta" > gpurun_out/test_2cta.log 2>&1; echo "2cta unit exit $?"; grep -E "2-CTA gemm|passed|failed|timeout|rror" gpurun_out/test_2cta.log | head -12
bash tools/gpu_gemm_bench.sh 2>&1 | grep -E "==|qkv|fc1|fc2|^o "
