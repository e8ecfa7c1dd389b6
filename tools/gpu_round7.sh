#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/hbm_kernels_bench.py > gpurun_out/hbm_kernels.log 2>&1; echo "hbm bench exit $?"; cat gpurun_out/hbm_kernels.log | tail -5
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 8 \
    -o gpurun_out/prof_gemm_v5 $CMD > gpurun_out/ncu_full.log 2>&1
echo "gemm capture exit $?"
