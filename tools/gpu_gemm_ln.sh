#!/bin/bash
mkdir -p gpurun_out
VZ_BENCH_LN=1 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_ln.log 2>&1; cat gpurun_out/gemm_ln.log
