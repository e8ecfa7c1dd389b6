#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 900 $PT tests/test_gpu_kernels.py tests/test_gpu_attention.py tests/test_gpu_e2e.py > gpurun_out/test_k_m.log 2>&1; echo "kernels+model exit $?"; grep -E "passed|failed" gpurun_out/test_k_m.log | tail -2
VZ_BENCH_LN=1 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_ln.log 2>&1; grep "M=" gpurun_out/gemm_ln.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-230; tail -1 gpurun_out/bench.log | grep -o '"roofline.*'
