#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 900 $PT tests/test_gpu_kernels.py tests/test_gpu_attention.py tests/test_gpu_e2e.py > gpurun_out/test_k_m.log 2>&1; echo "kernels+model exit $?"; grep -E "passed|failed" gpurun_out/test_k_m.log | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-330; tail -1 gpurun_out/bench.log | grep -o '"roofline.*'
