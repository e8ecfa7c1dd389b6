#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 900 $PT tests/test_gpu_pixels.py tests/test_gpu_attention.py tests/test_gpu_e2e.py > gpurun_out/test_px_attn.log 2>&1; echo "pixels+attention+e2e exit $?"; grep -E "vit attention impl|passed|failed" gpurun_out/test_px_attn.log | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-330
timeout 300 python tools/hbm_kernels_bench.py > gpurun_out/hbm_kernels.log 2>&1; tail -4 gpurun_out/hbm_kernels.log
