#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 600 $PT tests/test_gpu_attention.py -k "vit_attention_kernels" > gpurun_out/test_attn_unit.log 2>&1; echo "attn unit exit $?"
grep -E "vit attention impl|passed|failed|timeout|Error" gpurun_out/test_attn_unit.log | head
timeout 1500 $PT tests > gpurun_out/test_all.log 2>&1; echo "all gpu tests exit $?"; tail -3 gpurun_out/test_all.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-330
