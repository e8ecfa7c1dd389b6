#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 600 $PT tests/test_gpu_attention.py -k "vit_attention_kernels" > gpurun_out/test_attn_unit.log 2>&1; echo "attn unit exit $?"
grep -E "vit attention impl|passed|failed|timeout|Error" gpurun_out/test_attn_unit.log | head
timeout 900 $PT tests/test_gpu_attention.py tests/test_gpu_e2e.py > gpurun_out/test_gpu_model.log 2>&1; echo "model exit $?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-400
VZ_VIT_ATTN_LEGACY=1 timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_legacy_attn.log 2>&1; echo "bench legacy exit $?"; tail -1 gpurun_out/bench_legacy_attn.log | cut -c1-200
