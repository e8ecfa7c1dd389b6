#!/bin/bash
# preprocess iteration: pixel tests, kernel timings per rows-per-CTA setting, one ncu --set full capture of the two passes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pixels.py tests/test_visual_prompts.py -m gpu -q -s > gpurun_out/r2_pre_tests.log 2>&1
echo "pixel tests rc=$?" > gpurun_out/r2_pre_summary.txt
for rb in 8 16 32; do
  VZ_PRE_RB=$rb timeout 300 python tools/hbm_kernels_bench.py > gpurun_out/r2_pre_bench_rb$rb.log 2>&1
  echo "rb=$rb: $(grep 'preprocess config3' gpurun_out/r2_pre_bench_rb$rb.log)" >> gpurun_out/r2_pre_summary.txt
done
grep -h "config2\|splice" gpurun_out/r2_pre_bench_rb16.log >> gpurun_out/r2_pre_summary.txt
VZ_PRE_FORM=two timeout 300 python tools/hbm_kernels_bench.py > gpurun_out/r2_pre_bench_two.log 2>&1
echo "old two-kernel form: $(grep 'preprocess config3' gpurun_out/r2_pre_bench_two.log)" >> gpurun_out/r2_pre_summary.txt
cat gpurun_out/r2_pre_summary.txt
