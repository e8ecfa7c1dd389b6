#!/bin/bash
# probes: packed-exponential attention variant, ncu capture of a 32-row weight-streaming GEMM, launch list of the step
mkdir -p gpurun_out
for poly in 0 2; do
  VZ_ATTN_POLY=$poly timeout 300 python tools/attn_bench.py > gpurun_out/r2_attn_poly$poly.log 2>&1
  echo "VZ_ATTN_POLY=$poly: $(grep 'impl=1' gpurun_out/r2_attn_poly$poly.log | tr '\n' ' ')"
done
VZ_ATTN_POLY=2 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_e2e.py -m gpu -q -s > gpurun_out/r2_attn_poly2_tests.log 2>&1
echo "poly2 tests rc=$? $(tail -1 gpurun_out/r2_attn_poly2_tests.log)"; grep -h "min cos\|max_abs" gpurun_out/r2_attn_poly2_tests.log | head -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 --launch-skip 5 -c 1 -o gpurun_out/r2_gemm_m32 python tools/gemm_small.py "qf ffn2" > gpurun_out/r2_gemm_m32.log 2>&1
echo "ncu gemm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_bf16|vit_attn|qattn|layernorm|fuse|group_mean|softmax|row_stats|cls_rows|gather_rows|splice|text_|pre_|preprocess|patchify' \
  --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_launches_bench.log 2>&1
echo "ncu list rc=$?"
