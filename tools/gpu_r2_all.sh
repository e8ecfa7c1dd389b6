#!/bin/bash
# round-2 GPU check: every -m gpu test file on its own (no -x across files), then the default bench, then
# an ncu launch list of the single-image call.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
for f in test_gpu_pixels test_visual_prompts test_gpu_kernels test_gpu_attention test_gpu_splice test_text_inputs test_gpu_e2e test_gpu_llm test_gpu_train test_gpu_multi; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/r2_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2_summary.txt
  tail -3 gpurun_out/r2_$f.log | head -2
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_bf16|vit_attn|qattn|layernorm|fuse|group_mean|softmax|row_stats|cls_rows|gather_rows|splice|text_|pre_|preprocess|patchify' \
  --launch-skip 1200 -c 480 --csv --log-file gpurun_out/r2_c1_launches.csv python tools/latency_c1.py > gpurun_out/r2_c1_ncu.log 2>&1
echo "ncu c1 rc=$?" >> gpurun_out/r2_summary.txt
python tools/c1_breakdown.py 1 > gpurun_out/r2_c1_breakdown.log 2>&1
VZ_GRAPHS=0 python tools/latency_c1.py > gpurun_out/r2_c1_nograph.log 2>&1; python tools/latency_c1.py > gpurun_out/r2_c1.log 2>&1
cat gpurun_out/r2_c1_nograph.log gpurun_out/r2_c1.log | grep "config 1"
cat gpurun_out/r2_summary.txt
