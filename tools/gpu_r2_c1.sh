#!/bin/bash
# single-image latency: plain run, then the ncu launch list of ONE call (our kernels only, warm-up calls skipped)
mkdir -p gpurun_out
timeout 300 python tools/latency_c1.py > gpurun_out/r2_c1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_bf16|vit_attn|qattn|layernorm|fuse|softmax|row_stats|cls_rows|gather_rows|splice|text_|pre_|preprocess|patchify' \
  --launch-skip 1200 -c 480 --csv --log-file gpurun_out/r2_c1_launches.csv python tools/latency_c1.py > gpurun_out/r2_c1_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r2_c1.log | tail -2
