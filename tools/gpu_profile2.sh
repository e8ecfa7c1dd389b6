#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_kernel|splice_|text_|merge_rows|transpose_kernel|softmax_rows)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 10 -c 2 \
    -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
