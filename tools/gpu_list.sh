#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_kernel|splice_|text_|merge_rows|transpose_kernel|softmax_rows|row_stats)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
