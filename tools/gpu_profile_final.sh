#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_kernel|splice_|text_|merge_rows|transpose_kernel|softmax_rows|row_stats|collate)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 8 \
    -o gpurun_out/prof_gemm_v8 $CMD > gpurun_out/ncu_full.log 2>&1
echo "gemm capture exit $?"
