#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_|splice_|text_|merge_rows|softmax_rows|row_stats|collate)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none -k "regex:^(preprocess_|fuse_kernel|splice_scatter)" -s 4 -c 4 \
    -f -o gpurun_out/prof_hbm_v9 $CMD > gpurun_out/ncu_full_hbm.log 2>&1
echo "hbm capture exit $?"
( timeout 300 python tools/hbm_kernels_bench.py ) > gpurun_out/hbm_kernels.log 2>&1; tail -4 gpurun_out/hbm_kernels.log
