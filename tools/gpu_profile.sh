#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) full capture
# of the dominant kernel.  Each ncu run only after the same command exited 0 without ncu.
mkdir -p gpurun_out
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_kernel|splice_|text_|merge_rows)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 6 \
    -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30
