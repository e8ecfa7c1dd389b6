#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 10 -c 1 \
    -o gpurun_out/prof_attn2 $CMD > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
