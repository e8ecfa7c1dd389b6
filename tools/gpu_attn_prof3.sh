#!/bin/bash
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 3 -c 1 \
    -f -o gpurun_out/prof_attn_v6 python tools/attn_bench.py ) > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
