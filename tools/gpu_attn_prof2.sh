#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-600
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 3 -c 1 \
    -f -o gpurun_out/prof_attn_v5 python tools/attn_bench.py ) > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
