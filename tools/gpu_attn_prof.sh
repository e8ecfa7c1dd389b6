#!/bin/bash
# attention kernel: parity test, stand-alone timing, then one ncu --set full capture with source
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -m gpu ) > gpurun_out/test_gpu_attention.log 2>&1
echo "test_gpu_attention exit $?"
( timeout 300 python tools/attn_bench.py ) > gpurun_out/attn_bench.log 2>&1
echo "attn_bench exit $?"; grep impl gpurun_out/attn_bench.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 3 -c 2 \
    -f -o gpurun_out/prof_attn_v4 python tools/attn_bench.py ) > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
