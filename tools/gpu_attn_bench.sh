#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/attn_bench.py ) > gpurun_out/attn_bench.log 2>&1
cat gpurun_out/attn_bench.log | grep impl
