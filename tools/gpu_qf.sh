#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_attention.py -x -q -m gpu -rA ) > gpurun_out/test_gpu_e2e.log 2>&1
echo "e2e exit $?"; grep -E "cos|max_abs|passed|failed|Error" gpurun_out/test_gpu_e2e.log | tail -12
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | grep -E "smoke|Error" | tail -3
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-200; tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('launches', d['gpu_launches'], 'gemm ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])"
