#!/bin/bash
# whole GPU suite + smoke + bench (what the driver runs at round end)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/test_all.log 2>&1
echo "pytest -m gpu exit $?"; tail -4 gpurun_out/test_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
( timeout 900 python bench.py ) > gpurun_out/bench_full.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_full.log
