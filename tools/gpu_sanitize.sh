#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python tools/sanitize_smoke.py > gpurun_out/sanitize_run.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/sanitize_plain.log; tail -5 gpurun_out/memcheck.log
