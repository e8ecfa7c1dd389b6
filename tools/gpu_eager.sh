#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/eager_baseline.py > gpurun_out/eager.log 2>&1; echo "eager exit $?"; tail -3 gpurun_out/eager.log
