#!/bin/bash
# One GPU visit: every GPU test file in its own process (a trapped kernel poisons only its own
# context), logs under gpurun_out/.  Usage: tools/gpu_round.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
for f in test_gpu_kernels test_gpu_pixels test_gpu_splice; do
  timeout 900 $PT tests/$f.py > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
done
for f in test_gpu_attention test_gpu_e2e; do
  timeout 1200 $PT tests/$f.py > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  VZ_FORCE_SIMPLE_GEMM=1 timeout 1200 $PT tests/$f.py > gpurun_out/${f}_simple.log 2>&1; echo "$f simple-gemm exit $?" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/bench.log
