#!/bin/bash
# same-box A/B of the stream-K GEMM tail (VZ_GEMM_SK=0|1), alternating
mkdir -p gpurun_out
for rep in 1 2; do
  for v in 0 1; do
    VZ_GEMM_SK=$v timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_${v}_${rep}.log 2>&1
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_ab_${v}_${rep}.log").read().strip().splitlines()[-1])
print("SK=${v} rep=${rep}: %.1f images/s  %.3f ms/step  gemm %.3f ms/step  frac %.3f  sm_mhz %s" % (d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
  done
done
