#!/bin/bash
# A/B on the SAME box: bench with the 2-CTA GEMM form off and on, twice each (interleaved)
mkdir -p gpurun_out
for i in 1 2; do
  for two in 0 1; do
    VZ_GEMM_2CTA=$two timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ab_${two}_$i.log 2>&1
    echo "2CTA=$two run $i: $(tail -1 gpurun_out/bench_ab_${two}_$i.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['gemm_ms_per_step'],2), 'ms', round(d['roofline']['achieved']), 'TF/s', d['clocks']['sm_mhz'], 'MHz', d['clocks']['power_w_max'], 'W')")"
  done
done
