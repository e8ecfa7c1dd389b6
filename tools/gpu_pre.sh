#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_pixels.py -x -q -m gpu ) > gpurun_out/test_gpu_pixels.log 2>&1
echo "pixels exit $?"; tail -2 gpurun_out/test_gpu_pixels.log
( timeout 300 python tools/hbm_kernels_bench.py ) > gpurun_out/hbm_kernels.log 2>&1
echo "hbm bench exit $?"; cat gpurun_out/hbm_kernels.log | tail -4
( timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:preprocess_ -c 120 --csv --log-file gpurun_out/pre_launches.csv python tools/hbm_kernels_bench.py ) > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/pre_launches.csv")) if len(r)>5]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value"); i_g=h.index("Grid Size"); i_k=h.index("Kernel Name")
acc={}
for r in rows[1:]:
    acc.setdefault((r[i_k].split("(")[0][-22:], r[i_g], r[i_n]), []).append(float(r[i_v].replace(",","")))
for k,v in acc.items(): print(k, "median", sorted(v)[len(v)//2], "n", len(v))
PY
