#!/bin/bash
# round-1 final evidence: full GPU suite, smoke, bench, ncu launch list + full captures of the main kernels
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/test_all.log 2>&1
echo "pytest -m gpu exit $?"; tail -3 gpurun_out/test_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
( timeout 900 python bench.py ) > gpurun_out/bench_full.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_full.log
( timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.log 2>&1
echo "reference arm exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
( timeout 300 python tools/hbm_kernels_bench.py ) > gpurun_out/hbm_kernels.log 2>&1
echo "hbm bench exit $?"; cat gpurun_out/hbm_kernels.log | tail -4
( timeout 300 python tools/attn_bench.py ) > gpurun_out/attn_bench.log 2>&1; grep impl gpurun_out/attn_bench.log
( timeout 300 python tools/gemm_bench.py ) > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_|splice_|text_|merge_rows|softmax_rows|row_stats|collate)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 8 \
    -f -o gpurun_out/prof_gemm_v9 $CMD > gpurun_out/ncu_full.log 2>&1
echo "gemm capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:vit_attn_tc -s 5 -c 1 \
    -f -o gpurun_out/prof_attn_v9 $CMD > gpurun_out/ncu_full_attn.log 2>&1
echo "attn capture exit $?"
ncu --set full --clock-control none -k "regex:^(preprocess_|fuse_kernel|splice_scatter)" -s 4 -c 4 \
    -f -o gpurun_out/prof_hbm_v9 $CMD > gpurun_out/ncu_full_hbm.log 2>&1
echo "hbm capture exit $?"
