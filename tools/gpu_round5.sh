#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 900 $PT tests/test_gpu_pixels.py tests/test_gpu_attention.py > gpurun_out/test_px_attn.log 2>&1; echo "pixels+attention exit $?"; grep -E "vit attention impl|passed|failed" gpurun_out/test_px_attn.log | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-330
KREGEX='regex:^(gemm_bf16|layernorm_kernel|fuse_kernel|cls_rows|gather_rows|patchify|vit_attn|qattn32|preprocess_kernel|splice_|text_|merge_rows|transpose_kernel|softmax_rows)'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
