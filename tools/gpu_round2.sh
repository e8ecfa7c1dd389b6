#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --tb=short -rA -p no:cacheprovider -m gpu"
timeout 900 $PT tests/test_gpu_kernels.py > gpurun_out/test_gpu_kernels.log 2>&1; echo "kernels exit $?"
timeout 900 $PT tests/test_gpu_attention.py tests/test_gpu_e2e.py > gpurun_out/test_gpu_model.log 2>&1; echo "model exit $?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log
