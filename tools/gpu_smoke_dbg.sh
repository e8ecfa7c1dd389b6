#!/bin/bash
mkdir -p gpurun_out
for cfg in "VZ_GEMM_SK=1 VZ_VIT_ATTN_LEGACY=0" "VZ_GEMM_SK=0 VZ_VIT_ATTN_LEGACY=0" "VZ_GEMM_SK=1 VZ_VIT_ATTN_LEGACY=1"; do
  echo "== $cfg"
  ( env $cfg timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -E "smoke|Error|vz:" | head -5 )
done
