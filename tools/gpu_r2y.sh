#!/bin/bash
# L2 eviction-hint sweep for the F4C GEMM operands (0 normal, 1 evict_last, 2 evict_first)
set -u
T=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
for ha in 1 0 2; do for hb in 1 0 2; do
  echo "HINT_A=$ha HINT_B=$hb" >> $OUT/${T}_gemm_hints.log
  D3D_GEMM_HINT_A=$ha D3D_GEMM_HINT_B=$hb timeout 300 python tools/gemm_epi_bench.py 2>&1 | grep -E "qkv|proj \+ residual  |fc1 gelu  |fc2 \+" >> $OUT/${T}_gemm_hints.log
done; done
cat $OUT/${T}_gemm_hints.log
