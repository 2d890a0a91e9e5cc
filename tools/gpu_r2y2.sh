#!/bin/bash
set -u
T=${1:-r02y2}
OUT=gpurun_out
mkdir -p $OUT
for pr in 3 2 1 0 3; do
  echo "L2_PROMO=$pr" >> $OUT/${T}_gemm_promo.log
  D3D_TMA_L2_PROMO=$pr timeout 300 python tools/gemm_epi_bench.py 2>&1 | grep -E "qkv|proj \+ residual  |fc1 gelu  |fc2 \+" >> $OUT/${T}_gemm_promo.log
done
cat $OUT/${T}_gemm_promo.log
