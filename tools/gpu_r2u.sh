#!/bin/bash
set -u
T=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
for pf in 2 0; do
  echo "PREFETCH_A=$pf" >> $OUT/${T}_gemm_pf.log
  D3D_GEMM_PREFETCH_A=$pf timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_pf.log 2>&1
done
cat $OUT/${T}_gemm_pf.log
