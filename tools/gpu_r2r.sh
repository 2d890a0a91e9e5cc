#!/bin/bash
# round-2 session R: how much does the ring depth matter?  (170 KB budget: 4 stages with 8 epilogue warps, 3 with 16)
set -u
T=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/gemm_epi_bench.py > $OUT/${T}_gemm_epi.log 2>&1; cat $OUT/${T}_gemm_epi.log
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_smem170.so timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
D3D_GEMM_EW_GELU=8 timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
