#!/bin/bash
set -u
T=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
for ni in 1 2 1 2; do
  echo "N_INNER=$ni" >> $OUT/${T}_gemm_stagger.log
  D3D_GEMM_N_INNER=$ni timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_stagger.log 2>&1
done
cat $OUT/${T}_gemm_stagger.log
D3D_GEMM_N_INNER=2 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear or reduction or deferred" > $OUT/${T}_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -3 $OUT/${T}_pytest_ops.log
