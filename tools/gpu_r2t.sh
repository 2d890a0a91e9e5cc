#!/bin/bash
# round-2 session T: L2 prefetch of the next m-tile's A rows in the F4C GEMM
set -u
T=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
for pf in 1 0; do
  echo "PREFETCH_A=$pf" >> $OUT/${T}_gemm_pf.log
  D3D_GEMM_PREFETCH_A=$pf timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_pf.log 2>&1
done
cat $OUT/${T}_gemm_pf.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear or reduction" > $OUT/${T}_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -3 $OUT/${T}_pytest_ops.log
for v in "pf1:D3D_GEMM_PREFETCH_A=1" "pf0:D3D_GEMM_PREFETCH_A=0" "pf1_b:D3D_GEMM_PREFETCH_A=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json
done
