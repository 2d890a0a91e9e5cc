#!/bin/bash
# round-2 session J: residual update as a TMA reduction (EPI_F32_RED) vs the load-add-store epilogue
set -u
T=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear or reduction or deferred" > $OUT/${T}_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -5 $OUT/${T}_pytest_ops.log
timeout 300 python tools/gemm_epi_bench.py > $OUT/${T}_gemm_epi.log 2>&1; cat $OUT/${T}_gemm_epi.log
D3D_GEMM_RED=0 timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
for v in "defer1_red1:D3D_GEMM_RED=1" "defer0_red1:D3D_DEFER_LN2=0" "defer0_red0:D3D_DEFER_LN2=0 D3D_GEMM_RED=0" "defer1_red0:D3D_GEMM_RED=0" "defer0_red1_b:D3D_DEFER_LN2=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json
done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
