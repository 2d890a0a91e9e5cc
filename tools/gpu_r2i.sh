#!/bin/bash
# round-2 session I: deferred norm2 after the bit-invariance fix, 32/112 register split, 8 vs 16 emit warps
set -u
T=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/gemm_epi_bench.py > $OUT/${T}_gemm_epi.log 2>&1; cat $OUT/${T}_gemm_epi.log
D3D_GEMM_EW_EMIT=16 timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
for v in "defer0:D3D_DEFER_LN2=0" "defer1:D3D_DEFER_LN2=1" "defer1_ew16:D3D_GEMM_EW_EMIT=16" "defer0_b:D3D_DEFER_LN2=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json
done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
