#!/bin/bash
# round-2 session H: leaner F4C epilogues (1-MUFU GELU, hoisted store addressing, shared-space staging, smem bias),
# deferred norm2 with 8 / 16 emit warps: parity, microbench, bench A/B
set -u
T=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "deferred or f4c or gelu or linear" > $OUT/${T}_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -5 $OUT/${T}_pytest_ops.log
timeout 300 python tools/gemm_epi_bench.py > $OUT/${T}_gemm_epi.log 2>&1; cat $OUT/${T}_gemm_epi.log
D3D_GEMM_EW_EMIT=16 timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_gelu_as.so timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
for v in "defer1:D3D_DEFER_LN2=1" "defer0:D3D_DEFER_LN2=0" "defer1_ew16:D3D_GEMM_EW_EMIT=16"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json
done
D3D_DEFER_LN2=0 D3D_LIB=$PWD/diff3dhpe_b200/libd3d_gelu_as.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_defer0_gelu_as.json 2> $OUT/${T}_bench_defer0_gelu_as.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_defer0_gelu_as.json
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
