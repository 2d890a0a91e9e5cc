#!/bin/bash
# round-2 session G: deferred norm2 (proj emits fc1's operand, fc1 applies the LayerNorm): parity, microbench, bench A/B
set -u
T=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "deferred or f4c" > $OUT/${T}_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -5 $OUT/${T}_pytest_ops.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
timeout 300 python tools/gemm_epi_bench.py > $OUT/${T}_gemm_epi.log 2>&1; cat $OUT/${T}_gemm_epi.log
D3D_GEMM_EW_EMIT=16 timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_epi.log 2>&1; tail -9 $OUT/${T}_gemm_epi.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_defer1.json 2> $OUT/${T}_bench_defer1.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_defer1.json
D3D_DEFER_LN2=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_defer0.json 2> $OUT/${T}_bench_defer0.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_defer0.json
timeout 300 python tools/parity_report.py > $OUT/${T}_parity_report.log 2>&1; tail -12 $OUT/${T}_parity_report.log
