#!/bin/bash
# round-2 session B: bring-up of the block-scaled e2m1 (F4C) GEMM
set -u
T=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "f4c or F4C" > $OUT/${T}_pytest_f4c.log 2>&1; echo "pytest f4c rc=$?" >> $OUT/${T}_pytest_f4c.log; tail -25 $OUT/${T}_pytest_f4c.log
timeout 300 python tools/gemm_mode_bench.py 2115072 f8c f4c > $OUT/${T}_gemm_modes.log 2>&1; cat $OUT/${T}_gemm_modes.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -5 $OUT/${T}_pytest.log
timeout 600 python bench.py --gemm f4c --steps 3 --warmup 3 > $OUT/${T}_bench_f4c.json 2> $OUT/${T}_bench_f4c.err; echo "bench rc=$?"; cut -c1-300 $OUT/${T}_bench_f4c.json; tail -3 $OUT/${T}_bench_f4c.err
