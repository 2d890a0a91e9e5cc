#!/bin/bash
# round-2 session C: F4C with the tcgen05 attention epilogue + faster row-wise stores
set -u
T=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "f4c or F4C" > $OUT/${T}_pytest_f4c.log 2>&1; echo "pytest f4c rc=$?" >> $OUT/${T}_pytest_f4c.log; tail -25 $OUT/${T}_pytest_f4c.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -5 $OUT/${T}_pytest.log
timeout 600 python bench.py --gemm f4c --steps 3 --warmup 3 > $OUT/${T}_bench_f4c.json 2> $OUT/${T}_bench_f4c.err; echo "bench rc=$?"; cut -c1-300 $OUT/${T}_bench_f4c.json; tail -3 $OUT/${T}_bench_f4c.err
timeout 600 python bench.py --gemm f8c --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_f8c.json 2> $OUT/${T}_bench_f8c.err; echo "bench rc=$?"; cut -c1-300 $OUT/${T}_bench_f8c.json; tail -3 $OUT/${T}_bench_f8c.err
