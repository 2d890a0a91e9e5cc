#!/bin/bash
set -u
T=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench.json; tail -2 $OUT/${T}_bench.err
D3D_LN_ROWS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_rows1.json 2> $OUT/${T}_bench_rows1.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_rows1.json
