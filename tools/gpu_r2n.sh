#!/bin/bash
# round-2 session N: row-wise kernel diet (rsqrt + Newton, no clip division at eval), shared-space accesses in the attention
# epilogue: full GPU suite, bench, ncu of the lift / head kernels
set -u
T=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_b.json 2> $OUT/${T}_bench_b.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_b.json
for k in head_ddim lift_ln; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/${T}_full_$k -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_full_$k.log 2>&1; echo "full $k rc=$?"
done
