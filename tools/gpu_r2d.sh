#!/bin/bash
# round-2 session D: F4C as the default mode -- parity report, full tests, smoke, bench, ncu launch list + full GEMM captures
set -u
T=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/parity_report.py f8c f4c > $OUT/${T}_parity_report.log 2>&1; cat $OUT/${T}_parity_report.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_smoke.log 2>&1; tail -2 $OUT/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
# every kernel launch of one step of the bench command (cold-cache, serialised under ncu: compare SHARES)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1100 --csv --log-file $OUT/${T}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_launches.log 2>&1; echo "launches rc=$?"
# full captures of the four GEMMs of one block (qkv, proj, fc1, fc2) at the bench size
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o $OUT/${T}_full_gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_full_gemm.log 2>&1; echo "full gemm rc=$?"
ls -la $OUT | tail -12
