#!/bin/bash
# round-2 final evidence session: parity report, full GPU suite, smoke, bench (+ reference arm), cfg2 / cfg4, ncu launch list,
# ncu --set full of the four GEMMs of one block (-> profiles/gemm_traffic.json), memcheck of smoke()
set -u
T=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/parity_report.py f8c f4c > $OUT/${T}_parity_report.log 2>&1; cat $OUT/${T}_parity_report.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -6 $OUT/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_smoke.log 2>&1; tail -2 $OUT/${T}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/${T}_bench_reference.json 2> $OUT/${T}_bench_reference.err; echo "reference rc=$?"; cut -c1-200 $OUT/${T}_bench_reference.json
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_cfg2.json 2> $OUT/${T}_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg2.json
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_cfg4.json 2> $OUT/${T}_bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg4.json
# every kernel launch of one step of the bench command (cold-cache, serialised under ncu: compare SHARES)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1100 --csv --log-file $OUT/${T}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_launches.log 2>&1; echo "launches rc=$?"
# full captures of the four GEMMs of one block (qkv, proj, fc1, fc2) at the bench size
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o $OUT/${T}_full_gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_full_gemm.log 2>&1; echo "full gemm rc=$?"
for k in postnorm_add_ln2 ln_split2; do
  timeout 600 ncu --set full --clock-control none -k regex:$k -s 20 -c 1 -o $OUT/${T}_full_$k -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_full_$k.log 2>&1; echo "full $k rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_memcheck_smoke.log 2>&1; tail -4 $OUT/${T}_memcheck_smoke.log
ls -la $OUT | tail -5
