#!/bin/bash
# round-2 session A: new parity tests + bench with the timed-output check + ring-depth A/B of the F8C GEMM
set -u
T=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -5 $OUT/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_smoke.log 2>&1; tail -1 $OUT/${T}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${T}_bench.json; tail -3 $OUT/${T}_bench.err
for lib in libdiff3d_b200 libd3d_ring168 libd3d_ring136; do
  D3D_LIB=$PWD/diff3dhpe_b200/$lib.so timeout 300 python tools/gemm_mode_bench.py 2115072 f8c >> $OUT/${T}_ring_ab.log 2>&1
done
cat $OUT/${T}_ring_ab.log
timeout 400 python bench.py --config cfg2 --steps 3 --warmup 3 > $OUT/${T}_bench_cfg2.json 2> $OUT/${T}_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg2.json
timeout 400 python bench.py --config cfg4 --steps 3 --warmup 3 > $OUT/${T}_bench_cfg4.json 2> $OUT/${T}_bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg4.json
