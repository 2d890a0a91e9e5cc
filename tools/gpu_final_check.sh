#!/bin/bash
# last call of the round: the committed head on a fresh box -- CPU-free checks the driver will repeat
set -u
T=${1:-r02zz}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -4 $OUT/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_smoke.log 2>&1; tail -2 $OUT/${T}_smoke.log
timeout 900 python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench.json; tail -2 $OUT/${T}_bench.err
