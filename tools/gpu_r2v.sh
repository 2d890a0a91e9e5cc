#!/bin/bash
# round-2 session V: TMA stores of the attention kernel issued by an idle control warp per slot
set -u
T=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" > $OUT/${T}_pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -4 $OUT/${T}_pytest_attn.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -4 $OUT/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_sw1.json 2> $OUT/${T}_bench_sw1.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_sw1.json; tail -2 $OUT/${T}_bench_sw1.err
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_nostorewarp.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_sw0.json 2> $OUT/${T}_bench_sw0.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_sw0.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_sw1_b.json 2> $OUT/${T}_bench_sw1_b.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_sw1_b.json
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_cfg4.json 2> $OUT/${T}_bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg4.json
