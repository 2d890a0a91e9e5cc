#!/bin/bash
set -u
T=${1:-r02y3}
OUT=gpurun_out
mkdir -p $OUT
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_pn64.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_pn64.json 2> $OUT/${T}_bench_pn64.err; echo "bench rc=$?"; cut -c1-160 $OUT/${T}_bench_pn64.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_base.json 2> $OUT/${T}_bench_base.err; echo "bench rc=$?"; cut -c1-160 $OUT/${T}_bench_base.json
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_pn64.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_pn64_b.json 2> $OUT/${T}_bench_pn64_b.err; echo "bench rc=$?"; cut -c1-160 $OUT/${T}_bench_pn64_b.json
