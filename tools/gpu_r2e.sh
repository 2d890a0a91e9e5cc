#!/bin/bash
# round-2 session E: packed temporal attention (F <= 64), row-wise kernel diet, residual prefetch depth A/B
set -u
T=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_depth2.json 2> $OUT/${T}_bench_depth2.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_depth2.json
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_depth1.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_depth1.json 2> $OUT/${T}_bench_depth1.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_depth1.json
D3D_GEMM_EW_F32=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_ewf32_16.json 2> $OUT/${T}_bench_ewf32_16.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_ewf32_16.json
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 > $OUT/${T}_bench_cfg4.json 2> $OUT/${T}_bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg4.json
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 3 > $OUT/${T}_bench_cfg2.json 2> $OUT/${T}_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg2.json
