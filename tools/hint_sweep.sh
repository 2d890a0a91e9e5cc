#!/bin/bash
# L2 policy sweep of the GEMM (bench.py per-class GEMM time under each setting).  Usage: bash tools/hint_sweep.sh <tag>
TAG=${1:-sweep}
mkdir -p gpurun_out
for cfg in "0 0 0" "0 0 1" "1 0 1" "1 1 1" "2 1 1"; do
  set -- $cfg
  echo "== hint_a=$1 hint_b=$2 stream_out=$3"
  D3D_GEMM_HINT_A=$1 D3D_GEMM_HINT_B=$2 D3D_GEMM_STREAM_OUT=$3 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']), d['clocks']['sm_mhz'], d['roofline']['per_class_ms'])"
done 2>&1 | tee gpurun_out/${TAG}_hint_sweep.log
