#!/bin/bash
# round-2 session O: temporal attention with two softmax warpgroups per slot (attn_temporal_tc2_kernel)
set -u
T=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" > $OUT/${T}_pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -15 $OUT/${T}_pytest_attn.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
for v in "wg2_1:D3D_ATTN_WG2=1" "wg2_0:D3D_ATTN_WG2=0" "wg2_1b:D3D_ATTN_WG2=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json; tail -2 $OUT/${T}_bench_$name.err
done
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_cfg2.json 2> $OUT/${T}_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg2.json
