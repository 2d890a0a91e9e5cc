#!/bin/bash
# round-2 session P: two rows per warp in the LayerNorm kernels; tc2 attention v2 (unmasked chunks, prefetch behind the barrier)
set -u
T=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log; tail -8 $OUT/${T}_pytest.log
for v in "wg2_1_rows2:D3D_ATTN_WG2=1" "wg2_0_rows2:D3D_ATTN_WG2=0" "wg2_0_rows1:D3D_ATTN_WG2=0 D3D_LN_ROWS=1" "wg2_0_rows2_b:D3D_ATTN_WG2=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json; tail -2 $OUT/${T}_bench_$name.err
done
