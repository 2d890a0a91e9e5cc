#!/bin/bash
# round-2 session W: three TMEM slots in the single-tile attention modes
set -u
T=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" > $OUT/${T}_pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -4 $OUT/${T}_pytest_attn.log
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -x -k "alternative" > $OUT/${T}_pytest_alt.log 2>&1; echo "pytest alt rc=$?"; tail -4 $OUT/${T}_pytest_alt.log
for v in "slots3:D3D_ATTN_SLOTS=3" "slots2:D3D_ATTN_SLOTS=2" "slots3_b:D3D_ATTN_SLOTS=3"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_$name.json 2> $OUT/${T}_bench_$name.err; echo "bench $name rc=$?"; cut -c1-200 $OUT/${T}_bench_$name.json; tail -2 $OUT/${T}_bench_$name.err
done
for v in "slots3:D3D_ATTN_SLOTS=3" "slots2:D3D_ATTN_SLOTS=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_cfg4_$name.json 2> $OUT/${T}_bench_cfg4_$name.err; echo "cfg4 $name rc=$?"; cut -c1-200 $OUT/${T}_bench_cfg4_$name.json
done
