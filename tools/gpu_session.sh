#!/bin/bash
# One GPU-box session: parity tests, a bench line, the ncu launch list of the bench command and full captures
# of the dominant kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_session.sh <tag> [what...]
#   what: tests bench launches full   (default: all)
set -u
TAG=${1:-r01}
shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.csv 2>&1

for w in $WHAT; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
      echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
      tail -3 $OUT/${TAG}_pytest.log
      ;;
    bench)
      timeout 900 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.json
      ;;
    launches)
      # every kernel launch of ONE timed step of the bench command (warm-up launches skipped)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-6200} -c ${NCU_COUNT:-2100} \
        --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} \
        > $OUT/${TAG}_launches.log 2>&1
      echo "launches rc=$?"
      ;;
    full)
      for k in ${NCU_KERNELS:-gemm_tc attn_temporal attn_spatial postnorm_add_ln}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 2 \
          -o $OUT/${TAG}_full_$k -f python bench.py --steps 1 --warmup 3 --clips ${NCU_CLIPS:-64} --no-cpu-baseline ${BENCH_ARGS:-} \
          > $OUT/${TAG}_full_$k.log 2>&1
        echo "full $k rc=$?"
      done
      ;;
  esac
done
ls -la $OUT | tail -20
