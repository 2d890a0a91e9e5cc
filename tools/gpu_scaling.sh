#!/bin/bash
# Scaling exhibit: cfg5 strong scaling (2400 windows, fixed total) and cfg3 weak scaling at the rank
# counts the box has.  usage: tools/gpu_scaling.sh TAG "1 2 4"     (run under gpurun --gpus N)
set -u
T=${1:-r02k}
NS=${2:-"1 2 4"}
OUT=gpurun_out
mkdir -p $OUT
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --config cfg5 --gpus 1 --steps 1 --warmup 1 --lean --no-cpu-baseline > $OUT/${T}_cfg5_n1.json 2> $OUT/${T}_cfg5_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --config cfg5 --gpus $n --steps 1 --warmup 1 --lean --no-cpu-baseline > $OUT/${T}_cfg5_n$n.json 2> $OUT/${T}_cfg5_n$n.err
  fi
  echo "cfg5 n=$n rc=$?"; cut -c1-260 $OUT/${T}_cfg5_n$n.json
done
nmax=$(echo $NS | awk '{print $NF}')
if [ "$nmax" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $nmax --master-addr 127.0.0.1 --master-port 29400 \
    bench.py --gpus $nmax --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_cfg3_weak_n$nmax.json 2> $OUT/${T}_cfg3_weak_n$nmax.err
  echo "cfg3 weak n=$nmax rc=$?"; cut -c1-260 $OUT/${T}_cfg3_weak_n$nmax.json
fi
