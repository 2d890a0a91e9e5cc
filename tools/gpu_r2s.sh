#!/bin/bash
# round-2 session S: is the operand feed limited per SM or chip-wide?  the same GEMMs on 148 / 112 / 74 SMs
set -u
T=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
for sms in 148 112 74 36; do
  echo "SMS=$sms" >> $OUT/${T}_gemm_sms.log
  D3D_GEMM_SMS=$sms timeout 300 python tools/gemm_epi_bench.py >> $OUT/${T}_gemm_sms.log 2>&1
done
cat $OUT/${T}_gemm_sms.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_ew8.json 2> $OUT/${T}_bench_ew8.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_ew8.json
D3D_GEMM_EW_GELU=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_ew16.json 2> $OUT/${T}_bench_ew16.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_ew16.json
