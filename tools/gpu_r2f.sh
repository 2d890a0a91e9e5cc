#!/bin/bash
# round-2 session F: fc1 epilogue register A/B, qkv EW A/B, ncu full captures of the non-GEMM kernels, memcheck of smoke()
set -u
T=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
for lib in libdiff3d_b200 libd3d_regs112; do
  D3D_LIB=$PWD/diff3dhpe_b200/$lib.so timeout 300 python tools/gemm_mode_bench.py 2115072 f4c >> $OUT/${T}_gemm_regs_ab.log 2>&1
done
D3D_GEMM_EW_QKV=16 timeout 300 python tools/gemm_mode_bench.py 2115072 f4c >> $OUT/${T}_gemm_regs_ab.log 2>&1
cat $OUT/${T}_gemm_regs_ab.log
D3D_LIB=$PWD/diff3dhpe_b200/libd3d_regs112.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_regs112.json 2> $OUT/${T}_bench_regs112.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_regs112.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_base.json 2> $OUT/${T}_bench_base.err; echo "bench rc=$?"; cut -c1-200 $OUT/${T}_bench_base.json
for k in attn_temporal postnorm_add_ln ln_split; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 2 -o $OUT/${T}_full_$k -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${T}_full_$k.log 2>&1; echo "full $k rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${T}_memcheck_smoke.log 2>&1; tail -4 $OUT/${T}_memcheck_smoke.log
