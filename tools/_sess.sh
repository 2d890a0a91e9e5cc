mkdir -p gpurun_out
T=r01m
if timeout 180 python -m pytest tests/test_gpu_ops.py -k "gelu or f8c_tc_matches or tc_matches" -x -q > gpurun_out/${T}_pytest_gemm.log 2>&1; then echo "EW16 gemm tests OK"; else echo "EW16 gemm tests FAILED -> EW=8"; export D3D_GEMM_EW_GELU=8; fi
tail -3 gpurun_out/${T}_pytest_gemm.log
timeout 300 python -m pytest tests/test_gpu_ops.py -k "attention" -x -q > gpurun_out/${T}_pytest_attn.log 2>&1; tail -3 gpurun_out/${T}_pytest_attn.log
timeout 300 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/${T}_pytest_sampler.log 2>&1; tail -3 gpurun_out/${T}_pytest_sampler.log
for slots in 2 1; do
  D3D_ATTN_TC_SLOTS=$slots timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_slots$slots.json 2> gpurun_out/${T}_bench_slots$slots.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_slots$slots.json"))
    print("slots=$slots ew_gelu=${D3D_GEMM_EW_GELU:-16}", round(d["value"]), d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["per_class_ms"])
except Exception as e:
    print("bench slots=$slots failed", e)
PY
done
for k in attn_temporal_tc; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o gpurun_out/${T}_full_$k -f python bench.py --steps 1 --warmup 3 --clips 128 --no-cpu-baseline > gpurun_out/${T}_full_$k.log 2>&1
  echo "ncu $k rc=$?"; tail -3 gpurun_out/${T}_full_$k.log
done
