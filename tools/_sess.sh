mkdir -p gpurun_out
T=r01p
STAGES=attn_sp_tc_f243,attn_sp_tc_f27_split16,attn_tc_f81 timeout 300 python tools/gpu_first_contact.py > gpurun_out/${T}_contact.log 2>&1; cut -c1-200 gpurun_out/${T}_contact.log
timeout 300 python -m pytest tests/test_gpu_ops.py -k "attention" -x -q > gpurun_out/${T}_pytest_attn.log 2>&1; tail -3 gpurun_out/${T}_pytest_attn.log
timeout 300 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/${T}_pytest_sampler.log 2>&1; tail -3 gpurun_out/${T}_pytest_sampler.log
run_bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_$name.json"))
    print("$name", round(d["value"]), d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["per_class_ms"])
except Exception as e:
    print("bench $name failed", e)
PY
}
run_bench default D3D_X=0
run_bench ewqkv16 D3D_GEMM_EW_QKV=16
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_temporal_tc -s 20 -c 2 -o gpurun_out/${T}_full_attn_tc -f python bench.py --steps 1 --warmup 3 --clips 128 --no-cpu-baseline > gpurun_out/${T}_full_attn_tc.log 2>&1
echo "ncu rc=$?"
