mkdir -p gpurun_out
T=r01q
timeout 200 python -m pytest tests/test_gpu_ops.py -k "linear or gelu or matches" -x -q > gpurun_out/${T}_pytest_gemm_default.log 2>&1; tail -2 gpurun_out/${T}_pytest_gemm_default.log
D3D_GEMM_N_INNER=1 timeout 400 python -m pytest tests/test_gpu_ops.py -k "linear or gelu or matches" -x -q > gpurun_out/${T}_pytest_gemm_ninner.log 2>&1; tail -2 gpurun_out/${T}_pytest_gemm_ninner.log
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
run_bench ninner1 D3D_GEMM_N_INNER=1
run_bench ninner0 D3D_GEMM_N_INNER=0
D3D_GEMM_N_INNER=1 timeout 400 ncu --set full --clock-control none -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/${T}_full_gemm_ninner1 -f python bench.py --steps 1 --warmup 3 --clips 128 --no-cpu-baseline > gpurun_out/${T}_full_gemm_ninner1.log 2>&1
echo "ncu rc=$?"
