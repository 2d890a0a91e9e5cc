mkdir -p gpurun_out
T=r01x
timeout 300 python -m pytest tests/test_gpu_ops.py -k "residual_stream or linear" -x -q > gpurun_out/${T}_pytest_ops.log 2>&1; tail -2 gpurun_out/${T}_pytest_ops.log
timeout 400 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/${T}_pytest_sampler.log 2>&1; tail -2 gpurun_out/${T}_pytest_sampler.log
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
run_bench defer1 D3D_DEFER_POSTNORM=1
run_bench defer0 D3D_DEFER_POSTNORM=0
run_bench defer1b D3D_DEFER_POSTNORM=1
