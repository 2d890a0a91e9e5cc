mkdir -p gpurun_out
T=r01v
timeout 300 python -m pytest tests/test_gpu_ops.py -k "attention" -x -q > gpurun_out/${T}_pytest_attn.log 2>&1; tail -2 gpurun_out/${T}_pytest_attn.log
timeout 300 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/${T}_pytest_sampler.log 2>&1; tail -2 gpurun_out/${T}_pytest_sampler.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench.json"))
    print(round(d["value"]), d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["per_class_ms"])
except Exception as e:
    print("bench failed", e)
PY
