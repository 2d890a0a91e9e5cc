mkdir -p gpurun_out
T=r01r
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench.json"))
    print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["clocks"], d["roofline"]["per_class_ms"], d["roofline"]["frac"], d["cpu_baseline"])
except Exception as e:
    print("bench failed", e)
PY
timeout 600 python tools/run_configs.py cfg2 cfg4 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; echo "configs rc=$?"; cut -c1-250 gpurun_out/${T}_configs.jsonl
timeout 400 python tools/run_configs.py cfg5 --raw-sequences --cfg5-sequences 48 > gpurun_out/${T}_cfg5_raw48.jsonl 2> gpurun_out/${T}_cfg5_raw48.err; echo "cfg5 raw rc=$?"; cut -c1-400 gpurun_out/${T}_cfg5_raw48.jsonl
timeout 400 python tools/run_configs.py cfg5 --cfg5-sequences 48 > gpurun_out/${T}_cfg5_48.jsonl 2> gpurun_out/${T}_cfg5_48.err; echo "cfg5 rc=$?"; cut -c1-400 gpurun_out/${T}_cfg5_48.jsonl
