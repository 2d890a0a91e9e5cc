mkdir -p gpurun_out
T=r01u
timeout 200 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${T}_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck_smoke.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench.json"))
    print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["clocks"], d["roofline"]["per_class_ms"], d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("bench failed", e)
PY
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6200 -c 2100 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/${T}_full_gemm_fullsize -f python bench.py --steps 1 --warmup 3 --clips 256 --no-cpu-baseline > gpurun_out/${T}_full_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out | tail -12
