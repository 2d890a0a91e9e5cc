mkdir -p gpurun_out
T=r01t
timeout 200 python -m pytest tests/test_gpu_ops.py -k "metrics or window" -x -q > gpurun_out/${T}_pytest_n4.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/${T}_pytest_n4.log | cut -c1-300
