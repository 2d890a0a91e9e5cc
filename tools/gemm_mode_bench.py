"""GEMM microbenchmark of the four MixSTE linears at the bench's row count through the C ABI (d3d_op_linear_bench):
ms per launch, algorithmic TFLOP/s and fraction of the measured sustained bf16 peak, per precision mode.

    python tools/gemm_mode_bench.py [M] [mode ...]        modes: f8c f4c split3 fp16
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diff3dhpe_b200 import _lib  # noqa: E402
from diff3dhpe_b200.engine import Engine  # noqa: E402

args = sys.argv[1:]
M = int(args[0]) if args and args[0].isdigit() else 2115072
modes = [a for a in args if not a.isdigit()] or ["f8c"]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
eng = Engine(27, max_clips=1)
ids = {"f8c": _lib.GEMM_TC_F8C, "split3": _lib.GEMM_TC_SPLIT3, "fp16": _lib.GEMM_TC_FP16}
if hasattr(_lib, "GEMM_TC_F4C"):
    ids["f4c"] = _lib.GEMM_TC_F4C
print(f"M={M} lib={_lib.LIB_PATH}")
for name in modes:
    tot_ms, tot_fl = 0.0, 0.0
    for label, N, K, act in (("qkv", 1536, 512, 0), ("proj", 512, 512, 0), ("fc1", 1024, 512, 1), ("fc2", 512, 1024, 0)):
        ms = eng.op_linear_bench(M, N, K, act, ids[name], iters=5)
        fl = 2.0 * M * N * K
        tot_ms += ms
        tot_fl += fl
        print(f"{name:6s} {label:4s} N={N:4d} K={K:4d}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TF/s algorithmic = {fl / ms / 1e9 / peak:.3f} of {peak}",
              flush=True)
    print(f"{name:6s} block: {tot_ms:7.3f} ms  {tot_fl / tot_ms / 1e9:7.1f} TF/s algorithmic = {tot_fl / tot_ms / 1e9 / peak:.3f}", flush=True)
