"""GEMM kernel microbenchmark through the C ABI (d3d_op_linear_bench): ms / launch and TFLOP/s per shape, tile and mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff3dhpe_b200 import _lib  # noqa: E402
from diff3dhpe_b200.engine import Engine  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1057536
eng = Engine(27, max_clips=1)
print(f"M={M}")
for mode, name, passes in ((_lib.GEMM_TC_SPLIT3, "split3", 3), (_lib.GEMM_TC_FP16, "fp16", 1)):
    for N, K, act in ((1536, 512, 0), (512, 512, 0), (1024, 512, 1), (512, 1024, 0)):
        for bn in (128, 256):
            os.environ["D3D_GEMM_BN"] = str(bn)
            ms = eng.op_linear_bench(M, N, K, act, mode, iters=5)
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            print(f"{name:6s} N={N:4d} K={K:4d} act={act} BN={bn}: {ms:8.3f} ms  {tf:7.1f} TF/s algorithmic  {tf * passes:7.1f} executed",
                  flush=True)
