"""Prints the sampler-level parity numbers of every GEMM mode against the committed reference goldens (GPU box):
per-joint max-abs error and |MPJPE delta| for each golden case.   python tools/parity_report.py [mode ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diff3dhpe_b200 import _lib, synthetic  # noqa: E402
from oracle import diff3d_oracle as oracle  # noqa: E402  (checker only)

MODES = {"f8c": _lib.GEMM_TC_F8C, "f4c": _lib.GEMM_TC_F4C, "split3": _lib.GEMM_TC_SPLIT3, "fp16": _lib.GEMM_TC_FP16}
CASES = ["sampler_f27_b2_s3_clip", "sampler_f27_b2_s2_notime", "sampler_f81_b1_s2_noclip", "sampler_f243_b1_s1_clip",
         "sampler_f9_b2_s9_clip", "sampler_f81_b1_s9_clip", "sampler_f243_b1_s9_clip", "sampler_f27_b2_s9_notime"]
modes = sys.argv[1:] or ["f8c", "f4c"]
for name in CASES:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    F, B, S = int(g["F"]), int(g["B"]), int(g["S"])
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, S)
    ref = torch.from_numpy(g["pred"])
    row = []
    for m in modes:
        model = synthetic.make_model(F, with_time_emb=bool(g["with_time_emb"])).cuda()
        model.gemm_mode, model.max_clips_hint = MODES[m], B
        diff = synthetic.make_diffusion(model, sampling_timesteps=S, clip_denoised=bool(g["clip"])).cuda().eval()
        pred = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=(y_T.cuda(), None)).cpu()
        model._engine.close()
        err = (pred - ref).abs().max().item()
        dm = abs(oracle.mpjpe(pred, gt).item() - oracle.mpjpe(ref, gt).item())
        row.append(f"{m}: {err:.2e} / {dm:.1e}")
    print(f"{name:28s} " + "   ".join(row), flush=True)
