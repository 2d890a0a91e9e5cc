"""Golden values for the metric row: the UNMODIFIED reference common/loss.py functions on seeded
poses (including mirrored frames, which exercise the reflection branch of p_mpjpe), checked against the oracle
restatement and written to tests/golden/metrics.npz.

    python tools/make_golden_metrics.py          (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from common.loss import mean_velocity_error, mpjpe, n_mpjpe, p_mpjpe  # noqa: E402
from oracle import diff3d_oracle as oracle  # noqa: E402


def main():
    rng = np.random.RandomState(11)
    N, J = 300, 17
    gt = (0.3 * rng.randn(N, J, 3)).astype(np.float32)
    gt -= gt[:, :1]
    pred = (gt + 0.05 * rng.randn(N, J, 3)).astype(np.float32)
    pred[:7, :, 0] *= -1                                   # mirrored predictions: det(R) < 0 before the fix
    tp, tg = torch.from_numpy(pred).unsqueeze(1), torch.from_numpy(gt).unsqueeze(1)
    ref = dict(mpjpe=mpjpe(tp, tg).item(), n_mpjpe=n_mpjpe(tp, tg).item(), p_mpjpe=float(p_mpjpe(pred.copy(), gt.copy())),
               velocity=float(mean_velocity_error(pred, gt)))
    ours = dict(mpjpe=oracle.mpjpe(tp, tg).item(), n_mpjpe=oracle.n_mpjpe(tp, tg).item(), p_mpjpe=oracle.p_mpjpe(pred, gt),
                velocity=oracle.mean_velocity_error(pred, gt))
    assert ref == ours, (ref, ours)
    out = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(out, pred=pred, gt=gt, **{k: np.float64(v) for k, v in ref.items()})
    print(out, ref)


if __name__ == "__main__":
    main()
