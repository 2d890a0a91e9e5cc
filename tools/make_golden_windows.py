"""Golden vectors for the windowing row: drives the UNMODIFIED reference ChunkedGenerator
(common/nosiy_generators.py, numpy only) on three synthetic sequences, checks the oracle restatement against it
bit for bit and writes the reference's windows to tests/golden/windows_f9.npz.

    python tools/make_golden_windows.py          (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from common.nosiy_generators import ChunkedGenerator  # noqa: E402
from oracle import diff3d_oracle as oracle  # noqa: E402

F, J = 9, 17
LENS = [9, 20, 31]
L, R = oracle.H36M_JOINTS_LEFT, oracle.H36M_JOINTS_RIGHT


def main():
    rng = np.random.RandomState(7)
    keys = [("S9", "Walk", c) for c in range(len(LENS))]
    poses_2d = {k: rng.randn(n, J, 2).astype(np.float32) for k, n in zip(keys, LENS)}
    poses_3d = {k: rng.randn(n, J, 3).astype(np.float32) for k, n in zip(keys, LENS)}
    frame_id = {k: np.arange(n) for k, n in zip(keys, LENS)}
    gen = ChunkedGenerator(4, None, poses_3d, poses_2d, frame_id, chunk_length=F, pad=0, shuffle=False, augment=False,
                           kps_left=L, kps_right=R, joints_left=L, joints_right=R, out_all=True)
    x2d, x2d_flip, masks, starts, seq_ids = [], [], [], [], []
    base = {k: sum(LENS[:i]) for i, k in enumerate(keys)}
    for seq_name, s3, e3, st3, et3, flip, reverse in gen.pairs:
        key = (seq_name[0], seq_name[1], int(seq_name[2]))
        _, _, b2d, mask, _, _, _, _, _ = gen.get_batch_seq2seq(seq_name, s3, e3, st3, False, False)
        _, _, b2d_f, _, _, _, _, _, _ = gen.get_batch_seq2seq(seq_name, s3, e3, st3, True, False)
        # the oracle restatement must reproduce the reference bit for bit
        sc, stg = oracle.chunk_windows(LENS[key[2]], F)
        w = sc.index(int(s3))
        assert stg[w] == int(st3), (stg, st3)
        ob, om = oracle.window_batch(torch.from_numpy(poses_2d[key]), int(s3), F, int(st3), False)
        of, _ = oracle.window_batch(torch.from_numpy(poses_2d[key]), int(s3), F, int(st3), True)
        assert np.array_equal(ob.numpy(), b2d) and np.array_equal(of.numpy(), b2d_f) and np.array_equal(om.numpy(), mask)
        x2d.append(b2d), x2d_flip.append(b2d_f), masks.append(mask)
        starts.append(base[key] + int(s3)), seq_ids.append(key[2])
    out = os.path.join(ROOT, "tests", "golden", "windows_f9.npz")
    np.savez_compressed(out, seq2d=np.concatenate([poses_2d[k] for k in keys]), lens=np.array(LENS), F=F,
                        x2d=np.stack(x2d), x2d_flip=np.stack(x2d_flip), mask=np.stack(masks),
                        win_start=np.array(starts, dtype=np.int64), seq_id=np.array(seq_ids, dtype=np.int32))
    print(out, "windows:", len(starts), "starts:", starts, "first_valid:", [int((~m).sum()) for m in masks])


if __name__ == "__main__":
    main()
