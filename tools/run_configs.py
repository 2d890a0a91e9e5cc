#!/usr/bin/env python
"""Throughput + sanity of the BASELINE.json configurations other than the bench line (cfg3 is bench.py).

    python tools/run_configs.py cfg2 cfg4 cfg5                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_configs.py cfg5                                                # clips sharded over N ranks

cfg2: h36m_cpn shape, F=81, 9 DDIM steps, 256 clips on one GPU (no flip pass: BASELINE cfg2 as stated).
cfg4: 3dhp_gt shape, F=27, no time embedding (Experiments.sh:17), flip-TTA with the MPI-INF-3DHP joint lists,
      sampling_timesteps sweep 1/9/25/50 at 2048 and 512 clips (the launch / CUDA-graph bound regime).
cfg5: full-test-set-sized synthetic sweep: 240 sequences x 2250 frames = 540 000 frames, F=243 windows with the
      reference's windowing rule (last window back-shifted and masked, nosiy_generators.py:27-48) = 2400 windows,
      flip-TTA, sharded contiguously over the ranks; NCCL all-gather of the predictions + all-reduce of the MPJPE
      pair inside the timed region.  `--cfg5-sequences` shrinks it for smoke runs.
Prints one JSON line per measurement on rank 0.  Everything goes through the public API (evaluate.evaluate_shard ->
GaussianDiffusion.ddim_sample_loop -> C ABI); inputs start in pinned host memory.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from diff3dhpe_b200 import evaluate, synthetic  # noqa: E402

J = 17


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def sync(world, dev):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def emit(rank, **kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def run_eval(diff, x2d_h, gt_h, mask_h, *, dev, batch, tta, left, right, n_total, world, reps):
    """Times `reps` sweeps of this rank's shard through evaluate_shard + the single exchange step."""
    sampler = evaluate.DeviceSampler(diff)
    F = x2d_h.shape[1]

    def noise_fn(ids, flip):
        return diff.draw_noise([len(ids), F, J, 3], dev)

    def once():
        res = evaluate.evaluate_shard(sampler, x2d_h, gt_h, noise_fn, device=dev, batch_clips=batch, tta=tta, left=left,
                                      right=right, frame_mask=mask_h)
        return evaluate.gather_results(res["pred"], res["acc"], n_total)

    pred, mp = once()                       # warm-up: graph capture, weight upload
    sync(world, dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        pred, mp = once()
    sync(world, dev)
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    assert torch.isfinite(pred).all()
    return dt, mp, pred


def cfg2(args, world, rank, dev):
    F, B, S = 81, 256, 9
    model = synthetic.make_model(F).to(dev)
    model.max_clips_hint = B
    diff = synthetic.make_diffusion(model, sampling_timesteps=S).to(dev).eval()
    x2d, gt = synthetic.make_inputs(B, F, seed=20 + rank)
    dt, mp, _ = run_eval(diff, x2d.pin_memory(), gt.pin_memory(), None, dev=dev, batch=B, tta=False,
                         left=synthetic.H36M_JOINTS_LEFT, right=synthetic.H36M_JOINTS_RIGHT, n_total=B * world, world=1,
                         reps=args.reps)
    emit(rank, config="cfg2", F=F, clips=B, S=S, tta=False, n_gpus=1, seconds=dt, pose_frames_per_s=B * F / dt,
         mpjpe_vs_synthetic_gt=mp)
    model._engine.close()


def cfg4(args, world, rank, dev):
    F = 27
    L, R = synthetic.MPI3DHP_JOINTS_LEFT, synthetic.MPI3DHP_JOINTS_RIGHT
    for B in (2048, 512):
        model = synthetic.make_model(F, with_time_emb=False).to(dev)
        model.max_clips_hint = 2 * B
        for S in (1, 9, 25, 50):
            diff = synthetic.make_diffusion(model, sampling_timesteps=S).to(dev).eval()
            x2d, gt = synthetic.make_inputs(B, F, seed=40 + rank)
            dt, mp, _ = run_eval(diff, x2d.pin_memory(), gt.pin_memory(), None, dev=dev, batch=B, tta=True, left=L, right=R,
                                 n_total=B, world=1, reps=args.reps)
            emit(rank, config="cfg4", F=F, clips=B, S=S, tta=True, with_time_emb=False, n_gpus=1, seconds=dt,
                 pose_frames_per_s=B * F / dt, ms_per_ddim_step=1000 * dt / S, mpjpe_vs_synthetic_gt=mp)
        model._engine.close()


def cfg5_raw(args, world, rank, dev):
    """cfg5 from RAW packed sequences: windowing, flipping and the masked write-back run on the device (N3); sequences
    (not windows) are sharded contiguously over the ranks."""
    F, S = 243, 9
    n_seq, seq_len = args.cfg5_sequences, 2250
    assert n_seq % world == 0, "raw-sequence mode shards whole sequences: --cfg5-sequences must divide by the world size"
    start, count = evaluate.shard_range(n_seq, rank, world)
    g = torch.Generator().manual_seed(2000 + rank)
    seq2d = (0.3 * torch.randn(count * seq_len, J, 2, generator=g)).clamp_(-1, 1).pin_memory()
    gt = 0.3 * torch.randn(count * seq_len, J, 3, generator=g)
    gt = (gt - gt[:, :1]).pin_memory()
    model = synthetic.make_model(F).to(dev)
    model.max_clips_hint = 2 * args.batch
    diff = synthetic.make_diffusion(model, sampling_timesteps=S).to(dev).eval()
    sampler = evaluate.DeviceSampler(diff)

    def noise_fn(ids, flip):
        return diff.draw_noise([len(ids), F, J, 3], dev)

    def once():
        res = evaluate.evaluate_sequences(sampler, seq2d, gt, [seq_len] * count, noise_fn, device=dev, F=F,
                                          batch_clips=args.batch)
        pred, mp = evaluate.gather_results(res["pred"], res["acc"], n_seq * seq_len)      # gathered in units of frames
        return pred, mp, res["n_windows"]

    pred, mp, n_win = once()
    sync(world, dev)
    t0 = time.perf_counter()
    pred, mp, n_win = once()
    sync(world, dev)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    frames = n_seq * seq_len
    emit(rank, config="cfg5_raw_sequences", F=F, sequences=n_seq, frames=frames, windows_per_rank=n_win, S=S, tta=True,
         n_gpus=world, seconds=dt, pose_frames_per_s=frames / dt, mpjpe_vs_synthetic_gt=mp,
         gathered_shape=list(pred.shape), windowing="device (d3d_window_gather / d3d_window_scatter)")
    model._engine.close()


def cfg5(args, world, rank, dev):
    if args.raw_sequences:
        return cfg5_raw(args, world, rank, dev)
    F, S = 243, 9
    n_seq, seq_len = args.cfg5_sequences, 2250
    wins = evaluate.window_starts(seq_len, F)                      # 10 windows per sequence, last one back-shifted
    n_total = n_seq * len(wins)
    start, count = evaluate.shard_range(n_total, rank, world)
    # synthetic sequences are generated per window (seeded by the global window id) so every rank builds only its shard
    g = torch.Generator().manual_seed(1000 + rank)
    x2d = (0.3 * torch.randn(count, F, J, 2, generator=g)).clamp_(-1, 1)
    gt = 0.3 * torch.randn(count, F, J, 3, generator=g)
    gt = gt - gt[:, :, :1]
    mask = torch.ones(count, F, dtype=torch.uint8)
    for i in range(count):
        first_valid = wins[(start + i) % len(wins)][1]
        mask[i, :first_valid] = 0                                   # frames already predicted by the previous window
    model = synthetic.make_model(F).to(dev)
    model.max_clips_hint = 2 * args.batch
    diff = synthetic.make_diffusion(model, sampling_timesteps=S).to(dev).eval()
    dt, mp, pred = run_eval(diff, x2d.pin_memory(), gt.pin_memory(), mask.pin_memory(), dev=dev, batch=args.batch, tta=True,
                            left=synthetic.H36M_JOINTS_LEFT, right=synthetic.H36M_JOINTS_RIGHT, n_total=n_total,
                            world=world, reps=1)
    valid_frames = n_seq * seq_len
    emit(rank, config="cfg5", F=F, windows=n_total, frames=valid_frames, S=S, tta=True, n_gpus=world, seconds=dt,
         pose_frames_per_s=valid_frames / dt, window_frames_per_s=n_total * F / dt, mpjpe_vs_synthetic_gt=mp,
         gathered_shape=list(pred.shape), exchange="all_gather_into_tensor(pred) + all_reduce(sum err, count), timed")
    model._engine.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=["cfg2", "cfg4", "cfg5"])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=256, help="clips per sampler batch for cfg5 (x2 with the flip copies)")
    ap.add_argument("--cfg5-sequences", type=int, default=240)
    ap.add_argument("--raw-sequences", action="store_true",
                    help="cfg5 from raw packed sequences: windowing / flipping / write-back on the device")
    args = ap.parse_args()
    world, rank, dev = setup()
    for c in args.configs:
        if c != "cfg5" and rank != 0:
            continue                                                # cfg2 / cfg4 are single-GPU configurations
        {"cfg2": cfg2, "cfg4": cfg4, "cfg5": cfg5}[c](args, world, rank, dev)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
