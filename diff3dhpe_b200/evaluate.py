"""Sharded flip-TTA evaluation: the device-side equivalent of the reference's `evaluate()` inner loop
(RUN:557-606) with `nn.DataParallel` (RUN:216-218) replaced by one process per GPU.

Clips are independent, so rank r of R owns a contiguous block of clips (both flip variants of a clip stay on
the same rank); nothing is communicated inside the sampler.  One exchange step at the end: gather of the merged
predictions and an all-reduce of the fp64 (sum of joint errors, joint count) pair -- NCCL on GPUs, gloo in the
CPU tests of this plumbing.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import synthetic


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: returns (start, count) of rank's clips; counts differ by at most one."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def window_starts(n_frames: int, F: int) -> Sequence[Tuple[int, int]]:
    """Windowing rule of ChunkedGenerator (common/nosiy_generators.py:27-48) for one sequence of n_frames with
    stride == F: non-overlapping windows, the last one shifted back to end at the sequence end.  Returns
    (start_frame, first_valid_offset) per window: frames before first_valid_offset were already predicted by the
    previous window and are masked out (`target_mask`)."""
    if n_frames < F:
        raise ValueError("sequence shorter than one window")
    out = []
    n_full = n_frames // F
    for w in range(n_full):
        out.append((w * F, 0))
    if n_frames % F:
        start = n_frames - F
        out.append((start, n_full * F - start))
    return out


def plan_windows(seq_lengths: Sequence[int], F: int):
    """Window plan of a set of sequences packed back to back: (win_start int64 [n_win] = first PACKED frame of every
    window, first_valid int32 [n_win], seq_id int32 [n_win]) following `window_starts` per sequence, in the
    generator's order (sequence by sequence, windows in time order; nosiy_generators.py:27-48)."""
    starts, valid, sid = [], [], []
    base = 0
    for s, n in enumerate(seq_lengths):
        for st, fv in window_starts(int(n), F):
            starts.append(base + st)
            valid.append(fv)
            sid.append(s)
        base += int(n)
    return (torch.tensor(starts, dtype=torch.int64), torch.tensor(valid, dtype=torch.int32),
            torch.tensor(sid, dtype=torch.int32))


class DeviceSampler:
    """Adapter: (x2d [n,F,J,2] device, y_T, step_noise) -> y0, running GaussianDiffusion.ddim_sample_loop."""

    def __init__(self, diffusion):
        self.diffusion = diffusion

    def __call__(self, x2d, y_T, step_noise):
        return self.diffusion.ddim_sample_loop(x2d, list(y_T.shape), noise=(y_T, step_noise))

    def merge(self, y, y_flip, left, right, scale):
        return self.diffusion.model.engine(y.shape[0]).tta_merge(y, y_flip, left, right, scale)

    def mpjpe(self, pred, gt, acc, mask):
        return self.diffusion.model.engine(1).mpjpe_accumulate(pred, gt, acc, mask)


def evaluate_shard(sampler, x2d: torch.Tensor, gt: Optional[torch.Tensor], noise_fn: Callable, *, device,
                   batch_clips: int = 256, tta: bool = True, left=synthetic.H36M_JOINTS_LEFT,
                   right=synthetic.H36M_JOINTS_RIGHT, scale: float = 1.0, frame_mask: Optional[torch.Tensor] = None,
                   clip_offset: int = 0) -> Dict[str, torch.Tensor]:
    """Runs this rank's clips.  x2d/gt/frame_mask are HOST tensors of the local shard ([n,F,J,2], [n,F,J,3],
    [n,F] uint8); `noise_fn(global_clip_ids, flip) -> (y_T, step_noise|None)` returns device noise.
    Returns {'pred': [n,F,J,3] (device), 'acc': fp64[2] (device)}."""
    n, F, J, _ = x2d.shape
    pred = torch.empty((n, F, J, 3), device=device, dtype=torch.float32)
    acc = torch.zeros(2, device=device, dtype=torch.float64)
    for s in range(0, n, batch_clips):
        e = min(n, s + batch_clips)
        ids = torch.arange(clip_offset + s, clip_offset + e)
        xb = x2d[s:e].to(device, non_blocking=True)
        y_T, sn = noise_fn(ids, False)
        if tta:
            # orig || flip as one 2B batch (clips are independent), then un-flip + average (RUN:583-588)
            xf = synthetic.flip_2d(xb, left, right)
            yf_T, snf = noise_fn(ids, True)
            y_all = sampler(torch.cat([xb, xf]), torch.cat([y_T, yf_T]),
                            None if sn is None else torch.cat([sn, snf], dim=1))
            b = e - s
            out = sampler.merge(y_all[:b].contiguous(), y_all[b:].contiguous(), left, right, scale)
        else:
            out = sampler(xb, y_T, sn)
            if scale != 1.0:
                out = out * scale
        pred[s:e] = out
        if gt is not None:
            gb = gt[s:e].to(device, non_blocking=True)
            mb = None if frame_mask is None else frame_mask[s:e].to(device).reshape(-1).contiguous()
            sampler.mpjpe(out.contiguous(), gb.contiguous(), acc, mb)
    return {"pred": pred, "acc": acc}


def evaluate_sequences(sampler, seq2d: torch.Tensor, gt3d: Optional[torch.Tensor], seq_lengths: Sequence[int],
                       noise_fn: Callable, *, device, F: int, batch_clips: int = 256,
                       left=synthetic.H36M_JOINTS_LEFT, right=synthetic.H36M_JOINTS_RIGHT, scale: float = 1.0,
                       window_offset: int = 0) -> Dict[str, torch.Tensor]:
    """The evaluate() inner loop from RAW packed sequences: seq2d [N, J, 2] / gt3d [N, J, 3] are HOST
    tensors holding this rank's sequences back to back.  The F-frame windows and their flipped copies
    (nosiy_generators.py:27-48, 264-276) are built on the device, sampled with flip-TTA, merged (RUN:583-588) and written
    back into packed frame order with the overlap of each back-shifted last window masked (RUN:589-596).
    Returns {'pred': [N, J, 3] (device), 'acc': fp64[2] (device), 'n_windows': int}."""
    ws, fv, _ = plan_windows(seq_lengths, F)
    n_win = ws.numel()
    N, J, _ = seq2d.shape
    eng = sampler.diffusion.model.engine(2 * min(batch_clips, n_win))
    seq_d = seq2d.to(device, non_blocking=True)
    ws_d, fv_d = ws.to(device), fv.to(device)
    out = torch.zeros((N, J, 3), device=device, dtype=torch.float32)
    for s in range(0, n_win, batch_clips):
        e = min(n_win, s + batch_clips)
        ids = torch.arange(window_offset + s, window_offset + e)
        xb, xf = eng.window_gather(seq_d, ws_d[s:e].contiguous(), left, right)
        y_T, sn = noise_fn(ids, False)
        yf_T, snf = noise_fn(ids, True)
        y_all = sampler(torch.cat([xb, xf]), torch.cat([y_T, yf_T]), None if sn is None else torch.cat([sn, snf], dim=1))
        b = e - s
        merged = sampler.merge(y_all[:b].contiguous(), y_all[b:].contiguous(), left, right, scale)
        eng.window_scatter(merged, ws_d[s:e].contiguous(), fv_d[s:e].contiguous(), out)
    acc = torch.zeros(2, device=device, dtype=torch.float64)
    if gt3d is not None:
        sampler.mpjpe(out, gt3d.to(device, non_blocking=True).contiguous(), acc, None)
    return {"pred": out, "acc": acc, "n_windows": n_win}


def gather_results(local_pred: torch.Tensor, acc: torch.Tensor, n_total: int, group=None):
    """The single exchange step: all-gather of the (padded) prediction shards + all-reduce(sum) of the fp64
    (error sum, joint count) pair.  Returns (pred [n_total,F,J,3] on every rank, mpjpe float)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        cnt = acc[1].item()
        return local_pred, (acc[0].item() / cnt if cnt else float("nan"))
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_pred.shape[1:]), device=local_pred.device, dtype=local_pred.dtype)
    pad[: local_pred.shape[0]] = local_pred
    bucket = torch.empty((world * per,) + tuple(local_pred.shape[1:]), device=local_pred.device,
                         dtype=local_pred.dtype)
    dist.all_gather_into_tensor(bucket, pad, group=group) if bucket.is_cuda else dist.all_gather(
        list(bucket.view(world, per, *local_pred.shape[1:]).unbind(0)), pad, group=group)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    parts = []
    for r in range(world):
        _, c = shard_range(n_total, r, world)
        parts.append(bucket[r * per: r * per + c])
    cnt = acc[1].item()
    return torch.cat(parts), (acc[0].item() / cnt if cnt else float("nan"))
