"""CPU oracle for the Diff3DHPE DDIM / MixSTE-s2s sampler hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diff3dhpe_b200/`` imports this file; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and
only as the checker / the CPU baseline, never as the product path.

It is a *restatement* (functional, state-dict driven, fp32 torch-CPU arithmetic) of the reference's
algorithm, not an import of it.  Citations are ``file:line`` into the reference tree:

  MODEL = common/nets/model_conditional_diffusion_mixste_s2s_grand_linLift.py
  DIFF  = common/conditional_diffusion_ddim_normal_directPredict_variableLoss_both_crossFrames.py
  RUN   = run_conditionalDiffusionDDIM3dhpeNormalDirectPredictVariableLoss.py
  LOSS  = common/loss.py
  GEN   = common/nosiy_generators.py   (sic)

Parity pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against the *imported, unmodified reference* run in the build container:
``tools/make_golden.py`` runs the reference (with a 6-line ``timm.DropPath`` stub) and this oracle on the
same seeds, asserts they are bit-identical, and writes the reference's outputs to ``tests/golden/``;
``tests/test_oracle.py`` re-checks the oracle against those committed vectors on every run.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# H36M-17 and MPI-INF-3DHP-17 left/right joint lists (common/h36m_dataset.py:18-21,288 after the 32->17
# reduction; common/mpiinf3dhp_dataset.py:17-18).  Used by the flip-TTA tail (RUN:562-565, 583-585).
H36M_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]
H36M_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]
MPI3DHP_JOINTS_LEFT = [5, 6, 7, 11, 12, 13]
MPI3DHP_JOINTS_RIGHT = [2, 3, 4, 8, 9, 10]


def strip_prefix(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Accept GaussianDiffusion / DataParallel checkpoints: drop 'module.' and 'model.' prefixes and the
    schedule buffers (RUN:226-235 drops every key containing 'alphas'; all 14 buffers are rebuilt)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.startswith("model."):
            k = k[len("model."):]
        elif "." not in k:          # schedule buffers live at top level of GaussianDiffusion
            continue
        out[k] = v
    return out


# --------------------------------------------------------------------------------------------------
# Schedule (DIFF:58-68, 119-183, 270-273)
# --------------------------------------------------------------------------------------------------
def cosine_betas(timesteps: int, s: float = 0.008) -> Tensor:
    """DIFF:58-68, fp64."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def linear_betas(timesteps: int) -> Tensor:
    """DIFF:52-55."""
    return torch.linspace(0.0001, 0.02, timesteps, dtype=torch.float64)


def schedule_buffers(timesteps: int, beta_schedule: str = "cosine") -> Dict[str, Tensor]:
    """fp64 cumprod then cast to fp32 (DIFF:134-136, 151-163).  Only the two buffers the DDIM loop reads
    (plus sqrt_alphas_cumprod for q_sample) are produced."""
    if beta_schedule == "cosine":
        betas = cosine_betas(timesteps)
    elif beta_schedule == "linear":
        betas = linear_betas(timesteps)
    else:
        raise ValueError(f"unknown beta schedule {beta_schedule}")
    ac = torch.cumprod(1.0 - betas, dim=0)
    return {
        "alphas_cumprod": ac.to(torch.float32),
        "sqrt_alphas_cumprod": torch.sqrt(ac).to(torch.float32),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac).to(torch.float32),
    }


def ddim_times(timesteps: int, sampling_timesteps: int) -> List[int]:
    """DIFF:270-272: linspace(-1, T-1, S+1).int() reversed.  S=9,T=1000 ->
    [999, 887, 776, 665, 554, 443, 332, 221, 110, -1]."""
    t = torch.linspace(-1, timesteps - 1, steps=sampling_timesteps + 1)
    return list(reversed(t.int().tolist()))


def ddim_coefficients(bufs: Dict[str, Tensor], times: Sequence[int], eta: float) -> List[Optional[Dict[str, Tensor]]]:
    """Per-step fp32 scalars of DIFF:287-297, computed with the same fp32 tensor ops the reference uses.
    Entry is None for the final step (time_next < 0, DIFF:283-285)."""
    out = []
    for t, tn in zip(times[:-1], times[1:]):
        if tn < 0:
            out.append(None)
            continue
        alpha = bufs["alphas_cumprod"][t]
        alpha_next = bufs["alphas_cumprod"][tn]
        sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        out.append({
            "alpha": alpha, "sqrt_alpha_next": alpha_next.sqrt(), "c": c, "sigma": sigma,
            "sqrt_one_minus": bufs["sqrt_one_minus_alphas_cumprod"][t],
        })
    return out


# --------------------------------------------------------------------------------------------------
# Denoiser (MODEL:24-257)
# --------------------------------------------------------------------------------------------------
def sinusoidal_embedding(time: Tensor, dim: int) -> Tensor:
    """MODEL:29-36."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half) * -step)
    arg = time[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def time_mlp(sd: Dict[str, Tensor], time: Tensor, dim: int) -> Tensor:
    """Top-level time MLP, MODEL:169-174: sinusoid -> Linear -> GELU(erf) -> Linear."""
    e = sinusoidal_embedding(time, dim)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def attention_core(qkv: Tensor, heads: int) -> Tensor:
    """MODEL:75-83 without the linears: qkv [Bn, N, 3C] (channel = which*C + head*hd + d) -> [Bn, N, C]."""
    Bn, N, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    qkv = qkv.reshape(Bn, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    p = ((q @ k.transpose(-2, -1)) * (hd ** -0.5)).softmax(dim=-1)
    eye = torch.eye(N, dtype=p.dtype).view(1, 1, N, N).repeat(Bn, heads, 1, 1)
    return ((p - eye) @ v).transpose(1, 2).reshape(Bn, N, C)


def attention(sd: Dict[str, Tensor], pre: str, x: Tensor, heads: int) -> Tensor:
    """GRAND attention, MODEL:73-86.  x: [Bn, N, C]."""
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
    o = attention_core(qkv, heads)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp(sd: Dict[str, Tensor], pre: str, x: Tensor) -> Tensor:
    """MODEL:50-56 (dropout p=0)."""
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def block(sd: Dict[str, Tensor], pre: str, x: Tensor, spatial: bool, t_emb: Optional[Tensor], heads: int,
          eps: float = 1e-6) -> Tensor:
    """MODEL:111-135, eval path.  x: [B,F,J,C]."""
    b, f, j, c = x.shape
    if t_emb is not None and (pre + "time_mlp.1.weight") in sd:
        tv = F.linear(F.silu(t_emb), sd[pre + "time_mlp.1.weight"], sd[pre + "time_mlp.1.bias"])
        x = x + tv[:, None, None, :]
    if spatial:
        x = x.reshape(b * f, j, c)
    else:
        x = x.permute(0, 2, 1, 3).reshape(b * j, f, c)
    x = x + attention(sd, pre + "attn.", F.layer_norm(x, (c,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps), heads)
    x = x + mlp(sd, pre + "mlp.", F.layer_norm(x, (c,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    if spatial:
        return x.reshape(b, f, j, c)
    return x.reshape(b, j, f, c).permute(0, 2, 1, 3)


def forward_denoise(sd: Dict[str, Tensor], x5: Tensor, time: Tensor, heads: int = 8) -> Tensor:
    """MODEL:249-257 with ST_foward (MODEL:222-247) inlined.  x5: [B,F,J,5] -> [B,F,J,3]."""
    sd = strip_prefix(sd) if any(k.startswith(("module.", "model.")) for k in sd) else sd
    x = F.linear(x5, sd["fusion_layer.weight"], sd["fusion_layer.bias"])
    c = x.shape[-1]
    t_emb = time_mlp(sd, time, c) if "time_mlp.1.weight" in sd else None
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("STEblocks."))
    for i in range(depth):
        if i == 0:
            x = x + sd["Spatial_pos_embed"]                         # [1,J,C] broadcast over (B,F), MODEL:230-233
        x = block(sd, f"STEblocks.{i}.", x, True, t_emb, heads)
        x = F.layer_norm(x, (c,), sd["Spatial_norm.weight"], sd["Spatial_norm.bias"], 1e-6)
        if i == 0:
            x = x + sd["Temporal_pos_embed"].unsqueeze(2)           # [1,F,1,C], MODEL:239-242
        x = block(sd, f"TTEblocks.{i}.", x, False, t_emb, heads)
        x = F.layer_norm(x, (c,), sd["Temporal_norm.weight"], sd["Temporal_norm.bias"], 1e-6)
    x = F.layer_norm(x, (c,), sd["head.0.weight"], sd["head.0.bias"], 1e-5)       # MODEL:218 default eps
    return F.linear(x, sd["head.1.weight"], sd["head.1.bias"])


# --------------------------------------------------------------------------------------------------
# DDIM loop (DIFF:251-300) with explicit noise
# --------------------------------------------------------------------------------------------------
def draw_noise(shape: Sequence[int], sampling_timesteps: int, generator: Optional[torch.Generator] = None
               ) -> Tuple[Tensor, Tensor]:
    """Reproduce the reference's draw order for one sampler call: one randn for y_T (DIFF:275), then one
    randn_like per non-final step (DIFF:293) -- drawn even when eta == 0."""
    y_T = torch.randn(tuple(shape), generator=generator)
    steps = [torch.randn(tuple(shape), generator=generator) for _ in range(sampling_timesteps - 1)]
    step_noise = torch.stack(steps) if steps else torch.zeros((0, *shape))
    return y_T, step_noise


def ddim_sample_loop(sd: Dict[str, Tensor], x2d: Tensor, y_T: Tensor, step_noise: Optional[Tensor], *,
                     timesteps: int = 1000, sampling_timesteps: int = 9, eta: float = 0.0,
                     clip_denoised: bool = True, beta_schedule: str = "cosine", heads: int = 8,
                     trace: bool = False):
    """DIFF:263-300 (and 304-347 when trace=True).  Returns y0 [B,F,J,3] (and the y_t / x_start stacks)."""
    sd = strip_prefix(sd) if any(k.startswith(("module.", "model.")) for k in sd) else sd
    bufs = schedule_buffers(timesteps, beta_schedule)
    times = ddim_times(timesteps, sampling_timesteps)
    coefs = ddim_coefficients(bufs, times, eta)
    y = y_T
    ys, x0s = [], []
    B = y.shape[0]
    for step, (t, co) in enumerate(zip(times[:-1], coefs)):
        tt = torch.full((B,), t, dtype=torch.long)
        x0 = forward_denoise(sd, torch.cat([x2d, y], dim=-1), tt, heads)        # DIFF:254-255
        if clip_denoised:
            x0 = torch.clamp(x0, min=-1.0, max=1.0)                              # DIFF:252,256
        x0s.append(x0)
        if co is None:
            y = x0
        else:
            noise = step_noise[step] if step_noise is not None else torch.zeros_like(y)
            # literal DIFF:295-297 (note: alpha * x_start, not sqrt(alpha))
            y = x0 * co["sqrt_alpha_next"] + co["c"] * ((y - co["alpha"] * x0) / co["sqrt_one_minus"]) + co["sigma"] * noise
        ys.append(y)
    if trace:
        return y, torch.stack(ys, dim=-1), torch.stack(x0s, dim=-1)
    return y


# --------------------------------------------------------------------------------------------------
# forward() eval branch with output_loss / repeat_n (DIFF:360-366, 392-419, 421-449)
# --------------------------------------------------------------------------------------------------
def q_sample(bufs: Dict[str, Tensor], x_start: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """DIFF:360-366: sqrt(abar_t) x0 + sqrt(1 - abar_t) noise with per-sample t (``extract`` = gather + reshape)."""
    shape = (t.shape[0],) + (1,) * (x_start.dim() - 1)
    return (bufs["sqrt_alphas_cumprod"].gather(-1, t).reshape(shape) * x_start +
            bufs["sqrt_one_minus_alphas_cumprod"].gather(-1, t).reshape(shape) * noise)


def p_losses(sd: Dict[str, Tensor], x_start: Tensor, pose_2d: Tensor, t: Tensor, noise: Tensor, *,
             timesteps: int = 1000, beta_schedule: str = "cosine", loss_type: str = "l2", clip_loss: bool = True,
             heads: int = 8) -> Tensor:
    """DIFF:392-419 with the two random draws made explicit: ``t`` is the reference's
    ``torch.randint(0, T, (b,))`` and ``noise`` its ``randn_like(x_start)`` (drawn in that order).  One denoiser call
    with a per-sample t, then the un-reduced l1 / l2 error times ``1 + abar_t / sqrt(1 - abar_t)`` (clamped to 3 when
    ``clipLoss``).  Returns [B,F,J,3]."""
    sd = strip_prefix(sd) if any(k.startswith(("module.", "model.")) for k in sd) else sd
    bufs = schedule_buffers(timesteps, beta_schedule)
    x_noisy = q_sample(bufs, x_start, t, noise)
    model_out = forward_denoise(sd, torch.cat([pose_2d, x_noisy], dim=-1), t, heads)
    coef = 1.0 + bufs["alphas_cumprod"][t].view(-1, 1, 1, 1) / bufs["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1)
    if clip_loss:
        coef = torch.clamp(coef, max=3.0)
    if loss_type == "l1":
        err = F.l1_loss(model_out, x_start, reduction="none")
    elif loss_type == "l2":
        err = F.mse_loss(model_out, x_start, reduction="none")
    else:
        raise ValueError(f"invalid loss type {loss_type}")
    return err * coef


def forward_eval(sd: Dict[str, Tensor], clean_3d_pose: Tensor, noisy_2d_pose: Tensor, y_T: Tensor,
                 step_noise: Optional[Tensor], *, repeat_n: int = 1, loss_draws: Optional[Tuple[Tensor, Tensor]] = None,
                 **kw):
    """GaussianDiffusion.forward, eval branch (DIFF:427-449).  ``loss_draws = (t, noise)`` selects ``output_loss=True``
    (the default of the 3DHP evaluate(), RUN3:517-520); the 2D input is tiled ``repeat_n`` times along the batch
    (DIFF:434), the sampler runs on the ``repeat_n * B`` clips (``y_T`` / ``step_noise`` have that batch) and the
    prediction is the mean over the repeats (DIFF:448).  Returns (loss or None, pred [B,F,J,3])."""
    loss_kw = {k: kw.pop(k) for k in ("loss_type", "clip_loss") if k in kw}
    loss = None
    if loss_draws is not None:
        loss = p_losses(sd, clean_3d_pose, noisy_2d_pose, loss_draws[0], loss_draws[1],
                        timesteps=kw.get("timesteps", 1000), beta_schedule=kw.get("beta_schedule", "cosine"), **loss_kw)
    b, f, p, _ = clean_3d_pose.shape
    x = noisy_2d_pose.repeat(repeat_n, 1, 1, 1)
    pred = ddim_sample_loop(sd, x, y_T, step_noise, **kw)
    pred = torch.mean(pred.view(repeat_n, b, f, p, -1), dim=0, keepdim=True).squeeze(0)
    return loss, pred


# --------------------------------------------------------------------------------------------------
# flip-TTA tail and MPJPE (RUN:562-590, LOSS:15-27)
# --------------------------------------------------------------------------------------------------
def flip_2d(x2d: Tensor, left=H36M_JOINTS_LEFT, right=H36M_JOINTS_RIGHT) -> Tensor:
    """2D (or 3D) horizontal flip: negate x, swap L/R joints (common/nosiy_generators.py:273-276; RUN:562-565)."""
    o = x2d.clone()
    o[..., 0] *= -1
    o[:, :, left + right] = o[:, :, right + left]
    return o


def tta_merge(pred: Tensor, pred_flip: Tensor, scale: float = 1.0, left=H36M_JOINTS_LEFT,
              right=H36M_JOINTS_RIGHT) -> Tensor:
    """RUN:583-588: un-flip the flipped prediction, average, undo the 3D normalisation (x scale)."""
    pf = pred_flip.clone()
    pf[:, :, :, 0] *= -1
    pf[:, :, left + right] = pf[:, :, right + left]
    return ((pred + pf) / 2.0) * scale


def mpjpe(pred: Tensor, target: Tensor) -> Tensor:
    """LOSS:15-27, reduce='mean'."""
    return torch.mean(torch.norm(pred - target, dim=-1))


def sample_tta(sd, x2d, noise_pair, flip_noise_pair, **kw) -> Tensor:
    """evaluate()'s two sampler calls + merge (RUN:577-588) at scale 1."""
    scale = kw.pop("scale", 1.0)
    y = ddim_sample_loop(sd, x2d, *noise_pair, **kw)
    yf = ddim_sample_loop(sd, flip_2d(x2d), *flip_noise_pair, **kw)
    return tta_merge(y, yf, scale)


# ----------------------------------------------------------------------------------------------- windowing (N3)
def chunk_windows(n_seq_frame: int, chunk_length: int) -> Tuple[List[int], List[int]]:
    """Window bounds of one sequence for the seq2seq (`out_all`) generator, GEN:27-40: non-overlapping chunks, the
    last one shifted back to end at the sequence end.  Returns (start_index_chunk, start_index_chunk_target); the
    first `start - start_target` frames of a window are masked out of the targets (GEN:264-271)."""
    n_chunks = (n_seq_frame + chunk_length - 1) // chunk_length
    bounds = [c * chunk_length for c in range(n_chunks)]
    start_last = n_seq_frame - chunk_length
    target_offset = start_last - bounds[-1]
    start_chunk = bounds[:-1] + [start_last]
    start_target = bounds[:-1] + [start_last + target_offset]
    return start_chunk, start_target


def window_batch(seq_2d: Tensor, start: int, chunk_length: int, start_target: int, flip: bool,
                 kps_left=H36M_JOINTS_LEFT, kps_right=H36M_JOINTS_RIGHT) -> Tuple[Tensor, Tensor]:
    """One window of the generator (GEN:247-276, pad = causal_shift = 0, no edge padding: start >= 0): the 2D slice
    (flipped: x negated, left/right keypoints swapped) and its target_mask."""
    assert start >= 0 and start + chunk_length <= seq_2d.shape[0]
    batch_2d = seq_2d[start:start + chunk_length].clone()
    target_mask = torch.ones(chunk_length, dtype=torch.bool)
    n_unused = start - start_target
    assert n_unused >= 0
    if n_unused > 0:
        target_mask[:n_unused] = False
    if flip:
        batch_2d[:, :, 0] *= -1
        batch_2d[:, list(kps_left) + list(kps_right)] = batch_2d[:, list(kps_right) + list(kps_left)]
    return batch_2d, target_mask


# ----------------------------------------------------------------------------------------------- metrics (N4)
def n_mpjpe(predicted: Tensor, target: Tensor) -> Tensor:
    """Protocol #3, LOSS:84-94: per-frame least-squares scale, then MPJPE.  Inputs [N, 1, J, 3]."""
    norm_predicted = torch.mean(torch.sum(predicted ** 2, dim=3, keepdim=True), dim=2, keepdim=True)
    norm_target = torch.mean(torch.sum(target * predicted, dim=3, keepdim=True), dim=2, keepdim=True)
    return mpjpe(norm_target / norm_predicted * predicted, target)


def p_mpjpe(predicted, target) -> float:
    """Protocol #2, LOSS:43-82 (numpy, float64 if the inputs are): Procrustes alignment per frame.  Inputs [N, J, 3]."""
    import numpy as np
    predicted, target = np.asarray(predicted), np.asarray(target)
    muX = np.mean(target, axis=1, keepdims=True)
    muY = np.mean(predicted, axis=1, keepdims=True)
    X0, Y0 = target - muX, predicted - muY
    normX = np.sqrt(np.sum(X0 ** 2, axis=(1, 2), keepdims=True))
    normY = np.sqrt(np.sum(Y0 ** 2, axis=(1, 2), keepdims=True))
    X0, Y0 = X0 / normX, Y0 / normY
    H = np.matmul(X0.transpose(0, 2, 1), Y0)
    U, s, Vt = np.linalg.svd(H)
    V = Vt.transpose(0, 2, 1)
    R = np.matmul(V, U.transpose(0, 2, 1))
    sign_detR = np.sign(np.expand_dims(np.linalg.det(R), axis=1))
    V[:, :, -1] *= sign_detR
    s[:, -1] *= sign_detR.flatten()
    R = np.matmul(V, U.transpose(0, 2, 1))
    tr = np.expand_dims(np.sum(s, axis=1, keepdims=True), axis=2)
    a = tr * normX / normY
    t = muX - a * np.matmul(muY, R)
    aligned = a * np.matmul(predicted, R) + t
    return float(np.mean(np.linalg.norm(aligned - target, axis=2)))


def mean_velocity_error(predicted, target) -> float:
    """LOSS:133-142: mean norm of the difference of the first temporal differences.  Inputs [N, J, 3]."""
    import numpy as np
    vp, vt = np.diff(np.asarray(predicted), axis=0), np.diff(np.asarray(target), axis=0)
    return float(np.mean(np.linalg.norm(vp - vt, axis=2)))
