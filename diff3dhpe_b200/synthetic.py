"""Synthetic workload of SURVEY.md section 8(d): random-init weights of the reference architecture and
synthetic 2D/3D sequences (there is no network access for the licensed datasets or the checkpoints).
Everything is generated on the CPU with seeded torch generators, so it is identical on every machine."""
from __future__ import annotations

import torch

from .diffusion import GaussianDiffusion
from .model import ConditionalDiffusionMixSTES2SGRANDLinLift

H36M_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]      # common/h36m_dataset.py:18-21,288 after the 32->17 reduction
H36M_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]
MPI3DHP_JOINTS_LEFT = [5, 6, 7, 11, 12, 13]   # common/mpiinf3dhp_dataset.py:17-18
MPI3DHP_JOINTS_RIGHT = [2, 3, 4, 8, 9, 10]


def make_model(num_frame: int, depth: int = 8, with_time_emb: bool = True, seed: int = 0, embed_dim: int = 512):
    """Reference default init under torch.manual_seed(seed) (construction order matches the reference, so the
    weights equal the reference's under the same seed), plus pos-embeds ~ N(0, 0.02^2): the reference leaves
    them at zero (MODEL:193,205), which would not exercise the pos-embed path."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = ConditionalDiffusionMixSTES2SGRANDLinLift(num_frame=num_frame, num_joints=17, in_chans=2, embed_dim=embed_dim,
                                                  depth=depth, num_heads=8, mlp_ratio=2., qkv_bias=True,
                                                  drop_path_rate=0.1, with_time_emb=with_time_emb)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        m.Spatial_pos_embed.copy_(0.02 * torch.randn(m.Spatial_pos_embed.shape, generator=g))
        m.Temporal_pos_embed.copy_(0.02 * torch.randn(m.Temporal_pos_embed.shape, generator=g))
    torch.random.set_rng_state(state)
    return m.eval()


def make_diffusion(model, sampling_timesteps: int = 9, eta: float = 0.0, clip_denoised: bool = True,
                   timesteps: int = 1000):
    """The GaussianDiffusion of RUN:188-189 with the shipped JSON settings (cosine, T=1000, l2)."""
    return GaussianDiffusion(model, timesteps=timesteps, sampling_timesteps=sampling_timesteps, loss_type='l2',
                             clip_denoised=clip_denoised, beta_schedule='cosine', ddim_sampling_eta=eta,
                             clipLoss=True).eval()


def make_inputs(B: int, F: int, seed: int = 1234, J: int = 17):
    """2D keypoints ~ 0.3 N(0,1) clipped to [-1,1]; GT 3D ~ 0.3 N(0,1), root-centred (joint 0 = 0)."""
    g = torch.Generator().manual_seed(seed)
    x2d = (0.3 * torch.randn(B, F, J, 2, generator=g)).clamp_(-1, 1)
    gt = 0.3 * torch.randn(B, F, J, 3, generator=g)
    gt = gt - gt[:, :, :1]
    return x2d, gt


def make_noise(B: int, F: int, S: int, seed: int = 4321, J: int = 17):
    """Explicit DDIM noise in the reference's draw order (y_T then S-1 step tensors)."""
    g = torch.Generator().manual_seed(seed)
    y_T = torch.randn(B, F, J, 3, generator=g)
    steps = torch.stack([torch.randn(B, F, J, 3, generator=g) for _ in range(S - 1)]) if S > 1 else torch.zeros(0, B, F, J, 3)
    return y_T, steps


def flip_2d(x, left=H36M_JOINTS_LEFT, right=H36M_JOINTS_RIGHT):
    """Horizontal flip of a keypoint / pose tensor [B,F,J,*]: negate x, swap L/R (nosiy_generators.py:273-276)."""
    o = x.clone()
    o[..., 0] *= -1
    o[:, :, left + right] = o[:, :, right + left]
    return o
