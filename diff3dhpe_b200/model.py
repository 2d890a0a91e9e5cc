"""Drop-in for the reference's MixSTE seq2seq denoiser module.

Same constructor signature, attribute names and state-dict keys as
`common/nets/model_conditional_diffusion_mixste_s2s_grand_linLift.py:139-220`, so a reference checkpoint loads
unchanged (`load_state_dict(..., strict=False)` as in RUN:226-235) and `torch.manual_seed(s)` followed by
construction yields the same initial weights (sub-modules are created in the reference's order).  The
sub-modules are parameter containers only: `forward_denoise` (MODEL:249-257) is executed by
libdiff3d_b200.so through `Engine`.  Training is outside the scope of this package.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .engine import Engine


class _AttentionParams(nn.Module):
    """Parameter holder with the keys of `Attention` (MODEL:59-71): qkv.{weight,bias}, proj.{weight,bias}."""

    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    """Keys of `Mlp` (MODEL:40-48): fc1, fc2."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _BlockParams(nn.Module):
    """Keys of `Block` (MODEL:92-109): norm1, attn.*, norm2, time_mlp.1.*, mlp.*  (creation order matters for
    seed-for-seed identical initialisation)."""

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer, time_emb_dim):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _AttentionParams(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim)) if time_emb_dim is not None else None
        self.mlp = _MlpParams(dim, int(dim * mlp_ratio))


class ConditionalDiffusionMixSTES2SGRANDLinLift(nn.Module):
    def __init__(self, num_frame=9, num_joints=17, in_chans=2, embed_dim=32, depth=4,
                 num_heads=8, mlp_ratio=2., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None, with_time_emb=True, **kwargs):
        super().__init__()
        if qk_scale is not None:
            raise ValueError("qk_scale override is not supported by the B200 kernels (head_dim**-0.5 is baked in)")
        if in_chans != 2:
            raise ValueError("in_chans must be 2 (2D keypoints)")
        self.num_frame, self.num_joints, self.embed_dim = num_frame, num_joints, embed_dim
        self.num_heads, self.mlp_ratio = num_heads, mlp_ratio
        self.block_depth = depth
        self.with_time_emb = bool(with_time_emb)

        if with_time_emb:
            time_dim = embed_dim * 2
            # index 0 of the reference Sequential is the parameter-free SinusoidalPosEmb (MODEL:166-174)
            self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(embed_dim, time_dim), nn.GELU(),
                                          nn.Linear(time_dim, time_dim))
        else:
            time_dim = None
            self.time_mlp = None
        self.fusion_layer = nn.Linear(3 + in_chans, embed_dim)
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim))
        self.STEblocks = nn.ModuleList([
            _BlockParams(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer, time_dim) for _ in range(depth)])
        self.Spatial_norm = norm_layer(embed_dim)
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, embed_dim))
        self.TTEblocks = nn.ModuleList([
            _BlockParams(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer, time_dim) for _ in range(depth)])
        self.Temporal_norm = norm_layer(embed_dim)
        self.head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, 3))

        # engine state (not part of the state dict)
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._weights_key = None
        self._weights_epoch = 0
        self.gemm_mode = _lib.GEMM_DEFAULT
        self.attn_mode = _lib.ATTN_DEFAULT
        self.use_graph = True
        self.max_clips_hint = 1

    # ------------------------------------------------------------------ engine management
    def _weights_fingerprint(self):
        return (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_weights(self):
        """Forces a re-upload (re-packing) of the weights at the next call.  The engine notices parameter updates through
        `(data_ptr, _version)` of every parameter, which covers optimizer steps, `load_state_dict`, `copy_`, `.to()`
        and in-place ops on the parameter itself -- but NOT writes through `.data` (`p.data.copy_(ema)`, EMA / SWA
        weight swaps), which do not bump `_version`: call this after such a write.  (A content checksum would need a
        device -> host synchronisation on every sampler call, which the hot path must not have.)"""
        self._weights_epoch += 1

    def _load_from_state_dict(self, *args, **kwargs):
        self._weights_epoch += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def engine(self, batch: int = 1) -> Engine:
        """Returns the Engine for the device the parameters live on, (re)building it when the device, the
        kernel modes or the required capacity changed, and re-uploading weights when any parameter changed."""
        dev = self.fusion_layer.weight.device
        if dev.type != "cuda":
            raise RuntimeError("diff3dhpe_b200 runs on CUDA (sm_100a) only: move the module to a B200 with .cuda(); "
                               "there is no CPU fallback")
        need = max(batch, self.max_clips_hint)
        key = (dev, self.gemm_mode, self.attn_mode, self.use_graph)
        if self._engine is None or self._engine_key != key or self._engine.max_clips < need:
            if self._engine is not None:
                self._engine.close()
            mlp_hidden = int(self.embed_dim * self.mlp_ratio)
            self._engine = Engine(self.num_frame, self.num_joints, self.embed_dim, self.block_depth, self.num_heads,
                                  mlp_hidden, self.with_time_emb, need, dev, self.gemm_mode, self.attn_mode,
                                  self.use_graph)
            self._engine_key = key
            self._weights_key = None
        fp = self._weights_fingerprint()
        if self._weights_key != fp:
            self._engine.load_state_dict({k: v for k, v in self.state_dict().items()})
            self._weights_key = fp
        return self._engine

    # ------------------------------------------------------------------ reference API
    def forward_denoise(self, x, time):
        """MODEL:249-257.  x: [B,F,J,5] fp32, time: [B] -> [B,F,J,3]."""
        if self.training:
            raise NotImplementedError("diff3dhpe_b200 implements the inference path only: call .eval() first "
                                      "(training stays with the reference implementation)")
        assert x.dim() == 4, "shape is equal to 4"
        b, f, n, _ = x.shape
        if f != self.num_frame or n != self.num_joints:
            raise ValueError(f"expected [B,{self.num_frame},{self.num_joints},5], got {tuple(x.shape)}")
        eng = self.engine(b)
        return eng.forward_denoise(x.detach().to(torch.float32).contiguous(), time)

    def forward(self, x, time):
        return self.forward_denoise(x, time)
