"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on the
synthetic workload, and pins the oracle to it: every case asserts oracle == reference bit-for-bit on this
machine before the reference's output is written.  Run in the build container only:

    python tools/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_import import load_reference  # noqa: E402
from diff3dhpe_b200 import synthetic  # noqa: E402
from oracle import diff3d_oracle as oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
RefModel, RefDiffusion = load_reference()
torch.set_num_threads(os.cpu_count())


def ref_model_from(mine, F, with_time_emb):
    ref = RefModel(num_frame=F, num_joints=17, in_chans=2, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.,
                   qkv_bias=True, drop_path_rate=0.1, with_time_emb=with_time_emb)
    missing = ref.load_state_dict(mine.state_dict(), strict=True)
    return ref.eval()


class FixedNoise:
    """Feeds the explicit noise tensors to the reference's torch.randn / randn_like call sites in order."""

    def __init__(self, y_T, steps):
        self.queue = [y_T] + [s for s in steps]

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: self.queue.pop(0).clone()
        torch.randn_like = lambda *a, **k: self.queue.pop(0).clone()
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def sampler_case(name, F, B, S, eta, clip, with_time_emb=True, trace=False):
    mine = synthetic.make_model(F, with_time_emb=with_time_emb)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    ref = ref_model_from(mine, F, with_time_emb)
    diff = RefDiffusion(ref, timesteps=1000, sampling_timesteps=S, loss_type='l2', clip_denoised=clip,
                        beta_schedule='cosine', ddim_sampling_eta=eta, clipLoss=True).eval()
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, steps = synthetic.make_noise(B, F, S)
    with torch.no_grad(), FixedNoise(y_T, steps) as fn:
        if trace:
            _, pred, rev, x0s = diff(clean_3d_pose=gt, noisy_2d_pose=x2d, output_loss=False, output_reverse_diffusion_3d=True)
        else:
            _, pred = diff(clean_3d_pose=gt, noisy_2d_pose=x2d, output_loss=False)
        assert len(fn.queue) == 0, "reference consumed a different number of normal draws than S"
    with torch.no_grad():
        o = oracle.ddim_sample_loop(sd, x2d, y_T, steps, timesteps=1000, sampling_timesteps=S, eta=eta,
                                    clip_denoised=clip, trace=trace)
    if trace:
        o, orev, ox0 = o
        assert torch.equal(orev, rev) and torch.equal(ox0, x0s), f"{name}: oracle trace != reference"
    d = (o - pred).abs().max().item()
    print(f"{name}: |oracle - reference|max = {d:.3e}  |pred|max = {pred.abs().max():.3f}")
    assert d == 0.0, f"{name}: oracle is not bit-identical to the reference"
    out = dict(pred=pred.numpy(), F=F, B=B, S=S, eta=eta, clip=int(clip), with_time_emb=int(with_time_emb))
    if trace:
        out.update(rev=rev.numpy(), x0s=x0s.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def denoise_case(name, F, B, tvals, with_time_emb=True):
    mine = synthetic.make_model(F, with_time_emb=with_time_emb)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    ref = ref_model_from(mine, F, with_time_emb)
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, 1)
    x5 = torch.cat([x2d, y_T], dim=-1)
    t = torch.tensor(tvals, dtype=torch.long)
    with torch.no_grad():
        out = ref.forward_denoise(x5, t)
        o = oracle.forward_denoise(sd, x5, t)
        d = (o - out).abs().max().item()
        print(f"{name}: |oracle - reference|max = {d:.3e}")
        assert d == 0.0
        # residual stream after the first STE block / first TTE block (before the following post-norm),
        # sub-sampled every 8th token, for kernel-level localisation
        x = ref.fusion_layer(x5)
        temb = ref.time_mlp(t) if with_time_emb else None
        x = x + ref.Spatial_pos_embed
        x1 = ref.STEblocks[0](x, True, temb)
        x = ref.Spatial_norm(x1)
        x = x + ref.Temporal_pos_embed.unsqueeze(2)
        x2 = ref.TTEblocks[0](x, False, temb)
        tt = None
        if with_time_emb:
            tt = torch.stack([blk.time_mlp(temb) for pair in zip(ref.STEblocks, ref.TTEblocks) for blk in pair], dim=1)
    extra = dict(x_after_1=x1.reshape(-1, 512)[::8].numpy(), x_after_2=x2.reshape(-1, 512)[::8].numpy())
    if tt is not None:
        extra["time_table"] = tt.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), out=out.numpy(), t=np.array(tvals), F=F, B=B, **extra)


class FixedRandint:
    """Feeds a fixed t tensor to the reference's torch.randint call in p_losses (DIFF:395)."""

    def __init__(self, t):
        self.t, self.calls = t, 0

    def __enter__(self):
        self._randint = torch.randint

        def fake(*a, **k):
            self.calls += 1
            return self.t.clone()
        torch.randint = fake
        return self

    def __exit__(self, *exc):
        torch.randint = self._randint


def forward_case(name, F, B, S, repeat_n, tvals, loss_type="l2", clip_loss=True, with_time_emb=True):
    """GaussianDiffusion.forward in eval mode with output_loss=True (the 3DHP evaluate() default, RUN3:517-520) and
    repeat_n (DIFF:434,448): the reference's loss tensor and prediction with every random draw pinned."""
    mine = synthetic.make_model(F, with_time_emb=with_time_emb)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    ref = ref_model_from(mine, F, with_time_emb)
    diff = RefDiffusion(ref, timesteps=1000, sampling_timesteps=S, loss_type=loss_type, clip_denoised=True,
                        beta_schedule='cosine', ddim_sampling_eta=0.0, clipLoss=clip_loss).eval()
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, steps = synthetic.make_noise(repeat_n * B, F, S)
    g = torch.Generator().manual_seed(777)
    loss_noise = torch.randn(B, F, 17, 3, generator=g)
    t = torch.tensor(tvals, dtype=torch.long)
    # draw order inside forward(): randint (DIFF:395), randn_like (DIFF:397), then the sampler's S draws
    with torch.no_grad(), FixedRandint(t) as fr, FixedNoise(y_T, steps) as fn:
        fn.queue.insert(0, loss_noise)
        loss, pred = diff(clean_3d_pose=gt, noisy_2d_pose=x2d, repeat_n=repeat_n)
        assert fr.calls == 1 and len(fn.queue) == 0
    with torch.no_grad():
        oloss, opred = oracle.forward_eval(sd, gt, x2d, y_T, steps, repeat_n=repeat_n, loss_draws=(t, loss_noise),
                                           sampling_timesteps=S, loss_type=loss_type, clip_loss=clip_loss)
    assert torch.equal(oloss, loss) and torch.equal(opred, pred), f"{name}: oracle forward() != reference"
    print(f"{name}: oracle == reference; mean loss {loss.mean():.5f}, |pred|max {pred.abs().max():.3f}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), loss=loss.numpy(), pred=pred.numpy(), t=np.array(tvals), F=F,
                        B=B, S=S, repeat_n=repeat_n, loss_type=loss_type, clip_loss=int(clip_loss),
                        with_time_emb=int(with_time_emb))


def tail_case_3dhp():
    """The flip-TTA tail of the 3DHP evaluate() (RUN3:522-529) with the MPI-INF-3DHP joint lists imported from the
    reference (common/mpiinf3dhp_dataset.py:17-18)."""
    sys.path.insert(0, "/root/reference")
    from common.mpiinf3dhp_dataset import joints_left, joints_right
    assert joints_left == synthetic.MPI3DHP_JOINTS_LEFT and joints_right == synthetic.MPI3DHP_JOINTS_RIGHT
    g = torch.Generator().manual_seed(8)
    y = torch.randn(2, 27, 17, 3, generator=g)
    yf = torch.randn(2, 27, 17, 3, generator=g)
    pf = yf.clone()
    pf[:, :, :, 0] *= -1
    pf[:, :, joints_left + joints_right] = pf[:, :, joints_right + joints_left]
    merged = (y + pf) / 2.0 * 0.9
    assert torch.equal(oracle.tta_merge(y, yf, 0.9, joints_left, joints_right), merged)
    np.savez_compressed(os.path.join(OUT, "tta_tail_3dhp.npz"), y=y.numpy(), yf=yf.numpy(), merged=merged.numpy(),
                        scale=0.9, left=np.array(joints_left), right=np.array(joints_right))


def weights_checksum():
    rows = {}
    for F in (27, 81, 243):
        m = synthetic.make_model(F)
        ref = RefModel(num_frame=F, num_joints=17, in_chans=2, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.,
                       qkv_bias=True, drop_path_rate=0.1, with_time_emb=True)
        torch.manual_seed(0)
        ref2 = RefModel(num_frame=F, num_joints=17, in_chans=2, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.,
                        qkv_bias=True, drop_path_rate=0.1, with_time_emb=True)
        for (k, v), (k2, v2) in zip(m.state_dict().items(), ref2.state_dict().items()):
            assert k == k2
            if "pos_embed" not in k:
                assert torch.equal(v, v2), f"seeded init differs from the reference for {k}"
        sd = m.state_dict()
        rows[f"F{F}_sum"] = np.array([v.double().sum().item() for v in sd.values()])
        rows[f"F{F}_abs"] = np.array([v.double().abs().sum().item() for v in sd.values()])
        rows[f"F{F}_n"] = np.array(sum(v.numel() for v in sd.values()))
        print(f"F={F}: params = {int(rows[f'F{F}_n'])}")
    np.savez_compressed(os.path.join(OUT, "weights_checksum.npz"), **rows)


def tail_case():
    g = torch.Generator().manual_seed(7)
    y = torch.randn(3, 9, 17, 3, generator=g)
    yf = torch.randn(3, 9, 17, 3, generator=g)
    gt = torch.randn(3, 9, 17, 3, generator=g)
    L, R = synthetic.H36M_JOINTS_LEFT, synthetic.H36M_JOINTS_RIGHT
    # literal RUN:583-588 + common/loss.py:15-27
    pf = yf.clone()
    pf[:, :, :, 0] *= -1
    pf[:, :, L + R] = pf[:, :, R + L]
    merged = (y + pf) / 2.0 * 1.7
    sys.path.insert(0, "/root/reference")
    from common.loss import mpjpe
    e = mpjpe(merged, gt)
    assert torch.equal(oracle.tta_merge(y, yf, 1.7), merged)
    assert torch.equal(oracle.mpjpe(merged, gt), e)
    np.savez_compressed(os.path.join(OUT, "tta_tail.npz"), y=y.numpy(), yf=yf.numpy(), gt=gt.numpy(),
                        merged=merged.numpy(), mpjpe=e.numpy(), scale=1.7)


CASES = [
    ("weights_checksum", weights_checksum, ()),
    ("tta_tail", tail_case, ()),
    ("tta_tail_3dhp", tail_case_3dhp, ()),
    ("denoise_f27_b3", denoise_case, ("denoise_f27_b3", 27, 3, [999, 500, 3])),
    ("denoise_f27_b2_notime", denoise_case, ("denoise_f27_b2_notime", 27, 2, [10, 20], False)),
    ("sampler_f27_b2_s3_clip", sampler_case, ("sampler_f27_b2_s3_clip", 27, 2, 3, 0.0, True)),
    ("sampler_f27_b2_s3_eta", sampler_case, ("sampler_f27_b2_s3_eta", 27, 2, 3, 0.5, True, True, True)),
    ("sampler_f27_b2_s2_notime", sampler_case, ("sampler_f27_b2_s2_notime", 27, 2, 2, 0.0, True, False)),
    ("sampler_f81_b1_s2_noclip", sampler_case, ("sampler_f81_b1_s2_noclip", 81, 1, 2, 0.0, False)),
    ("sampler_f243_b1_s1_clip", sampler_case, ("sampler_f243_b1_s1_clip", 243, 1, 1, 0.0, True)),
    ("sampler_f9_b2_s9_clip", sampler_case, ("sampler_f9_b2_s9_clip", 9, 2, 9, 0.0, True)),
    # the BASELINE configurations' own frame counts at the full 9 DDIM steps (cfg2: F = 81, cfg3 / cfg5: F = 243)
    ("sampler_f81_b1_s9_clip", sampler_case, ("sampler_f81_b1_s9_clip", 81, 1, 9, 0.0, True)),
    ("sampler_f243_b1_s9_clip", sampler_case, ("sampler_f243_b1_s9_clip", 243, 1, 9, 0.0, True)),
    ("sampler_f27_b2_s9_notime", sampler_case, ("sampler_f27_b2_s9_notime", 27, 2, 9, 0.0, True, False)),
    # forward() with output_loss=True (N1) and repeat_n
    ("forward_f27_b3_s2_loss", forward_case, ("forward_f27_b3_s2_loss", 27, 3, 2, 1, [999, 431, 7])),
    ("forward_f9_b2_s2_rep2_l1", forward_case, ("forward_f9_b2_s2_rep2_l1", 9, 2, 2, 2, [650, 12], "l1", False)),
]

if __name__ == "__main__":
    want = sys.argv[1:]
    for key, fn, a in CASES:
        if not want or any(w in key for w in want):
            fn(*a)
    print("golden vectors written to", OUT)
# tests/golden/state_dict_keys.txt (the 254 keys + shapes of the reference GaussianDiffusion state dict at F=27) is
# written by:  python - <<< "see git history / DESIGN.md";  it is regenerated below when run as a script.
def state_dict_keys():
    m = RefModel(num_frame=27, num_joints=17, in_chans=2, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.,
                 qkv_bias=True, drop_path_rate=0.1, with_time_emb=True)
    d = RefDiffusion(m, timesteps=1000, sampling_timesteps=9, loss_type='l2', clip_denoised=True,
                     beta_schedule='cosine', ddim_sampling_eta=0., clipLoss=True)
    with open(os.path.join(OUT, "state_dict_keys.txt"), "w") as f:
        f.write("\n".join(f"{k} {tuple(v.shape)}" for k, v in d.state_dict().items()) + "\n")


if __name__ == "__main__":
    state_dict_keys()
