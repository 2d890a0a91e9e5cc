"""CPU: the oracle against the committed golden vectors (generated from the imported, unmodified reference by
tools/make_golden.py) and against the known answers of SURVEY.md section 4."""
import numpy as np
import pytest
import torch

from diff3dhpe_b200 import synthetic
from oracle import diff3d_oracle as oracle

# fp32 CPU GEMM blocking differs between hosts (AVX2 / AVX-512 code paths), so cross-machine agreement with the
# golden vectors is to rounding, not bit-exact (it IS bit-exact on the machine that generated them).
TOL = 2e-5


def _sd(F, with_time_emb=True):
    m = synthetic.make_model(F, with_time_emb=with_time_emb)
    return {k: v.detach() for k, v in m.state_dict().items()}


def test_schedule_known_answers():
    assert oracle.ddim_times(1000, 9) == [999, 887, 776, 665, 554, 443, 332, 221, 110, -1]
    assert oracle.ddim_times(1000, 1) == [999, -1]
    assert oracle.ddim_times(1000, 1000)[:3] == [999, 998, 997]
    b = oracle.schedule_buffers(1000)
    ac = b["alphas_cumprod"]
    assert ac.dtype == torch.float32 and ac.shape == (1000,)
    assert torch.all(ac[1:] < ac[:-1]) and 0.99 < ac[0] < 1 and ac[-1] < 1e-4
    co = oracle.ddim_coefficients(b, oracle.ddim_times(1000, 9), 0.0)
    assert co[-1] is None and all(c["sigma"] == 0 for c in co[:-1])
    # eta = 0: c = sqrt(1 - alpha_next)
    assert torch.equal(co[0]["c"], (1 - ac[887]).sqrt())


def test_weights_match_reference_seeded_init(golden):
    g = golden("weights_checksum")
    for F, n in ((27, 43646467), (81, 43674115), (243, 43757059)):
        sd = synthetic.make_model(F).state_dict()
        assert sum(v.numel() for v in sd.values()) == n == int(g[f"F{F}_n"])
        s = np.array([v.double().sum().item() for v in sd.values()])
        a = np.array([v.double().abs().sum().item() for v in sd.values()])
        np.testing.assert_allclose(s, g[f"F{F}_sum"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(a, g[f"F{F}_abs"], rtol=1e-12)


@pytest.mark.parametrize("name", ["denoise_f27_b3", "denoise_f27_b2_notime"])
def test_forward_denoise_golden(golden, name):
    g = golden(name)
    F, B = int(g["F"]), int(g["B"])
    wt = "notime" not in name
    sd = _sd(F, wt)
    x2d, _ = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, 1)
    with torch.no_grad():
        out = oracle.forward_denoise(sd, torch.cat([x2d, y_T], -1), torch.tensor(g["t"], dtype=torch.long))
    assert np.abs(out.numpy() - g["out"]).max() < TOL


@pytest.mark.parametrize("name", ["sampler_f27_b2_s3_clip", "sampler_f27_b2_s3_eta", "sampler_f27_b2_s2_notime",
                                  "sampler_f9_b2_s9_clip"])
def test_sampler_golden(golden, name):
    g = golden(name)
    F, B, S = int(g["F"]), int(g["B"]), int(g["S"])
    sd = _sd(F, bool(g["with_time_emb"]))
    x2d, _ = synthetic.make_inputs(B, F)
    y_T, steps = synthetic.make_noise(B, F, S)
    with torch.no_grad():
        out = oracle.ddim_sample_loop(sd, x2d, y_T, steps, sampling_timesteps=S, eta=float(g["eta"]),
                                      clip_denoised=bool(g["clip"]), trace="rev" in g)
    if "rev" in g:
        out, rev, x0s = out
        assert np.abs(rev.numpy() - g["rev"]).max() < 5e-5
        assert np.abs(x0s.numpy() - g["x0s"]).max() < 5e-5
    assert np.abs(out.numpy() - g["pred"]).max() < 5e-5


def test_tta_tail_golden(golden):
    g = golden("tta_tail")
    merged = oracle.tta_merge(torch.from_numpy(g["y"]), torch.from_numpy(g["yf"]), float(g["scale"]))
    assert np.array_equal(merged.numpy(), g["merged"])
    e = oracle.mpjpe(merged, torch.from_numpy(g["gt"]))
    assert abs(e.item() - float(g["mpjpe"])) < 1e-6


@pytest.mark.parametrize("name", ["forward_f27_b3_s2_loss", "forward_f9_b2_s2_rep2_l1"])
def test_forward_output_loss_and_repeat_n_golden(golden, name):
    """q_sample / p_losses / the eval branch of forward() with repeat_n (DIFF:360-366, 392-419, 427-449) against the
    imported reference's loss tensor and prediction (tools/make_golden.py forward_case pins randint and every normal
    draw)."""
    g = golden(name)
    F, B, S, rep = int(g["F"]), int(g["B"]), int(g["S"]), int(g["repeat_n"])
    sd = _sd(F)
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, steps = synthetic.make_noise(rep * B, F, S)
    loss_noise = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(777))
    with torch.no_grad():
        loss, pred = oracle.forward_eval(sd, gt, x2d, y_T, steps, repeat_n=rep,
                                         loss_draws=(torch.tensor(g["t"], dtype=torch.long), loss_noise),
                                         sampling_timesteps=S, loss_type=str(g["loss_type"]), clip_loss=bool(g["clip_loss"]))
    assert pred.shape == (B, F, 17, 3)
    assert np.abs(pred.numpy() - g["pred"]).max() < 5e-5
    assert np.abs(loss.numpy() - g["loss"]).max() < 5e-5 * max(1.0, float(np.abs(g["loss"]).max()))
    # the clamp of DIFF:413-414 is active in the first case (t = 7: 1 + abar / sqrt(1 - abar) > 3) and off in the second
    bufs = oracle.schedule_buffers(1000)
    coef = 1.0 + bufs["alphas_cumprod"][torch.tensor(g["t"])] / bufs["sqrt_one_minus_alphas_cumprod"][torch.tensor(g["t"])]
    assert (coef.max() > 3.0) and bool(g["clip_loss"]) == (name == "forward_f27_b3_s2_loss")


def test_tta_tail_3dhp_golden(golden):
    g = golden("tta_tail_3dhp")
    merged = oracle.tta_merge(torch.from_numpy(g["y"]), torch.from_numpy(g["yf"]), float(g["scale"]),
                              oracle.MPI3DHP_JOINTS_LEFT, oracle.MPI3DHP_JOINTS_RIGHT)
    assert np.array_equal(merged.numpy(), g["merged"])
    assert g["left"].tolist() == oracle.MPI3DHP_JOINTS_LEFT and g["right"].tolist() == oracle.MPI3DHP_JOINTS_RIGHT


def test_draw_order_matches_reference_count():
    y_T, steps = oracle.draw_noise((2, 9, 17, 3), 9, torch.Generator().manual_seed(3))
    assert y_T.shape == (2, 9, 17, 3) and steps.shape == (8, 2, 9, 17, 3)
    g = torch.Generator().manual_seed(3)
    assert torch.equal(y_T, torch.randn(2, 9, 17, 3, generator=g))
    assert torch.equal(steps[0], torch.randn(2, 9, 17, 3, generator=g))


def test_grand_identity_and_flip_batching():
    """(P - I) V == P V - V, and orig||flip as one 2B batch equals two separate calls."""
    g = torch.Generator().manual_seed(0)
    p = torch.randn(4, 8, 17, 17, generator=g).softmax(-1)
    v = torch.randn(4, 8, 17, 64, generator=g)
    eye = torch.eye(17).view(1, 1, 17, 17)
    assert ((p - eye) @ v - (p @ v - v)).abs().max() < 2e-6
    sd = _sd(9)
    x2d, _ = synthetic.make_inputs(2, 9)
    xf = oracle.flip_2d(x2d)
    y_T, _ = synthetic.make_noise(4, 9, 1)
    kw = dict(sampling_timesteps=1)
    with torch.no_grad():
        both = oracle.ddim_sample_loop(sd, torch.cat([x2d, xf]), y_T, None, **kw)
        a = oracle.ddim_sample_loop(sd, x2d, y_T[:2], None, **kw)
        b = oracle.ddim_sample_loop(sd, xf, y_T[2:], None, **kw)
    assert (both - torch.cat([a, b])).abs().max() < 1e-5


def test_flip_is_involution_and_uses_h36m_lists():
    x, _ = synthetic.make_inputs(2, 3)
    assert torch.equal(oracle.flip_2d(oracle.flip_2d(x)), x)
    assert oracle.H36M_JOINTS_LEFT == [4, 5, 6, 11, 12, 13] and oracle.H36M_JOINTS_RIGHT == [1, 2, 3, 14, 15, 16]
    assert torch.equal(oracle.flip_2d(x), synthetic.flip_2d(x))


def test_windowing_golden(golden):
    """Window bounds, 2D slices, flipped copies and target masks against the reference ChunkedGenerator's own output
    (tools/make_golden_windows.py): three sequences of 9 / 20 / 31 frames, F = 9."""
    g = golden("windows_f9")
    F, lens = int(g["F"]), g["lens"].tolist()
    seq = torch.from_numpy(g["seq2d"])
    w, base = 0, 0
    for n in lens:
        starts, targets = oracle.chunk_windows(n, F)
        for st, tg in zip(starts, targets):
            assert base + st == int(g["win_start"][w])
            x, m = oracle.window_batch(seq[base:base + n], st, F, tg, False)
            xf, _ = oracle.window_batch(seq[base:base + n], st, F, tg, True)
            assert np.array_equal(x.numpy(), g["x2d"][w]) and np.array_equal(xf.numpy(), g["x2d_flip"][w])
            assert np.array_equal(m.numpy(), g["mask"][w])
            w += 1
        base += n
    assert w == g["x2d"].shape[0]


def test_metrics_golden(golden):
    """mpjpe / n_mpjpe / p_mpjpe / velocity restatements against the reference's own common/loss.py outputs
    (tools/make_golden_metrics.py; the first 7 frames are mirrored to exercise the reflection branch of p_mpjpe)."""
    g = golden("metrics")
    tp, tg = torch.from_numpy(g["pred"]).unsqueeze(1), torch.from_numpy(g["gt"]).unsqueeze(1)
    assert oracle.mpjpe(tp, tg).item() == float(g["mpjpe"])
    assert oracle.n_mpjpe(tp, tg).item() == float(g["n_mpjpe"])
    assert oracle.p_mpjpe(g["pred"], g["gt"]) == float(g["p_mpjpe"])
    assert oracle.mean_velocity_error(g["pred"], g["gt"]) == float(g["velocity"])
