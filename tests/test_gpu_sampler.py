"""GPU (-m gpu): the hot path (forward_denoise, DDIM loop, flip-TTA) through the drop-in modules / C ABI against
the golden vectors of the imported reference and the CPU oracle run live on the same seeded inputs.

Tolerances (north star): per-joint max-abs error <= 1e-2 (pose scale 1) and MPJPE delta <= 0.1 mm = 1e-4.  The
default mode (3-pass split-fp16 linears, single-pass fp16 attention with an exact "- V" term) is asserted against
a 4x tighter bound; the CPU precision emulation (tools/precision_probe.py) predicts max-abs 6e-4 / 8e-4 / 1.5e-3
at F = 27 / 81 / 243 with 9 steps."""
import numpy as np
import pytest
import torch

from diff3dhpe_b200 import _lib, synthetic
from oracle import diff3d_oracle as oracle

pytestmark = pytest.mark.gpu

MAXABS_BAR, MPJPE_BAR = 1e-2, 1e-4
MARGIN = 4.0


def _diffusion(F, S, eta=0.0, clip=True, with_time_emb=True, gemm_mode=_lib.GEMM_TC_F8C, attn_mode=_lib.ATTN_DEFAULT,
               use_graph=True, max_clips=1):
    m = synthetic.make_model(F, with_time_emb=with_time_emb).cuda()
    m.gemm_mode, m.attn_mode, m.use_graph, m.max_clips_hint = gemm_mode, attn_mode, use_graph, max_clips
    return synthetic.make_diffusion(m, sampling_timesteps=S, eta=eta, clip_denoised=clip).cuda().eval()


def _mpjpe_delta(a, b, gt):
    return abs(oracle.mpjpe(a, gt).item() - oracle.mpjpe(b, gt).item())


@pytest.mark.parametrize("name", ["denoise_f27_b3", "denoise_f27_b2_notime"])
@pytest.mark.parametrize("gemm_mode", [_lib.GEMM_SIMT_FP32, _lib.GEMM_TC_SPLIT3, _lib.GEMM_TC_F8C])
def test_forward_denoise_golden(golden, name, gemm_mode):
    g = golden(name)
    F, B = int(g["F"]), int(g["B"])
    diff = _diffusion(F, 1, with_time_emb="notime" not in name, gemm_mode=gemm_mode)
    x2d, _ = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, 1)
    out = diff.model.forward_denoise(torch.cat([x2d, y_T], -1).cuda(), torch.tensor(g["t"]).cuda()).cpu().numpy()
    assert np.abs(out - g["out"]).max() < MAXABS_BAR / MARGIN


@pytest.mark.parametrize("name", ["sampler_f27_b2_s3_clip", "sampler_f27_b2_s3_eta", "sampler_f27_b2_s2_notime",
                                  "sampler_f81_b1_s2_noclip", "sampler_f243_b1_s1_clip", "sampler_f9_b2_s9_clip"])
@pytest.mark.parametrize("use_graph,gemm_mode", [(False, _lib.GEMM_TC_F8C), (True, _lib.GEMM_TC_F8C),
                                                 (True, _lib.GEMM_TC_SPLIT3)])
def test_sampler_golden(golden, name, use_graph, gemm_mode):
    g = golden(name)
    F, B, S, eta = int(g["F"]), int(g["B"]), int(g["S"]), float(g["eta"])
    diff = _diffusion(F, S, eta, bool(g["clip"]), bool(g["with_time_emb"]), gemm_mode=gemm_mode, use_graph=use_graph,
                      max_clips=B)
    x2d, gt = synthetic.make_inputs(B, F)
    y_T, steps = synthetic.make_noise(B, F, S)
    noise = (y_T.cuda(), steps.cuda() if eta != 0 else None)
    trace = "rev" in g
    if trace:
        pred, rev, x0s = diff.ddim_sample_loop_ouput_reverse_diffusion(x2d.cuda(), [B, F, 17, 3], noise=noise)
        assert np.abs(rev.cpu().numpy() - g["rev"]).max() < MAXABS_BAR / MARGIN
        assert np.abs(x0s.cpu().numpy() - g["x0s"]).max() < MAXABS_BAR / MARGIN
    else:
        pred = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=noise)
        pred2 = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=noise)      # graph replay / determinism
        assert torch.equal(pred, pred2)
    pred = pred.cpu()
    ref = torch.from_numpy(g["pred"])
    assert (pred - ref).abs().max().item() < MAXABS_BAR / MARGIN
    assert _mpjpe_delta(pred, ref, gt) < MPJPE_BAR / MARGIN


def test_fp16_fast_mode_is_within_maxabs_bar(golden):
    g = golden("sampler_f27_b2_s3_clip")
    diff = _diffusion(27, 3, gemm_mode=_lib.GEMM_TC_FP16, max_clips=2)
    x2d, _ = synthetic.make_inputs(2, 27)
    y_T, _ = synthetic.make_noise(2, 27, 3)
    pred = diff.ddim_sample_loop(x2d.cuda(), [2, 27, 17, 3], noise=(y_T.cuda(), None)).cpu().numpy()
    assert np.abs(pred - g["pred"]).max() < 3e-2      # documented fast mode: looser than the parity bar


def test_forward_api_flip_tta_against_oracle():
    """evaluate()'s two GaussianDiffusion.forward calls + un-flip/average (RUN:577-588) vs the oracle, S=9."""
    F, B, S = 27, 2, 9
    diff = _diffusion(F, S, max_clips=B)
    x2d, gt = synthetic.make_inputs(B, F)
    xf = synthetic.flip_2d(x2d)
    n1, n2 = synthetic.make_noise(B, F, S, seed=1), synthetic.make_noise(B, F, S, seed=2)
    y = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=(n1[0].cuda(), None))
    yf = diff.ddim_sample_loop(xf.cuda(), [B, F, 17, 3], noise=(n2[0].cuda(), None))
    merged = diff.model.engine(B).tta_merge(y, yf, synthetic.H36M_JOINTS_LEFT, synthetic.H36M_JOINTS_RIGHT, 1.0).cpu()
    sd = {k: v.detach().cpu() for k, v in diff.model.state_dict().items()}
    with torch.no_grad():
        ref = oracle.sample_tta(sd, x2d, n1, n2, sampling_timesteps=S)
    assert (merged - ref).abs().max().item() < MAXABS_BAR / MARGIN
    assert _mpjpe_delta(merged, ref, gt) < MPJPE_BAR / MARGIN
    # forward(): same signature / return convention as DIFF:421-449, and it consumes S normal draws in order
    torch.manual_seed(77)
    loss, pred = diff(clean_3d_pose=gt.cuda(), noisy_2d_pose=x2d.cuda(), output_loss=False)
    torch.manual_seed(77)
    y_T = torch.randn(B, F, 17, 3, device="cuda")
    for _ in range(S - 1):
        torch.randn_like(y_T)
    after = torch.randn(3, device="cuda")
    pred2 = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=(y_T, None))
    assert loss is None and torch.equal(pred, pred2)
    torch.manual_seed(77)
    diff(clean_3d_pose=gt.cuda(), noisy_2d_pose=x2d.cuda(), output_loss=False)
    assert torch.equal(after, torch.randn(3, device="cuda"))
    loss, _ = diff(clean_3d_pose=gt.cuda(), noisy_2d_pose=x2d.cuda(), output_loss=True)     # N1 row: p_losses
    assert loss.shape == (B, F, 17, 3) and torch.isfinite(loss).all()


def test_batch_split_invariance_and_host_api():
    """Clips are independent: sampling 6 clips at once == sampling 4 + 2 (bit-exact on the GPU), which is what
    makes rank-sharding exact; the host-buffer entry point returns the same bytes as the device one."""
    F, B, S = 81, 6, 2
    diff = _diffusion(F, S, max_clips=B)
    x2d, _ = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, S)
    xd, nd = x2d.cuda(), y_T.cuda()
    full = diff.ddim_sample_loop(xd, [B, F, 17, 3], noise=(nd, None))
    a = diff.ddim_sample_loop(xd[:4].contiguous(), [4, F, 17, 3], noise=(nd[:4].contiguous(), None))
    b = diff.ddim_sample_loop(xd[4:].contiguous(), [2, F, 17, 3], noise=(nd[4:].contiguous(), None))
    assert torch.equal(full, torch.cat([a, b]))
    eng = diff._engine(B)
    xh, nh, yh = x2d.pin_memory(), y_T.pin_memory(), torch.empty(B, F, 17, 3).pin_memory()
    eng.ddim_sample_host(xh, nh, None, yh)
    assert torch.equal(yh, full.cpu())
    assert eng.launch_count() > 0


def test_flip_equivariance_property_full_size():
    """Size-independent property at a BASELINE-size batch (cfg2 shape: F=81, many clips): the merged TTA output of
    the flipped input is the flip of the merged output (the model need not be equivariant, the merge is)."""
    F, B, S = 81, 32, 1
    diff = _diffusion(F, S, max_clips=2 * B)
    x2d, _ = synthetic.make_inputs(B, F)
    xf = synthetic.flip_2d(x2d)
    y_T, _ = synthetic.make_noise(2 * B, F, S)
    eng = diff._engine(2 * B)
    L, R = synthetic.H36M_JOINTS_LEFT, synthetic.H36M_JOINTS_RIGHT
    both = diff.ddim_sample_loop(torch.cat([x2d, xf]).cuda(), [2 * B, F, 17, 3], noise=(y_T.cuda(), None))
    m1 = eng.tta_merge(both[:B].contiguous(), both[B:].contiguous(), L, R, 1.0)
    m2 = eng.tta_merge(both[B:].contiguous(), both[:B].contiguous(), L, R, 1.0)
    assert torch.allclose(m2, synthetic.flip_2d(m1.cpu()).cuda(), atol=1e-6)
    assert torch.isfinite(both).all() and both.abs().max() <= 1.0


def test_evaluate_sequences_matches_host_windowing():
    """N3: the whole evaluate() inner loop from raw packed sequences (device windowing + flip + sampler + merge + masked
    write-back) against the same windows built on the host the way the reference's generator does; bit-identical."""
    from diff3dhpe_b200 import evaluate
    F, S, lens = 9, 2, [9, 20, 31]
    N = sum(lens)
    dev = torch.device("cuda", 0)
    model = synthetic.make_model(F).cuda()
    model.max_clips_hint = 16
    diff = synthetic.make_diffusion(model, sampling_timesteps=S).cuda().eval()
    sampler = evaluate.DeviceSampler(diff)
    g = torch.Generator().manual_seed(5)
    seq2d = (0.3 * torch.randn(N, 17, 2, generator=g)).clamp_(-1, 1)
    gt = 0.3 * torch.randn(N, 17, 3, generator=g)

    def noise_fn(ids, flip):
        ys = [synthetic.make_noise(1, F, S, seed=100 + 2 * int(i) + int(flip))[0] for i in ids]
        return torch.cat(ys).to(dev), None

    res = evaluate.evaluate_sequences(sampler, seq2d, gt, lens, noise_fn, device=dev, F=F, batch_clips=3)
    # host-side windows (oracle restatement of the generator) through evaluate_shard
    ws, fv, _ = evaluate.plan_windows(lens, F)
    x_h = torch.stack([seq2d[int(s):int(s) + F] for s in ws])
    gt_h = torch.stack([gt[int(s):int(s) + F] for s in ws])
    mask = torch.ones(ws.numel(), F, dtype=torch.uint8)
    for w, v in enumerate(fv.tolist()):
        mask[w, :v] = 0
    ref = evaluate.evaluate_shard(sampler, x_h, gt_h, noise_fn, device=dev, batch_clips=3, tta=True, frame_mask=mask)
    packed = torch.zeros(N, 17, 3)
    for w, (s, v) in enumerate(zip(ws.tolist(), fv.tolist())):
        packed[s + v:s + F] = ref["pred"][w, v:].cpu()
    assert res["n_windows"] == 8
    assert torch.equal(res["pred"].cpu(), packed)
    a, b = res["acc"].cpu(), ref["acc"].cpu()
    assert a[1].item() == b[1].item() == N * 17 and abs(a[0].item() - b[0].item()) < 1e-9 * b[0].item()
