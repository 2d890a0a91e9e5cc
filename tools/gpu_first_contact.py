"""Staged bring-up on a real B200: every stage runs in its own subprocess with a timeout, so a trap / sticky
CUDA error / hang in one kernel does not hide the others.  Prints one line per stage."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def _imports():
    sys.path.insert(0, ROOT)
    import torch
    from diff3dhpe_b200 import _lib, synthetic
    from diff3dhpe_b200.engine import Engine
    return torch, _lib, synthetic, Engine


def _gemm(mode, M, N, K, bn, act=0, cg=2, residual=False, cs=2):
    torch, _lib, synthetic, Engine = _imports()
    os.environ["D3D_GEMM_BN"] = str(bn)
    os.environ["D3D_GEMM_CG"] = str(cg)
    os.environ["D3D_GEMM_CS"] = str(cs)
    eng = Engine(27, max_clips=1)
    g = torch.Generator().manual_seed(1)
    a, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g) if residual else None
    ref = a.double() @ w.double().T + b.double()
    if residual:
        ref = ref + res.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    out = eng.op_linear(a.cuda(), w.cuda(), b.cuda(), residual=res.cuda() if residual else None, act=act,
                        gemm_mode=mode).cpu().double()
    err = (out - ref).abs()
    bad = (err > 1e-2).nonzero()
    return {"max_err": err.max().item(), "mean_err": err.mean().item(), "n_bad": int(bad.shape[0]),
            "first_bad": bad[:4].tolist(), "out00": out[0, :4].tolist(), "ref00": ref[0, :4].tolist()}


@stage
def gemm_simt():
    return _gemm(2, 300, 512, 512, 128)


@stage
def gemm_cg1_split3_bn256():
    return _gemm(0, 5000, 1536, 512, 256, cg=1)


@stage
def gemm_cg1_residual_bn128():
    return _gemm(0, 300, 512, 512, 128, cg=1, residual=True)


@stage
def gemm_cg2_split3_one_tile():
    return _gemm(0, 256, 256, 64, 256, cg=2)


@stage
def gemm_cg2_split3_small():
    return _gemm(0, 300, 512, 512, 256, cg=2)


@stage
def gemm_cg2_split3_residual():
    return _gemm(0, 5000, 512, 1024, 256, cg=2, residual=True)


@stage
def gemm_cg2_gelu():
    return _gemm(0, 5000, 1024, 512, 256, act=1, cg=2)


@stage
def gemm_cg2_fp16():
    return _gemm(1, 5000, 1536, 512, 256, cg=2)


@stage
def gemm_simt_f8c():
    return _gemm(4, 300, 512, 512, 256)


@stage
def gemm_cg1_f8c():
    return _gemm(3, 300, 512, 512, 256, cg=1)


@stage
def gemm_cg2_f8c_one_tile():
    return _gemm(3, 256, 256, 64, 256, cg=2)


@stage
def gemm_cg2_f8c_residual_cs1():
    return _gemm(3, 5000, 512, 1024, 256, cg=2, residual=True, cs=1)


@stage
def gemm_cs2_f8c_two_tiles():
    return _gemm(3, 512, 256, 64, 256, cg=2, cs=2)


@stage
def gemm_cs2_f8c_odd_tiles():
    return _gemm(3, 700, 512, 512, 256, cg=2, cs=2)


@stage
def gemm_cs2_f8c_residual():
    return _gemm(3, 5000, 512, 1024, 256, cg=2, residual=True, cs=2)


@stage
def gemm_cs2_f8c_gelu():
    return _gemm(3, 5000, 1024, 512, 256, act=1, cg=2, cs=2)


@stage
def gemm_cs2_f8c_many_tiles():
    return _gemm(3, 70001, 1536, 512, 256, cg=2, residual=True, cs=2)


@stage
def gemm_cg2_many_tiles():
    return _gemm(0, 70001, 1536, 512, 256, cg=2, residual=True)


def _attn(F, spatial, mode):
    torch, _lib, synthetic, Engine = _imports()
    sys.path.insert(0, ROOT)
    from oracle import diff3d_oracle as oracle
    B, J, C = 2, 17, 512
    eng = Engine(F, max_clips=B)
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * F * J, 3 * C, generator=g) * 1.5
    qkv[:, :2 * C] = qkv[:, :2 * C].half().float()      # the packed layout keeps q, k as fp16
    x = qkv.view(B, F, J, 3 * C)
    seqs = x.reshape(B * F, J, 3 * C) if spatial else x.permute(0, 2, 1, 3).reshape(B * J, F, 3 * C)
    ref = oracle.attention_core(seqs, 8)
    ref = ref.reshape(B, F, J, C) if spatial else ref.reshape(B, J, F, C).permute(0, 2, 1, 3)
    out = eng.op_attention(qkv.cuda(), B, spatial, mode).cpu().view(B, F, J, C)
    err = (out - ref).abs()
    return {"max_err": err.max().item(), "mean_err": err.mean().item(), "argmax": list(map(int, (err == err.max()).nonzero()[0]))}


@stage
def attn_spatial_simt():
    return _attn(27, True, 1)


@stage
def attn_spatial_mma():
    return _attn(27, True, 0)


@stage
def attn_temporal_simt_f81():
    return _attn(81, False, 1)


@stage
def attn_temporal_mma_f27():
    return _attn(27, False, 0)


@stage
def attn_temporal_mma_f243():
    return _attn(243, False, 0)


def _sampler(gemm_mode, attn_mode, use_graph, name="sampler_f27_b2_s3_clip", cg=2):
    torch, _lib, synthetic, Engine = _imports()
    import numpy as np
    os.environ["D3D_GEMM_CG"] = str(cg)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    F, B, S = int(g["F"]), int(g["B"]), int(g["S"])
    m = synthetic.make_model(F).cuda()
    m.gemm_mode, m.attn_mode, m.use_graph, m.max_clips_hint = gemm_mode, attn_mode, use_graph, B
    diff = synthetic.make_diffusion(m, sampling_timesteps=S, clip_denoised=bool(g["clip"])).cuda()
    x2d, _ = synthetic.make_inputs(B, F)
    y_T, _ = synthetic.make_noise(B, F, S)
    pred = diff.ddim_sample_loop(x2d.cuda(), [B, F, 17, 3], noise=(y_T.cuda(), None)).cpu().numpy()
    return {"max_err": float(np.abs(pred - g["pred"]).max())}


@stage
def sampler_simt_simt():
    return _sampler(2, 1, False)


@stage
def sampler_cg1_simtattn():
    return _sampler(0, 1, False, cg=1)


@stage
def sampler_cg2_simtattn():
    return _sampler(0, 1, False, cg=2)


@stage
def sampler_default_graph():
    return _sampler(0, 0, True)


@stage
def sampler_f243():
    return _sampler(0, 0, True, "sampler_f243_b1_s1_clip")


@stage
def sampler_f9_s9():
    return _sampler(0, 0, True, "sampler_f9_b2_s9_clip")


@stage
def sampler_f8c_simt():
    return _sampler(4, 1, False)


@stage
def sampler_f8c():
    return _sampler(3, 0, True)


@stage
def sampler_f8c_f243():
    return _sampler(3, 0, True, "sampler_f243_b1_s1_clip")


@stage
def sampler_f8c_f9_s9():
    return _sampler(3, 0, True, "sampler_f9_b2_s9_clip")


def _attn_tc(F, B, gemm_mode=3, spatial=False):
    """tcgen05 temporal kernel vs the CUDA-core kernel: error statistics by query row / head / channel to localise
    layout bugs (descriptor, swizzle, TMEM packing)."""
    torch, _lib, synthetic, Engine = _imports()
    J, C = 17, 512
    eng = Engine(F, max_clips=B, gemm_mode=gemm_mode)
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(B * F * J, 3 * C, generator=g) * 1.5
    qkv[:, :2 * C] = qkv[:, :2 * C].half().float()
    ref = eng.op_attention(qkv.cuda(), B, spatial, 1).cpu().view(B, F, J, 8, 64)
    out = eng.op_attention(qkv.cuda(), B, spatial, 0).cpu().view(B, F, J, 8, 64)
    hi, second = eng.debug_attention_operand(qkv.cuda(), B, spatial, 0)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    res = {"max_err": err.max().item(), "mean_err": err.mean().item(), "finite": bool(torch.isfinite(out).all()),
           "err_by_clip": [round(v, 5) for v in err.amax(dim=(1, 2, 3, 4)).tolist()],
           "err_by_head": [round(v, 5) for v in err.amax(dim=(0, 1, 2, 4)).tolist()],
           "err_by_frame_first8": [round(v, 5) for v in err.amax(dim=(0, 2, 3, 4)).tolist()[:8]],
           "err_by_frame_128_136": [round(v, 5) for v in err.amax(dim=(0, 2, 3, 4)).tolist()[128:136]],
           "err_by_frame_last4": [round(v, 5) for v in err.amax(dim=(0, 2, 3, 4)).tolist()[-4:]],
           "err_by_chan_first8": [round(v, 5) for v in err.amax(dim=(0, 1, 2, 3)).tolist()[:8]],
           "err_by_joint": [round(v, 5) for v in err.amax(dim=(0, 1, 3, 4)).tolist()],
           "out0": out[0, 0, 0, 0, :4].tolist(), "ref0": ref[0, 0, 0, 0, :4].tolist()}
    hi = hi.float().cpu().view(B, F, J, 8, 64)
    res["hi_err"] = (hi - ref).abs().max().item()
    if gemm_mode == 3:
        a8 = second[:, :C].view(torch.float8_e5m2).float().cpu().view(B, F, J, 8, 64) * 256.0
        res["a8_rel_err"] = ((a8 - ref).abs() / (ref.abs() + 1e-2)).max().item()
    return res


@stage
def attn_tc_f243():
    return _attn_tc(243, 5)


@stage
def attn_sp_tc_f243():
    return _attn_tc(243, 5, spatial=True)


@stage
def attn_sp_tc_f27_split16():
    return _attn_tc(27, 3, gemm_mode=0, spatial=True)


@stage
def attn_tc_f81():
    return _attn_tc(81, 3)


@stage
def attn_tc_f243_split16():
    return _attn_tc(243, 2, gemm_mode=0)


@stage
def bench_attention_tc():
    """ms / call of the temporal attention alone at 256 clips x 243 frames: tcgen05 kernel vs mma.sync kernel, timed
    on the operand path (pack kernel excluded by timing a pair of calls with and without attention is not possible
    through the ABI, so the pack + copies are included in both and reported separately)."""
    torch, _lib, synthetic, Engine = _imports()
    B, F = 256, 243
    eng = Engine(F, max_clips=B)
    qkv = torch.randn(B * F * 17, 1536, device="cuda")
    out = {}
    for name, mode in (("tc", 0), ("mma_sync", 2), ("spatial", -1)):
        args = (qkv, B, mode < 0, 0 if mode < 0 else mode)
        for _ in range(2):
            eng.debug_attention_operand(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.debug_attention_operand(*args)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_plus_pack_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    return out


@stage
def bench_gemm():
    """ms / launch of the GEMM kernel alone at the cfg3 half-batch size (M = 1 057 536 tokens)."""
    torch, _lib, synthetic, Engine = _imports()
    eng = Engine(27, max_clips=1)
    out = {}
    M = 1057536
    os.environ["D3D_GEMM_BN"] = "256"
    for cg, cs in ((1, 1), (2, 1), (2, 2)):
        os.environ["D3D_GEMM_CG"] = str(cg)
        os.environ["D3D_GEMM_CS"] = str(cs)
        for mode, name, passes in ((_lib.GEMM_TC_SPLIT3, "split3", 3), (_lib.GEMM_TC_F8C, "f8c", 2), (_lib.GEMM_TC_FP16, "fp16", 1)):
            if cs == 2 and name != "f8c":
                continue
            for N, K, act in ((1536, 512, 0), (512, 512, 0), (1024, 512, 1), (512, 1024, 0)):
                ms = eng.op_linear_bench(M, N, K, act, mode, iters=5)
                tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
                out[f"cg{cg}cs{cs}_{name}_N{N}_K{K}"] = [round(ms, 3), round(tf, 1), round(tf * passes, 1)]
    return out


@stage
def bench_attention():
    """ms / launch of the attention kernels at 128 clips x 243 frames (529 k tokens)."""
    torch, _lib, synthetic, Engine = _imports()
    B, F = 128, 243
    eng = Engine(F, max_clips=B)
    qkv = torch.randn(B * F * 17, 1536, device="cuda")
    out = {}
    for spatial in (True, False):
        for _ in range(2):
            eng.op_attention(qkv, B, spatial, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.op_attention(qkv, B, spatial, 0)
        e1.record()
        torch.cuda.synchronize()
        out["spatial" if spatial else "temporal"] = round(e0.elapsed_time(e1) / 5, 3)    # includes the fp32->fp16 pack kernel
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(STAGES[sys.argv[1]]()), flush=True)
        sys.exit(0)
    only = os.environ.get("STAGES")
    for name in STAGES:
        if only and name not in only.split(","):
            continue
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=400)
            res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            tail = (p.stderr.strip().splitlines() or [""])[-1][:300]
            print(f"{name:32s} rc={p.returncode} {res[0][7:] if res else 'NO RESULT: ' + tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"{name:32s} TIMEOUT", flush=True)
