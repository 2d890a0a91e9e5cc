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


def _gemm(mode, M, N, K, bn, act=0):
    torch, _lib, synthetic, Engine = _imports()
    os.environ["D3D_GEMM_BN"] = str(bn)
    eng = Engine(27, max_clips=1)
    g = torch.Generator().manual_seed(1)
    a, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g)
    ref = a.double() @ w.double().T + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    out = eng.op_linear(a.cuda(), w.cuda(), b.cuda(), act=act, gemm_mode=mode).cpu().double()
    err = (out - ref).abs()
    bad = (err > 1e-2).nonzero()
    return {"max_err": err.max().item(), "mean_err": err.mean().item(), "n_bad": int(bad.shape[0]),
            "first_bad": bad[:4].tolist(), "out00": out[0, :4].tolist(), "ref00": ref[0, :4].tolist()}


@stage
def gemm_simt():
    return _gemm(2, 300, 512, 512, 128)


@stage
def gemm_tc_fp16_bn128():
    return _gemm(1, 300, 512, 512, 128)


@stage
def gemm_tc_split3_bn128():
    return _gemm(0, 300, 512, 512, 128)


@stage
def gemm_tc_split3_bn256():
    return _gemm(0, 5000, 1536, 512, 256)


@stage
def gemm_tc_fp16_bn256_k1024():
    return _gemm(1, 5000, 512, 1024, 256)


@stage
def gemm_tc_gelu_bn256():
    return _gemm(0, 5000, 1024, 512, 256, act=1)


@stage
def gemm_tc_many_tiles():
    return _gemm(0, 70000, 1536, 512, 256)


def _attn(F, spatial, mode):
    torch, _lib, synthetic, Engine = _imports()
    sys.path.insert(0, ROOT)
    from oracle import diff3d_oracle as oracle
    B, J, C = 2, 17, 512
    eng = Engine(F, max_clips=B)
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * F * J, 3 * C, generator=g) * 1.5
    x = qkv.view(B, F, J, 3 * C)
    seqs = x.reshape(B * F, J, 3 * C) if spatial else x.permute(0, 2, 1, 3).reshape(B * J, F, 3 * C)
    ref = oracle.attention_core(seqs, 8)
    ref = ref.reshape(B, F, J, C) if spatial else ref.reshape(B, J, F, C).permute(0, 2, 1, 3)
    out = eng.op_attention(qkv.cuda(), B, spatial, mode).cpu().view(B, F, J, C)
    return {"max_err": (out - ref).abs().max().item()}


@stage
def attn_spatial():
    return _attn(27, True, 0)


@stage
def attn_spatial_simt():
    return _attn(27, True, 1)


@stage
def attn_temporal_simt_f81():
    return _attn(81, False, 1)


@stage
def attn_temporal_mma_f27():
    return _attn(27, False, 0)


@stage
def attn_temporal_mma_f243():
    return _attn(243, False, 0)


def _sampler(gemm_mode, attn_mode, use_graph, name="sampler_f27_b2_s3_clip"):
    torch, _lib, synthetic, Engine = _imports()
    import numpy as np
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
def sampler_tc_simtattn():
    return _sampler(0, 1, False)


@stage
def sampler_tc_default_graph():
    return _sampler(0, 0, True)


@stage
def sampler_f243():
    return _sampler(0, 0, True, "sampler_f243_b1_s1_clip")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(STAGES[sys.argv[1]]()), flush=True)
        sys.exit(0)
    for name in STAGES:
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            tail = (p.stderr.strip().splitlines() or [""])[-1][:300]
            print(f"{name:32s} rc={p.returncode} {res[0][7:] if res else 'NO RESULT: ' + tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"{name:32s} TIMEOUT", flush=True)
