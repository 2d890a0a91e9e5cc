"""Predict the parity of a GEMM / attention operand-precision policy WITHOUT a GPU (SURVEY.md 7.4 recipe).

A TorchFunctionMode intercepts F.linear (K >= 64) and torch.matmul while the CPU oracle's DDIM sampler runs,
replaces the operands by their rounded / split versions and accumulates in fp32, then compares the result with
the plain fp32 run on the same weights, inputs and noise.

    python tools/precision_probe.py F B [S]

Linear modes:
  fp16      a_hi.b_hi                                             (1 fp16 pass)
  split3    a_hi.b_hi + a_hi.b_lo + a_lo.b_hi                     (3 fp16 passes)
  f8corr    a_hi.b_hi + 2^-16 (e4m3(a).e4m3(2^16 b_lo) + e4m3(2^12 a_lo).e4m3(2^4 b))   (1 fp16 + 2 fp8 passes, 2 accumulators)
  f8c52     a_hi.b_hi + e5m2(2^-8 a).e5m2(2^8 b_lo) + e5m2(2^4 a_lo).e5m2(2^-4 b)       (same, ONE accumulator)
  f4c       a_hi.b_hi + q4(a).q4(b_lo) + q4(a_lo).q4(b), q4 = e2m1 with one power-of-two scale per 32 elements of K
            (kind::mxf4, 4x the fp16 rate: 1.5 tensor-pipe units);  f4c_16: power-of-two scale per 16 elements
            (kind::mxf4nvf4 with ue8m0 scales);  f4c_nv: e4m3 scale per 16 elements (nvfp4) -- NOTE its emulation
            applies a free per-tensor power of two, which one shared accumulator cannot give: the a . w_lo product
            would need scale values 21 binades apart and ue4m3 spans 17
Attention modes: fp32 | fp16 | split3 | qk3pv1 (QK^T split3, PV single fp16 pass)
"""
import os
import sys

import torch
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff3dhpe_b200 import synthetic  # noqa: E402
from oracle import diff3d_oracle as oracle  # noqa: E402


def h16(x):
    return x.to(torch.float16).float()


def split16(x):
    hi = h16(x)
    return hi, h16(x - hi)


def e4m3(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def e5m2(x):
    return x.clamp(-57344.0, 57344.0).to(torch.float8_e5m2).float()


def e2m1(x):
    """Round-to-nearest-even onto the e2m1 grid {0, .5, 1, 1.5, 2, 3, 4, 6} (saturating): steps of 0.5 below 2, 1 below
    4, 2 above (torch.round is ties-to-even, which is ties-to-even-mantissa on this grid)."""
    ax = x.abs().clamp(max=6.0)
    q = torch.where(ax < 2.0, torch.round(ax * 2.0) * 0.5, torch.where(ax < 4.0, torch.round(ax), torch.round(ax * 0.5) * 2.0))
    return torch.copysign(q, x)


def blockq4(x, dim, block, scale):
    """Block-scaled e2m1 along `dim` (the contraction index), as tcgen05.mma kind::mxf4 / mxf4nvf4 consumes it:
    scale = "ue8m0": one power-of-two scale per `block` elements (mxfp4, block 32), chosen so that the block maximum
    lands in (3, 6]; "ue4m3": one e4m3 scale per block (nvfp4, block 16), amax / 6 rounded to e4m3."""
    xm = x.movedim(dim, -1)
    K = xm.shape[-1]
    pad = (-K) % block
    if pad:
        xm = F.pad(xm, (0, pad))
    xb = xm.reshape(*xm.shape[:-1], -1, block)
    amax = xb.abs().amax(dim=-1, keepdim=True).clamp(min=1e-30)
    if scale == "ue8m0":
        sf = torch.exp2(torch.ceil(torch.log2(amax / 6.0)))
    elif scale == "ue8m0_best":
        # the better (block squared error) of the non-saturating power of two and the one below it (block maximum in
        # (6, 12] saturates to 6): costs the producer a second quantisation of the block
        s1 = torch.exp2(torch.ceil(torch.log2(amax / 6.0)))
        s0 = s1 * 0.5
        e1 = ((e2m1(xb / s1) * s1 - xb) ** 2).sum(-1, keepdim=True)
        e0 = ((e2m1(xb / s0) * s0 - xb) ** 2).sum(-1, keepdim=True)
        sf = torch.where(e0 < e1, s0, s1)
    else:
        # per-tensor power-of-two pre-scale (a compile-time constant per operand in a kernel, as kActHiScale is today)
        # puts the largest block scale at 256 < 448; smaller ones use e4m3's 2^15 range, then subnormals, then flush
        pre = torch.exp2(torch.floor(torch.log2(256.0 / (amax.max() / 6.0))))
        sf = (amax / 6.0 * pre).to(torch.float8_e4m3fn).float().clamp(min=2.0 ** -9) / pre
    q = (e2m1(xb / sf) * sf).reshape(xm.shape)[..., :K]
    return q.movedim(-1, dim)


def mm_mode(a, bt, mode):
    """a [.., M, K] @ bt [.., K, N] with operand rounding per mode, fp32 accumulate."""
    if mode == "fp32":
        return torch.matmul(a, bt)
    ah, al = split16(a)
    bh, bl = split16(bt)
    if mode == "fp16":
        return torch.matmul(ah, bh)
    if mode == "split3":
        return torch.matmul(ah, bh) + torch.matmul(ah, bl) + torch.matmul(al, bh)
    if mode == "split2a":           # activations exact-ish, weights rounded
        return torch.matmul(ah, bh) + torch.matmul(al, bh)
    if mode == "split2b":           # weights exact-ish, activations rounded
        return torch.matmul(ah, bh) + torch.matmul(ah, bl)
    if mode == "f8corr":
        corr = torch.matmul(e4m3(a), e4m3(bl * 65536.0)) + torch.matmul(e4m3(al * 4096.0), e4m3(bt * 16.0))
        return torch.matmul(ah, bh) + corr * (1.0 / 65536.0)
    if mode == "f8c52":            # single accumulator: scale products are 1, all four fp8 operands e5m2
        return (torch.matmul(ah, bh) + torch.matmul(e5m2(a * 2.0 ** -8), e5m2(bl * 2.0 ** 8)) +
                torch.matmul(e5m2(al * 2.0 ** 4), e5m2(bt * 2.0 ** -4)))
    if mode.startswith("f4c"):      # fp16 main + BOTH correction products in block-scaled e2m1 (1.5 tensor-pipe units)
        block, sc = {"f4c": (32, "ue8m0"), "f4c_16": (16, "ue8m0"), "f4c_16b": (16, "ue8m0_best"),
                     "f4c_nv": (16, "ue4m3")}[mode]
        qa = lambda t: blockq4(t, -1, block, sc)      # noqa: E731  (A operands: K is the last dim)
        qb = lambda t: blockq4(t, -2, block, sc)      # noqa: E731  (B^T operands: K is dim -2)
        return torch.matmul(ah, bh) + torch.matmul(qa(a), qb(bl)) + torch.matmul(qa(al), qb(bt))
    if mode.startswith("f84c"):     # A side e5m2 as shipped, B side (weights) block-scaled e2m1 -- bytes, not rate
        qb = lambda t: blockq4(t, -2, 32, "ue8m0")    # noqa: E731
        return torch.matmul(ah, bh) + torch.matmul(e5m2(a * 2.0 ** -8) * 2.0 ** 8, qb(bl)) + torch.matmul(
            e5m2(al * 2.0 ** 4) * 2.0 ** -4, qb(bt))
    if mode == "f8c43":            # a, a_lo in e4m3 (x 2^-4 / x 2^8), b_lo, b in e5m2 (x 2^4 / x 2^-8)
        return (torch.matmul(ah, bh) + torch.matmul(e4m3(a * 2.0 ** -4), e5m2(bl * 2.0 ** 4)) +
                torch.matmul(e4m3(al * 2.0 ** 8), e5m2(bt * 2.0 ** -8)))
    raise ValueError(mode)


class Policy(TorchFunctionMode):
    def __init__(self, lin, attn, n_calls=0):
        """lin = "early>last" runs the first n_calls-1 denoiser calls of the sampler with the `early` linear mode and
        only the last one with `last` (DDIM damps the errors of the early steps)."""
        super().__init__()
        self.lin_early, self.lin_last = (lin.split(">") + [lin])[:2] if ">" in lin else (lin, lin)
        self.lin, self.attn = self.lin_early, attn
        self.n_qk = 0
        self.n_calls, self.call_idx = n_calls, 0

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is F.linear and args[1].shape[1] == 5:        # fusion_layer: a new denoiser call starts
            self.call_idx += 1
            self.lin = self.lin_last if self.call_idx >= self.n_calls else self.lin_early
        if func is F.linear and args[1].shape[1] >= 64 and self.lin != "fp32":
            x, w = args[0], args[1]
            b = args[2] if len(args) > 2 else kwargs.get("bias")
            with torch._C.DisableTorchFunction():
                if self.lin.endswith("_qk1") and w.shape[0] == 1536:
                    # q | k columns: ONE fp16 pass (they are rounded to fp16 for the attention MMAs anyway); v: corrected
                    base = self.lin[:-4]
                    out = torch.cat([mm_mode(x, w[:1024].t(), "fp16"), mm_mode(x, w[1024:].t(), base)], dim=-1)
                else:
                    out = mm_mode(x, w.t(), self.lin[:-4] if self.lin.endswith("_qk1") else self.lin)
                return out + b if b is not None else out
        if func in (torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__) and self.attn != "fp32":
            a, b = args
            with torch._C.DisableTorchFunction():
                # oracle.attention_core: first matmul is q @ k^T, second is (p - I) @ v
                is_qk = (self.n_qk % 2) == 0
                self.n_qk += 1
                if self.attn == "qk3pv1":
                    return mm_mode(a, b, "split3" if is_qk else "fp16")
                if self.attn == "fp16xv":      # shipped kernel: fp16 QK^T and PV, the GRAND "- V" term exact
                    if is_qk:
                        return mm_mode(a, b, "fp16")
                    eye = torch.eye(a.shape[-1], dtype=a.dtype)
                    return mm_mode(a + eye, b, "fp16") - b
                return mm_mode(a, b, self.attn)
        return func(*args, **kwargs)


def main():
    Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 9
    # PROBE_THREADS=1 on oversubscribed VMs: OpenMP barriers there make every small elementwise op 40x slower
    torch.set_num_threads(int(os.environ.get("PROBE_THREADS", os.cpu_count())))
    m = synthetic.make_model(Fr)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x2d, gt = synthetic.make_inputs(B, Fr)
    y_T, steps = synthetic.make_noise(B, Fr, S)
    for clip in (False, True):
        with torch.no_grad():
            ref = oracle.ddim_sample_loop(sd, x2d, y_T, steps, sampling_timesteps=S, clip_denoised=clip)
        print(f"F={Fr} B={B} S={S} clip_denoised={clip}  |ref|max={ref.abs().max():.3f}", flush=True)
        policies = (("fp16", "fp16"), ("split3", "fp16"), ("split3", "split3"), ("f8corr", "fp16"),
                    ("f8corr", "split3"), ("f8corr", "qk3pv1"), ("split2a", "split3"))
        if os.environ.get("PROBE_POLICIES"):     # e.g. PROBE_POLICIES=f8c52:fp16xv,f8c52_qk1:fp16xv
            policies = tuple(tuple(p.split(":")) for p in os.environ["PROBE_POLICIES"].split(","))
        for lin, attn in policies:
            with torch.no_grad(), Policy(lin, attn, S):
                out = oracle.ddim_sample_loop(sd, x2d, y_T, steps, sampling_timesteps=S, clip_denoised=clip)
            err = (out - ref).abs()
            dm = abs(oracle.mpjpe(out, gt).item() - oracle.mpjpe(ref, gt).item())
            print(f"  linear={lin:8s} attn={attn:7s} max-abs {err.max():.3e}  mean-joint-L2 "
                  f"{torch.norm(out - ref, dim=-1).mean():.3e}  |dMPJPE| {dm:.3e}", flush=True)


if __name__ == "__main__":
    main()
