"""Bring-up: HBM-bound kernels (LayerNorm, QK-norm + RoPE generic and hd64 fast path, column sums) vs torch references,
with timings at the production shape.  Run on the B200 box; small=1 keeps shapes tiny (for compute-sanitizer)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import ops  # noqa: E402

SMALL = len(sys.argv) > 1 and sys.argv[1] == "small"
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9)).item()


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


def ln_case(rows, D):
    x = torch.randn(rows, D, device=dev, generator=g).bfloat16()
    dy = torch.randn(rows, D, device=dev, generator=g).bfloat16()
    dres = torch.randn(rows, D, device=dev, generator=g).bfloat16()
    gam = torch.randn(D, device=dev, generator=g)
    bet = torch.randn(D, device=dev, generator=g)
    y, mean, rstd = ops.layernorm_fwd(x, gam, bet)
    xr = x.float().requires_grad_()
    gr = gam.clone().requires_grad_()
    br = bet.clone().requires_grad_()
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps=1e-6)
    yr.backward(dy.float())
    dgam = torch.zeros(D, device=dev)
    dbet = torch.zeros(D, device=dev)
    dx = ops.layernorm_bwd(dy, x, mean, rstd, gam, dres, dgam, dbet)
    torch.cuda.synchronize()
    r = dict(op="layernorm", rows=rows, D=D, y=rel(y, yr), dx=rel(dx, xr.grad + dres.float()), dgamma=rel(dgam, gr.grad),
             dbeta=rel(dbet, br.grad))
    if not SMALL:
        r["fwd_ms"] = timeit(lambda: ops.layernorm_fwd(x, gam, bet))
        r["bwd_ms"] = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres, dgam, dbet))
        r["fwd_GBs"] = 2 * rows * D * 2 / r["fwd_ms"] / 1e6
        r["bwd_GBs"] = 4 * rows * D * 2 / r["bwd_ms"] / 1e6
    emit(**r)


def qk_reference(qkv, qs, ks, cos, sin, H, hd, pos):
    """fp32 reference with the bf16 rounding points of the kernel (x~ and each product rounded)."""
    rows = qkv.shape[0]
    x = qkv[:, :2 * H * hd].float().reshape(rows, 2, H, hd)
    mu = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mu * mu
    xh = (x - mu) * torch.rsqrt(var.clamp_min(0) + 1e-6)
    sc = torch.stack([qs, ks])[None, :, None, :]
    xn = (xh * sc).bfloat16().float()
    c = cos.float()[pos][:, None, None, :]
    s = sin.float()[pos][:, None, None, :]
    half = hd // 2
    rot = torch.cat([-xn[..., half:], xn[..., :half]], -1)
    y = (xn * c).bfloat16().float() + (rot * s).bfloat16().float()
    return y.reshape(rows, 2 * H * hd)


def qk_case(rows, H, hd, pos_div, pos_mod):
    qkv = torch.randn(rows, 3 * H * hd, device=dev, generator=g).bfloat16()
    qs = 1 + 0.1 * torch.randn(hd, device=dev, generator=g)
    ks = 1 + 0.1 * torch.randn(hd, device=dev, generator=g)
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd))
    fr = torch.arange(pos_mod, device=dev).float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos().bfloat16().contiguous(), emb.sin().bfloat16().contiguous()
    pos = (torch.arange(rows, device=dev) // pos_div) % pos_mod
    out = ops.qknorm_rope_fwd(qkv, qs, ks, cos, sin, H, hd, pos_div, pos_mod)
    # reference through autograd for the backward
    qkv_r = qkv.float().requires_grad_()
    qs_r, ks_r = qs.clone().requires_grad_(), ks.clone().requires_grad_()
    x = qkv_r[:, :2 * H * hd].reshape(rows, 2, H, hd)
    mu = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mu * mu
    xn = (x - mu) * torch.rsqrt(var.clamp_min(0) + 1e-6) * torch.stack([qs_r, ks_r])[None, :, None, :]
    c = cos.float()[pos][:, None, None, :]
    s = sin.float()[pos][:, None, None, :]
    half = hd // 2
    y_r = xn * c + torch.cat([-xn[..., half:], xn[..., :half]], -1) * s
    dy = torch.randn(rows, 2 * H * hd, device=dev, generator=g).bfloat16()
    y_r.reshape(rows, -1).backward(dy.float())
    dqkv = torch.zeros(rows, 3 * H * hd, device=dev, dtype=torch.bfloat16)
    dqkv[:, :2 * H * hd] = dy
    dqs, dks = torch.zeros(hd, device=dev), torch.zeros(hd, device=dev)
    ops.qknorm_rope_bwd_(dqkv, qkv, qs, ks, cos, sin, dqs, dks, H, hd, pos_div, pos_mod)
    torch.cuda.synchronize()
    r = dict(op="qknorm_rope", rows=rows, H=H, hd=hd, y=rel(out, qk_reference(qkv, qs, ks, cos, sin, H, hd, pos)),
             dx=rel(dqkv[:, :2 * H * hd], qkv_r.grad[:, :2 * H * hd]), dqs=rel(dqs, qs_r.grad), dks=rel(dks, ks_r.grad))
    if not SMALL:
        r["fwd_ms"] = timeit(lambda: ops.qknorm_rope_fwd(qkv, qs, ks, cos, sin, H, hd, pos_div, pos_mod))
        r["bwd_ms"] = timeit(lambda: ops.qknorm_rope_bwd_(dqkv, qkv, qs, ks, cos, sin, dqs, dks, H, hd, pos_div, pos_mod))
        r["fwd_GBs"] = 2 * rows * 2 * H * hd * 2 / r["fwd_ms"] / 1e6
        r["bwd_GBs"] = 3 * rows * 2 * H * hd * 2 / r["bwd_ms"] / 1e6
    emit(**r)


def colsum_case(rows, n, ld=None):
    ld = ld or n
    base = torch.randn(rows, ld, device=dev, generator=g).bfloat16()
    x = base[:, :n]
    out = torch.zeros(n, device=dev)
    ops.colsum_accum(x, out)
    torch.cuda.synchronize()
    r = dict(op="colsum", rows=rows, n=n, ld=ld, err=rel(out, x.float().sum(0)))
    if not SMALL:
        r["ms"] = timeit(lambda: ops.colsum_accum(x, out))
        r["GBs"] = rows * n * 2 / r["ms"] / 1e6
    emit(**r)


R = 600 if SMALL else 32768
ln_case(R, 768)
ln_case(R // 2 + 3, 256)
qk_case(R, 8, 64, 1, 256)        # spatial: position = row % 256 (fast path)
qk_case(R, 8, 64, 256, 16)       # temporal: position = (row / hw) % t
qk_case(R // 4 * 4, 4, 32, 1, 16)  # generic kernel
qk_case(R // 4 * 4, 2, 64, 16, 8)  # fast path, 2 heads
colsum_case(R, 1536)
colsum_case(R, 768)
colsum_case(R + 5, 96)
colsum_case(R, 512, ld=1536)
