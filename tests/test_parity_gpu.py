"""GPU parity tests: libvvae (through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): fp32 relative error <= 1e-4; bf16 relative error <= 2e-2 on the loss and on
the latent mean / log-variance.  "Relative error" here = max|a - ref| / max|ref| per tensor.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def rel_err(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.fixture(scope="module")
def V():
    import video_vae_b200 as V
    from video_vae_b200 import _ffi
    _ffi.require_device()
    return V


# ---------------------------------------------------------------------------------------------- single operators
def test_gemm_all_modes_fp32_and_bf16(V):
    from video_vae_b200 import ops
    from video_vae_b200._ffi import BACKEND_SIMT, BACKEND_TCGEN05
    g = _gen(0)
    for (M, N, K) in [(300, 200, 136), (130, 72, 64), (64, 3, 12)]:
        A = torch.randn(M, K, generator=g)
        B = torch.randn(K, N, generator=g)
        ref = A @ B
        out = ops.gemm(A.cuda(), B.cuda())
        assert rel_err(out, ref) < FP32_TOL
        out = ops.gemm(A.t().contiguous().cuda(), B.t().contiguous().cuda(), transA=True, transB=True)
        assert rel_err(out, ref) < FP32_TOL
    # bf16: tcgen05 vs SIMT vs fp32 reference of the bf16-rounded operands
    M, N, K = 384, 256, 320
    A = torch.randn(M, K, generator=g).bfloat16()
    B = torch.randn(K, N, generator=g).bfloat16()
    ref = A.float() @ B.float()
    o_tc = ops.gemm(A.cuda(), B.cuda(), backend=BACKEND_TCGEN05, out_dtype=torch.float32)
    o_si = ops.gemm(A.cuda(), B.cuda(), backend=BACKEND_SIMT, out_dtype=torch.float32)
    assert rel_err(o_tc, ref) < 1e-5 and rel_err(o_si, ref) < 1e-5


def test_attention_mask_tests_property(V):
    """train/attention_mask_tests.py: b=17,s=15,h=19,d=13, keys 10..14 masked == truncated to 10; vs oracle too."""
    from oracle import nn as onn
    from video_vae_b200 import ops
    from video_vae_b200.ops import AttnGeom, AttnMask
    b, h, s, d = 17, 19, 15, 13
    g = _gen(1)
    q, k, v = (torch.randn(b, s, h, d, generator=g) for _ in range(3))
    mask = torch.ones(b, h, s, s, dtype=torch.bool)
    mask[..., 10:] = False
    ref = onn.dot_product_attention(q, k, v, mask=mask)
    mu8 = mask.to(torch.uint8).cuda()
    am = AttnMask(mu8, 1, mu8.stride(0), mu8.stride(1), mu8.stride(2), mu8.stride(3))
    qc, kc, vc = (t.reshape(b * s, h * d).cuda() for t in (q, k, v))
    o, _ = ops.attn_fwd(AttnGeom(b, 1, s, s, 0, 1), h, d, qc, kc, vc, am, 1.0 / math.sqrt(d))
    o = o.view(b, s, h, d)
    assert rel_err(o, ref) < FP32_TOL
    qt, kt, vt = (t[:, :10].reshape(b * 10, h * d).contiguous().cuda() for t in (q, k, v))
    ot, _ = ops.attn_fwd(AttnGeom(b, 1, 10, 10, 0, 1), h, d, qt, kt, vt, None, 1.0 / math.sqrt(d))
    assert torch.allclose(o[:, :10].cpu(), ot.view(b, 10, h, d).cpu(), rtol=1e-5, atol=1e-6)


def test_attention_fully_masked_rows_uniform(V):
    from video_vae_b200 import ops
    from video_vae_b200.ops import AttnGeom, AttnMask
    g = _gen(2)
    q, k, v = (torch.randn(2 * 4, 2 * 8, generator=g).cuda() for _ in range(3))
    m = torch.zeros(2, 4, dtype=torch.uint8).cuda()
    o, _ = ops.attn_fwd(AttnGeom(2, 1, 4, 4, 0, 1), 2, 8, q, k, v, AttnMask(m, 1, 4, 0, 0, 1), 1.0)
    ref = v.view(2, 4, 16).mean(dim=1, keepdim=True).expand(2, 4, 16).reshape(8, 16)
    assert torch.allclose(o, ref, atol=1e-5)


def _copy_params(dst, src):
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    missing = dst.load_state_dict(sd, strict=True)
    return missing


def _grads_close(model, oracle, tol, min_checked=1):
    worst, worst_name, n = 0.0, None, 0
    og = dict(oracle.named_parameters())
    for name, p in model.named_parameters():
        ref = og[name].grad
        if ref is None or ref.abs().max() == 0:
            continue
        assert p.grad is not None, f"no gradient for {name}"
        e = rel_err(p.grad, ref)
        n += 1
        if e > worst:
            worst, worst_name = e, name
    assert n >= min_checked
    assert worst < tol, f"worst gradient rel err {worst:.3e} at {worst_name}"
    return worst


@pytest.mark.parametrize("masked", [False, True])
def test_factored_attention_fwd_bwd_fp32(V, masked):
    from oracle import Rngs as ORngs
    from oracle.layers import FactoredAttention as OFA
    o = OFA(256, 768, 4, 128, 32, 16, ORngs(0))
    m = V.FactoredAttention(256, 768, 4, 128, 32, 16, V.Rngs(0), dtype=torch.float32)
    _copy_params(m, o)
    g = _gen(3)
    b, t, hw = 2, 6, 16
    x = torch.randn(b, t, hw, 768, generator=g)
    mask = torch.ones(b, t, dtype=torch.bool)
    if masked:
        mask[0, 4:] = False
        mask[1, 5:] = False
    w = torch.randn(b, t, hw, 768, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = o(xo, mask[:, None, None, :])
    (yo * w).sum().backward()
    xm = x.cuda().requires_grad_(True)
    ym = m(xm, mask[:, None, None, :].cuda())
    (ym * w.cuda()).sum().backward()
    assert rel_err(ym, yo) < FP32_TOL
    assert rel_err(xm.grad, xo.grad) < FP32_TOL
    _grads_close(m, o, 5e-4, min_checked=20)
    # the ((b hw),1,1,t) convention of train/ gives the same result
    from video_vae_b200 import expand_mask
    ym2 = m(x.cuda(), expand_mask(mask, hw).cuda())
    assert torch.equal(ym2, ym.detach())


def test_unet_fwd_bwd_fp32(V):
    from oracle import Rngs as ORngs
    from oracle.unet import UNet as OUNet
    o = OUNet(12, 16, 3, 3, ORngs(0))
    with torch.no_grad():
        o.final_conv.kernel.copy_(torch.randn(o.final_conv.kernel.shape, generator=_gen(9)) * 0.2)
        for n_, p in o.named_parameters():
            if n_.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=_gen(len(n_))) * 0.1)
    m = V.UNet(12, 16, 3, 3, V.Rngs(0), dtype=torch.float32)
    _copy_params(m, o)
    g = _gen(4)
    x = torch.randn(1, 3, 16, 24, 12, generator=g)
    res = torch.randn(1, 3, 16, 24, 3, generator=g)
    w = torch.randn(1, 3, 16, 24, 3, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = o(xo) + res
    (yo * w).sum().backward()
    xm = x.cuda().requires_grad_(True)
    rm = res.cuda().requires_grad_(True)
    ym = m(xm, residual=rm)
    (ym * w.cuda()).sum().backward()
    assert rel_err(ym, yo) < FP32_TOL
    assert rel_err(xm.grad, xo.grad) < 5e-4
    assert rel_err(rm.grad, w) < 1e-6
    _grads_close(m, o, 5e-4, min_checked=30)


def test_unet_bf16_tensor_core_convs_match_generic_and_oracle(V):
    """bf16 U-Net: the tcgen05 implicit-GEMM convs (fwd + dgrad) against the generic kernels on the same bf16
    data (tight) and against the fp32 oracle (bf16 tolerance), forward and every parameter gradient."""
    from oracle import Rngs as ORngs
    from oracle.unet import UNet as OUNet
    from video_vae_b200 import _ffi, ops
    o = OUNet(12, 16, 3, 3, ORngs(0))
    with torch.no_grad():
        o.final_conv.kernel.copy_(torch.randn(o.final_conv.kernel.shape, generator=_gen(9)) * 0.2)
    g = _gen(4)
    shape = (2, 3, 32, 64)
    x = torch.randn(*shape, 12, generator=g)
    res = torch.randn(*shape, 3, generator=g)
    w = torch.randn(*shape, 3, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = o(xo) + res
    (yo * w).sum().backward()
    outs = {}
    for backend in (_ffi.BACKEND_SIMT, _ffi.BACKEND_AUTO):
        ops.CONV_BACKEND = backend
        try:
            m = V.UNet(12, 16, 3, 3, V.Rngs(0), dtype=torch.bfloat16)
            _copy_params(m, o)
            xm = x.cuda().bfloat16().requires_grad_(True)
            ym = m(xm, residual=res.cuda().bfloat16())
            (ym.float() * w.cuda()).sum().backward()
            outs[backend] = (ym.float().cpu(), xm.grad.float().cpu(), {n: p.grad.cpu() for n, p in m.named_parameters()})
        finally:
            ops.CONV_BACKEND = _ffi.BACKEND_AUTO
    ys, dxs, gs = outs[_ffi.BACKEND_SIMT]
    yt, dxt, gt = outs[_ffi.BACKEND_AUTO]
    assert rel_err(yt, ys) < 1e-2 and rel_err(dxt, dxs) < 2e-2
    assert rel_err(yt, yo) < BF16_TOL
    # gradients through 15 bf16 conv+GroupNorm layers: the tensor-core path must be as close to the fp32 oracle as
    # the generic bf16 path is (both carry the same bf16 rounding points)
    assert rel_err(dxt, xo.grad) < max(0.05, 1.5 * rel_err(dxs, xo.grad))
    og = dict(o.named_parameters())
    for n_, gg in gt.items():
        assert torch.isfinite(gg).all(), n_
        assert rel_err(gg, og[n_].grad) < max(0.05, 1.5 * rel_err(gs[n_], og[n_].grad)), n_


def _small_pair(V, dtype, enc=2, dec=2, seed=2):
    from oracle import Rngs as ORngs
    from oracle.model import VideoVAE as OVAE
    o = OVAE(64, 64, 3, 16, enc, dec, 256, 4, 128, 32, 8, 4, ORngs(seed))
    with torch.no_grad():
        o.decoder.unet.final_conv.kernel.copy_(torch.randn(o.decoder.unet.final_conv.kernel.shape, generator=_gen(5)) * 0.05)
    m = V.VideoVAE(64, 64, 3, 16, enc, dec, 256, 4, 128, 32, 8, 4, V.Rngs(seed), dtype=dtype)
    _copy_params(m, o)
    return m, o


def _inputs(b=2, t=4, hw=16, lat=96):
    g = _gen(6)
    video = torch.rand(b, t, 64, 64, 3, generator=g)
    mask = torch.ones(b, t, dtype=torch.bool)
    mask[0, 3:] = False
    noise = torch.randn(b, t, hw, lat, generator=g)
    u = torch.rand(b, t, 1, generator=g)
    return video, mask, noise, u


def test_videovae_loss_and_grads_fp32(V):
    """cfg-1-style correctness: whole model forward + loss + backward, fp32, vs the oracle (reduced depth)."""
    from oracle import Rngs as ORngs
    from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn as o_loss_fn
    m, o = _small_pair(V, torch.float32)
    video, mask, noise, u = _inputs()
    lo, auxo = o_loss_fn(o, video, expand_mask(mask, 16), mask, ORngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=u)
    lo.backward()
    lm, auxm = V.loss_fn(m, video.cuda(), mask[:, None, None, :].cuda(), mask.cuda(), V.Rngs(0), V.DEFAULT_HPARAMS,
                         noise=noise.cuda(), gumbel_u=u.cuda())
    lm.backward()
    assert torch.equal(auxm["selection"].cpu().reshape(-1), auxo["selection"].reshape(-1))
    for key in ("mean", "logvar", "reconstruction", "compressed"):
        assert rel_err(auxm[key], auxo[key]) < FP32_TOL, key
    for key in ("MSE", "kl_loss", "selection_loss", "MAE"):
        assert abs(auxm[key].item() - auxo[key].item()) <= FP32_TOL * max(abs(auxo[key].item()), 1e-6), key
    assert abs(lm.item() - lo.item()) <= FP32_TOL * abs(lo.item())
    _grads_close(m, o, 1e-3, min_checked=100)


def test_videovae_bf16_loss_and_latents(V):
    from oracle import Rngs as ORngs
    from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn as o_loss_fn
    m, o = _small_pair(V, torch.bfloat16)
    video, mask, noise, u = _inputs()
    lo, auxo = o_loss_fn(o, video, expand_mask(mask, 16), mask, ORngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=u)
    lm, auxm = V.loss_fn(m, video.cuda(), mask[:, None, None, :].cuda(), mask.cuda(), V.Rngs(0), V.DEFAULT_HPARAMS,
                         noise=noise.cuda(), gumbel_u=u.cuda())
    lm.backward()
    assert auxm["mean"].dtype == torch.bfloat16 and auxm["reconstruction"].dtype == torch.bfloat16
    assert rel_err(auxm["mean"], auxo["mean"]) < BF16_TOL
    assert rel_err(auxm["logvar"], auxo["logvar"]) < BF16_TOL
    if torch.equal(auxm["selection"].cpu().reshape(-1), auxo["selection"].reshape(-1)):
        assert abs(lm.item() - lo.item()) <= BF16_TOL * abs(lo.item())
    for n_, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n_


def test_eval_mode_and_encoder_masked_equals_truncated(V):
    """train/llm_tests.py:390-474: masked frames do not change the unmasked frames' latents (fp32, depth 2)."""
    m, _ = _small_pair(V, torch.float32)
    g = _gen(7)
    x = (torch.randn(2, 8, 64, 64, 3, generator=g) * 0.02).cuda()
    mask = torch.ones(2, 8, dtype=torch.bool).cuda()
    mask[:, 5:] = False
    with torch.no_grad():
        mu_m, lv_m, _ = m.encoder(x, mask[:, None, None, :], V.Rngs(1), train=False)
        mu_c, lv_c, _ = m.encoder(x[:, :5].contiguous(), mask[:, None, None, :5], V.Rngs(1), train=False)
    assert torch.allclose(mu_m[:, :5], mu_c, atol=5e-2) and torch.allclose(lv_m[:, :5], lv_c, atol=5e-2)


def test_philox_noise_statistics_and_determinism(V):
    m, _ = _small_pair(V, torch.float32, enc=1, dec=1)
    video, mask, _, _ = _inputs()
    with torch.no_grad():
        out1 = m(video.cuda(), mask[:, None, None, :].cuda(), V.Rngs(3), train=True)
        out2 = m(video.cuda(), mask[:, None, None, :].cuda(), V.Rngs(3), train=True)
        out3 = m(video.cuda(), mask[:, None, None, :].cuda(), V.Rngs(4), train=True)
    assert torch.equal(out1[1], out2[1])
    assert not torch.equal(out1[1], out3[1])
    sel = out1[2]
    assert set(torch.unique(sel).tolist()) <= {0.0, 1.0}
