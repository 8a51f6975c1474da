"""CPU tests of the oracle: pinned against the float64 numpy restatement (oracle/np_ref.py)
and against the reference's own property tests."""
import numpy as np
import pytest
import torch

from oracle import Rngs, np_ref
from oracle import nn as onn
from oracle.layers import (Attention, FactoredAttention, GumbelSigmoidSTE, PatchEmbedding, PatchUnEmbedding,
                           RotaryEmbedding, round_ste)
from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn
from oracle.model import Encoder, VideoVAE
from oracle.unet import UNet

F64 = dict(dtype=torch.float64, param_dtype=torch.float64)


def _rand(*shape, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def test_layernorm_vs_numpy():
    x = _rand(5, 7, 48, seed=1)
    s, b = _rand(48, seed=2), _rand(48, seed=3)
    y = onn.layer_norm(x, s, b, torch.float64)
    np.testing.assert_allclose(y.numpy(), np_ref.layer_norm(x.numpy(), s.numpy(), b.numpy()), rtol=1e-12, atol=1e-12)


def test_groupnorm_vs_numpy():
    gn = onn.GroupNorm(4, 16, **F64)
    with torch.no_grad():
        gn.scale.copy_(_rand(16, seed=4))
        gn.bias.copy_(_rand(16, seed=5))
    x = _rand(2, 3, 4, 6, 16, seed=6)
    ref = np_ref.group_norm(x.numpy(), 4, gn.scale.detach().numpy(), gn.bias.detach().numpy())
    np.testing.assert_allclose(gn(x).detach().numpy(), ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("ks", [(3, 3, 3), (3, 7, 7), (1, 1, 1)])
def test_conv_vs_numpy(ks):
    conv = onn.Conv(5, 6, ks, Rngs(0), **F64)
    with torch.no_grad():
        conv.bias.copy_(_rand(6, seed=7))
    x = _rand(2, 4, 9, 8, 5, seed=8)
    ref = np_ref.conv3d_same(x.numpy(), conv.kernel.detach().numpy(), conv.bias.detach().numpy())
    np.testing.assert_allclose(conv(x).detach().numpy(), ref, rtol=1e-10, atol=1e-10)


def test_conv_transpose_vs_numpy():
    ct = onn.ConvTranspose122(5, 3, Rngs(1), **F64)
    with torch.no_grad():
        ct.bias.copy_(_rand(3, seed=9))
    x = _rand(2, 3, 4, 5, 5, seed=10)
    ref = np_ref.conv_transpose_122(x.numpy(), ct.kernel.detach().numpy(), ct.bias.detach().numpy())
    np.testing.assert_allclose(ct(x).detach().numpy(), ref, rtol=1e-10, atol=1e-10)


def test_maxpool_rope_softplus_vs_numpy():
    x = _rand(2, 3, 6, 8, 5, seed=11)
    np.testing.assert_allclose(onn.max_pool_122(x).numpy(), np_ref.max_pool_122(x.numpy()))
    q = _rand(3, 11, 4, 16, seed=12)
    rope = RotaryEmbedding(16, max_len=32)
    rope.cos_cached, rope.sin_cached = rope.cos_cached.double(), rope.sin_cached.double()
    qr, _ = rope.rotate_queries_and_keys(q, q)
    np.testing.assert_allclose(qr.numpy(), np_ref.rope(q.numpy()), rtol=1e-5, atol=1e-6)  # tables built in fp32
    z = _rand(100, seed=13) * 10
    np.testing.assert_allclose(onn.softplus(z).numpy(), np_ref.softplus(z.numpy()), rtol=1e-12)


def test_attention_vs_numpy_and_mask_property():
    """train/attention_mask_tests.py: masking keys 10..14 == truncating to 10 (same odd shapes)."""
    b, h, s, d = 17, 19, 15, 13
    q, k, v = _rand(b, s, h, d, seed=1), _rand(b, s, h, d, seed=2), _rand(b, s, h, d, seed=3)
    mask = torch.ones(b, h, s, s, dtype=torch.bool)
    mask[..., 10:] = False
    om = onn.dot_product_attention(q, k, v, mask=mask)
    np.testing.assert_allclose(om.numpy(), np_ref.attention(q.numpy(), k.numpy(), v.numpy(), mask.numpy()),
                               rtol=1e-10, atol=1e-12)
    ou = onn.dot_product_attention(q[:, :10], k[:, :10], v[:, :10])
    assert torch.allclose(om[:, :10], ou, rtol=1e-5, atol=1e-8)


def test_attention_fully_masked_row_is_uniform():
    q, k, v = _rand(2, 4, 2, 8, seed=1), _rand(2, 4, 2, 8, seed=2), _rand(2, 4, 2, 8, seed=3)
    mask = torch.zeros(2, 1, 1, 4, dtype=torch.bool)
    o = onn.dot_product_attention(q, k, v, mask=mask)
    assert torch.allclose(o, v.mean(dim=1, keepdim=True).expand_as(o), atol=1e-12)


def _small_vae(dtype=torch.float32, enc=2, dec=2, seed=2):
    return VideoVAE(64, 64, 3, 16, enc, dec, 256, 4, 128, 32, 8, 4, Rngs(seed), dtype=dtype, param_dtype=torch.float32)


def test_shapes_test_rl_model_config():
    """claude_distributed/test_rl_model.py:49-139,193-239 shapes (64x64, P16, 2/2, mlp 256, 4 heads, qkv 128)."""
    rngs = Rngs(0)
    pe = PatchEmbedding(64, 64, 3, 16, rngs)
    x = _rand(2, 4, 64, 64, 3, dtype=torch.float32)
    e = pe(x)
    assert e.shape == (2, 4, 16, 768)
    feats, rgb = PatchUnEmbedding(64, 64, 3, 16, 4, rngs)(e)
    assert feats.shape == (2, 4, 64, 64, 12) and rgb.shape == (2, 4, 64, 64, 3)
    fa = FactoredAttention(256, 768, 4, 128, 32, 16, rngs)
    assert fa(e, torch.ones(2 * 16, 1, 1, 4, dtype=torch.bool)).shape == e.shape
    assert UNet(12, 16, 3, 3, rngs)(feats).shape == (2, 4, 64, 64, 3)
    vae = _small_vae()
    mask = torch.ones(2, 4, dtype=torch.bool)
    rec, comp, sel, lv, mu = vae(x, expand_mask(mask, 16), Rngs(3), train=True)
    assert rec.shape == x.shape and comp.shape == (2, 4, 16, 96) and sel.shape == (2, 4, 1, 1)
    assert lv.shape == mu.shape == (2, 4, 16, 96)
    assert set(torch.unique(sel).tolist()) <= {0.0, 1.0}


def test_mask_conventions_agree():
    """train/ passes ((b hw),1,1,t); claude_distributed/ passes (b,1,1,t) (layers.py:213-214)."""
    fa = FactoredAttention(64, 48, 2, 32, 8, 4, Rngs(0))
    x = _rand(2, 6, 4, 48, dtype=torch.float32)
    m = torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    assert torch.allclose(fa(x, expand_mask(m, 4)), fa(x, m[:, None, None, :]))


@pytest.mark.parametrize("depth,atol", [(1, 5e-3), (2, 5e-2)])
def test_encoder_masked_equals_truncated(depth, atol):
    """train/llm_tests.py:390-474,499-502 (fp32 tolerances quoted there)."""
    enc = Encoder(64, 64, 3, 16, depth, 256, 4, 128, 32, 8, Rngs(0))
    x = _rand(2, 8, 64, 64, 3, dtype=torch.float32) * 0.02
    m = torch.ones(2, 8, dtype=torch.bool)
    m[:, 5:] = False
    mu_m, lv_m, _ = enc(x, expand_mask(m, 16), Rngs(1), train=False)
    mu_c, lv_c, _ = enc(x[:, :5], expand_mask(m[:, :5], 16), Rngs(1), train=False)
    assert torch.allclose(mu_m[:, :5], mu_c, atol=atol) and torch.allclose(lv_m[:, :5], lv_c, atol=atol)


def test_batch_isolation():
    """train/human_tests.py:84-92 (atol 1e-1 there; exact here up to fp32 noise)."""
    vae = _small_vae(enc=1, dec=1)
    x = _rand(3, 4, 64, 64, 3, dtype=torch.float32) * 0.02
    m = torch.ones(3, 4, dtype=torch.bool)
    full = vae(x, expand_mask(m, 16), Rngs(0), train=False)[0]
    one = vae(x[1:2], expand_mask(m[1:2], 16), Rngs(0), train=False)[0]
    assert torch.allclose(full[1:2], one, atol=1e-4)


def test_round_ste_and_gumbel():
    """claude_distributed/test_rl_model.py:173-191."""
    x = torch.tensor([0.2, 0.5, 1.5, 2.5, -0.7], requires_grad=True)
    y = round_ste(x)
    assert y.tolist() == [0.0, 0.0, 2.0, 2.0, -1.0]  # half to even
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    g = GumbelSigmoidSTE()
    lg = torch.randn(64, requires_grad=True)
    s = g(lg, Rngs(0), train=True)
    assert set(torch.unique(s.detach()).tolist()) <= {0.0, 1.0}
    s.sum().backward()
    assert lg.grad.abs().sum() > 0
    assert torch.equal(g(torch.tensor([-1.0, 0.0, 3.0]), Rngs(0), train=False), torch.tensor([0.0, 0.0, 1.0]))


def test_loss_finite_grads_nonzero_and_decreases():
    """claude_distributed/test_rl_model.py:151-171, test_training_loop.py:137-202."""
    torch.manual_seed(0)
    vae = _small_vae(enc=1, dec=1)
    with torch.no_grad():
        variance = vae.decoder.unet.final_conv.kernel
        variance.copy_(torch.randn_like(variance) * 0.05)   # exercise the UNet backward (zero-init otherwise)
    x = torch.rand(2, 4, 64, 64, 3)
    m = torch.tensor([[1, 1, 1, 0], [1, 1, 1, 1]], dtype=torch.bool)
    noise = torch.randn(2, 4, 16, 96)
    u = torch.rand(2, 4, 1)
    opt = torch.optim.Adam(vae.parameters(), lr=1e-3)
    losses = []
    for _ in range(10):
        opt.zero_grad()
        loss, aux = loss_fn(vae, x, expand_mask(m, 16), m, Rngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=u)
        loss.backward()
        if not losses:
            assert torch.isfinite(loss) and loss > 0
            for n, p in vae.named_parameters():
                assert p.grad is not None and torch.isfinite(p.grad).all(), n
            assert vae.decoder.unet.patch_mixer.kernel.grad.abs().sum() > 0
            assert vae.encoder.layers[0].TemporalAttention.q_norm.scale.grad.abs().sum() > 0
        torch.nn.utils.clip_grad_norm_(vae.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_oracle_reproduces_committed_golden_fixture():
    """tests/golden/videovae_cfg64_fp32.npz was written by tests/golden/make_golden.py from this oracle; a change in
    the oracle's arithmetic shows up here (CPU, fp32; BLAS reduction order differs between hosts -> rtol 2e-5)."""
    import importlib.util
    import os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "videovae_cfg64_fp32.npz"))
    now = mg.run_oracle()
    now.update(mg.run_attention_kat())
    assert sorted(now) == sorted(gold.files)
    for key in gold.files:
        if key == "grad_names":
            assert list(now[key]) == list(gold[key])
        elif key == "selection":
            assert np.array_equal(now[key], gold[key])
        else:
            ref = np.asarray(gold[key], dtype=np.float64)
            got = np.asarray(now[key], dtype=np.float64)
            assert np.allclose(got, ref, rtol=2e-4, atol=2e-5 * max(1e-6, np.abs(ref).max())), key


def test_optimizer_oracle_matches_torch_adam_with_global_norm_clip():
    """oracle/optim.ClipAdam restates optax.chain(clip_by_global_norm, adam) (rl_nonadversarial.py:248-251); pin it
    against torch.optim.Adam (same bias-corrected update, eps outside the sqrt) + clip_grad_norm_."""
    from oracle.optim import ClipAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(7, 5), (13,), (3, 3, 3)]
    p_o = [torch.randn(s, generator=g) for s in shapes]
    p_t = [torch.nn.Parameter(p.clone()) for p in p_o]
    opt_o = ClipAdam(p_o, lr=1e-2, clip=1.0)
    opt_t = torch.optim.Adam(p_t, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    for step in range(6):
        scale = 3.0 if step % 2 == 0 else 0.05                   # clip active on even steps, inactive on odd ones
        grads = [torch.randn(s, generator=g) * scale for s in shapes]
        gn = opt_o.step([x.clone() for x in grads])
        for p, x in zip(p_t, grads):
            p.grad = x.clone()
        torch.nn.utils.clip_grad_norm_(p_t, 1.0)                 # torch adds 1e-6 to the norm: tolerance below
        opt_t.step()
        assert (gn > 1.0) == (step % 2 == 0)
        for a, b in zip(p_o, p_t):
            assert torch.allclose(a, b.detach(), rtol=1e-4, atol=1e-6)


def test_warmup_cosine_schedule_known_points():
    """optax.warmup_cosine_decay_schedule(0, peak, warmup, decay, peak/10) (rl_nonadversarial.py:241-247): closed-form
    points, and the host implementation the product uses must agree with the oracle everywhere."""
    import importlib.util
    import math
    import os
    from oracle.optim import warmup_cosine_decay_schedule as o_sched
    spec = importlib.util.spec_from_file_location(
        "vvae_optim", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video_vae_b200", "optim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    peak, warm, decay = 2e-5, 20000 // math.sqrt(2), 1_000_000
    so, sp = o_sched(0.0, peak, warm, decay, peak / 10), mod.reference_schedule(2)
    assert so(0) == 0.0 and abs(so(warm / 2) - peak / 2) < 1e-12 and abs(so(warm) - peak) < 1e-12
    mid = warm + (decay - warm) / 2
    assert abs(so(mid) - peak * (0.9 * 0.5 + 0.1)) < 1e-12
    assert abs(so(decay) - peak / 10) < 1e-12 and abs(so(10 * decay) - peak / 10) < 1e-12
    for c in (0, 1, 17, warm - 1, warm, warm + 1, 12345, 500000, decay, decay + 5):
        assert so(c) == sp(c)


def test_rl_model_variant_shapes_and_binary_keep_mask():
    """claude_distributed/test_rl_model.py Tests 1-3 on the oracle's rl_model restatement: encoder selection is a
    probability of shape (b, t, 1); the VAE doubles the batch, the keep-mask is binary, 6 outputs."""
    from oracle import Rngs
    from oracle.rl_model import Encoder, VideoVAE
    B, T, H = 2, 4, 64
    enc = Encoder(H, H, 3, 16, 2, 256, 4, 128, 32, 8, Rngs(42))
    x = torch.randn(B, T, H, H, 3, generator=torch.Generator().manual_seed(0)) * 0.02
    mask = torch.ones(B, 1, 1, T, dtype=torch.bool)
    mean, logvar, sel = enc(x, mask, Rngs(0), train=True)
    assert mean.shape == (B, T, 16, 96) and logvar.shape == mean.shape and sel.shape == (B, T, 1)
    assert (sel > 0).all() and (sel < 1).all()
    vae = VideoVAE(H, H, 3, 16, 2, 2, 256, 4, 128, 32, 8, 4, Rngs(42))
    recon, comp, s, smask, lv, mu = vae(x, mask, Rngs(42), train=True)
    assert recon.shape == (2 * B, T, H, H, 3) and comp.shape == (2 * B, T, 16, 96)
    assert s.shape == (2 * B, T, 1, 1) and smask.shape == (2 * B, T, 1, 1) and lv.shape == mu.shape == comp.shape
    assert set(smask.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(mu[0], mu[1]) and torch.equal(s[2], s[3])            # 'b ... -> (b 2) ...': copies are adjacent
    recon.square().mean().backward()
    g = vae.encoder.spatial_compression.kernel.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0


def test_distributed_rl_model_returns_variance():
    """claude_distributed/rl_model.py:55-60,125-128,147 (the data-parallel trainer's copy): same weights, same draws ->
    the same reconstruction as train/rl_model.py, with variance = exp(log_variance) in the 5th slot."""
    from oracle import Rngs
    from oracle.distributed_rl_model import Encoder, VideoVAE as DVAE
    from oracle.rl_model import VideoVAE as RVAE
    B, T, H = 2, 4, 64
    cfg = (H, H, 3, 16, 2, 2, 256, 4, 128, 32, 8, 4)
    a, d = RVAE(*cfg, Rngs(42)), DVAE(*cfg, Rngs(42))
    d.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, H, H, 3, generator=g) * 0.5
    mask = torch.ones(B, 1, 1, T, dtype=torch.bool)
    noise, bu = torch.randn(B, T, 16, 96, generator=g), torch.rand(2 * B, T, 1, 1, generator=g)
    ra = a(x, mask, Rngs(0), train=True, noise=noise, bernoulli_u=bu)
    rd = d(x, mask, Rngs(0), train=True, noise=noise, bernoulli_u=bu)
    assert (rd[4] > 0).all() and torch.allclose(rd[4], torch.exp(ra[4]), rtol=1e-5, atol=0)
    for i in (0, 1, 2, 3, 5):
        assert ((rd[i] - ra[i]).abs().max() / ra[i].abs().max()).item() < 1e-5, i
    mean, var, sel = Encoder(H, H, 3, 16, 2, 256, 4, 128, 32, 8, Rngs(1))(x, mask, Rngs(0))
    assert var.shape == mean.shape and (var > 0).all() and sel.shape == (B, T, 1)


def test_plain_eval_step_is_deterministic_and_uses_the_mean():
    """training_loop_adversarial.py:139-148: eval_step runs loss_fn with train=False -- no draws: two calls with
    different rngs agree, the gate is binary, the compressed representation is the mean on kept frames."""
    from oracle import Rngs
    from oracle.losses import DEFAULT_HPARAMS, eval_step
    from oracle.model import VideoVAE
    cfg = (64, 64, 3, 16, 1, 1, 128, 2, 64, 16, 8, 4)
    m = VideoVAE(*cfg, Rngs(3))
    g = torch.Generator().manual_seed(1)
    video = torch.rand(2, 4, 64, 64, 3, generator=g)
    mask = torch.tensor([[True] * 4, [True, True, False, False]])
    l1, a1 = eval_step(m, video, mask, DEFAULT_HPARAMS, 16, Rngs(0))
    l2, a2 = eval_step(m, video, mask, DEFAULT_HPARAMS, 16, Rngs(99))
    assert torch.equal(l1, l2) and torch.equal(a1["reconstruction"], a2["reconstruction"]) and not l1.requires_grad
    sel = a1["selection"]
    assert set(sel.unique().tolist()) <= {0.0, 1.0}
    kept = sel.reshape(2, 4) > 0
    assert torch.equal(a1["compressed"][kept], a1["mean"][kept])


def test_rl_loss_training_loop_checks():
    """claude_distributed/test_training_loop.py Tests 1, 3, 4 (64x64, P16, 2/2, mlp 256, 4 heads, qkv 128, scr 4, up 2,
    t=8, lr 1e-3, gamma3=0 with a zero perceptual term) on the oracle restatement of rl_nonadversarial.py:100-186:
    finite positive loss and aux, finite non-zero gradients, loss decreasing over 10 clip+Adam steps."""
    from oracle import Rngs
    from oracle.optim import ClipAdam
    from oracle.rl_losses import loss_fn
    from oracle.rl_model import VideoVAE
    hp = {"gamma1": 0.2, "gamma2": 0.001, "gamma3": 0.0, "gamma4": 0.05, "max_compression_rate": 2,
          "magnify_negatives_rate": 100, "rl_loss_weight": 0.01}
    vae = VideoVAE(64, 64, 3, 16, 2, 2, 256, 4, 128, 16, 4, 2, Rngs(42))
    g = torch.Generator().manual_seed(0)
    B, T = 2, 8
    video = torch.randn(B, T, 64, 64, 3, generator=g) * 0.1
    mask = torch.ones(B, T, dtype=torch.bool)
    zero_perceptual = lambda params, x, target: torch.zeros(x.shape[0])       # noqa: E731  dummy_perceptual (:81-82)
    loss, aux = loss_fn(vae, video, mask[:, None, None, :], mask, Rngs(1), hp, zero_perceptual, None)
    assert torch.isfinite(loss) and loss > 0
    for k in ("MSE", "selection_loss", "kl_loss", "rl_loss", "per_sample_MAE", "kept_frame_density",
              "mean_trajectory_prob"):
        assert torch.isfinite(aux[k]), k
    # normalised probs are 1 in value: the RL term is the mean disadvantage, which is 0 for every pair
    assert abs(float(aux["rl_loss"].detach())) < 1e-3
    loss.backward()
    grads = [p.grad for p in vae.parameters() if p.grad is not None]
    assert all(torch.isfinite(x).all() for x in grads) and max(float(x.abs().max()) for x in grads) > 0
    # the gate logits get gradient only through the trajectory-probability term
    assert vae.encoder.selection_layer2.kernel.grad.abs().max() > 0
    params = [p for p in vae.parameters()]
    opt = ClipAdam(params, lr=1e-3, clip=1.0)
    losses = []
    for step in range(10):
        for p in params:
            p.grad = None
        loss, _ = loss_fn(vae, video, mask[:, None, None, :], mask, Rngs(step + 100), hp, zero_perceptual, None)
        loss.backward()
        opt.step([p.grad if p.grad is not None else torch.zeros_like(p) for p in params])
        losses.append(float(loss))
    assert sum(losses[5:]) < sum(losses[:5]), losses


def test_rl_loss_pairwise_terms_hand_checked():
    """rl_nonadversarial.py:149-174 on a hand-sized case: disadvantages of a pair are +-1 (population std), the gradient
    of the RL term w.r.t. a frame's keep-probability is weight * disadvantage * sign / P(action) / (b*2), masked frames
    and clipped probabilities get none."""
    from oracle.rl_losses import loss_terms
    hp = {"gamma1": 0.0, "gamma2": 0.0, "gamma3": 0.0, "gamma4": 0.0, "max_compression_rate": 2,
          "magnify_negatives_rate": 100, "rl_loss_weight": 0.5}
    b, t = 1, 3
    video = torch.zeros(b, t, 2, 2, 1)
    recon = torch.zeros(2 * b, t, 2, 2, 1)
    recon[0] += 1.0                                                    # sample 0 is worse than its twin
    sel = torch.tensor([0.3, 0.6, 0.0]).repeat(2, 1).reshape(2, t, 1, 1).requires_grad_()
    smask = torch.tensor([[1.0, 0.0, 1.0], [0.0, 1.0, 1.0]]).reshape(2, t, 1, 1)
    mask = torch.tensor([[True, True, True]])
    lv = torch.zeros(2, t, 1, 4)
    loss, aux = loss_terms(video, recon, sel, smask, lv, lv.clone(), mask, hp)
    assert abs(float(aux["MSE"]) - 0.5) < 1e-6 and abs(float(loss) - 0.5) < 1e-5
    loss.backward()
    gsel = sel.grad.reshape(2, t)
    w = 0.5 / 2
    # sample 0 (disadvantage +1): kept frame 0 (P=0.3, d/dp=+1), dropped frame 1 (P=0.4, d/dp=-1), frame 2 clipped
    exp0 = torch.tensor([w / 0.3, -w / 0.4, 0.0])
    exp1 = torch.tensor([-w * -1 / 0.7, -w / 0.6, 0.0])               # disadvantage -1: dropped (P=.7), kept (P=.6)
    assert torch.allclose(gsel[0], exp0, rtol=1e-4, atol=1e-6), gsel
    assert torch.allclose(gsel[1], exp1, rtol=1e-4, atol=1e-6), gsel
    # a masked frame contributes nothing
    sel2 = sel.detach().clone().requires_grad_()
    loss2, _ = loss_terms(video, recon, sel2, smask, lv, lv.clone(), torch.tensor([[True, False, True]]), hp)
    loss2.backward()
    assert float(sel2.grad.reshape(2, t)[:, 1].abs().max()) == 0.0


def test_vgg_perceptual_oracle_shapes_and_loss():
    """train/vgg_tests.py __main__ checks (:134-201) on the oracle restatement: activation shapes at 64x64 and 128x128,
    a finite scalar loss for ones vs 0.5*ones, per-sample form averaging to the scalar form, finite non-zero gradient
    of the input's shape; plus a hand-check of the ImageNet normalisation and 'SAME' zero padding on relu1_1."""
    from oracle import Rngs
    from oracle.perceptual import (IMAGENET_MEAN, IMAGENET_STD, VGG16Features, get_adversarial_perceptual_loss_fn,
                                   get_perceptual_loss_fn)
    vgg = VGG16Features(Rngs(0))
    b, t, c = 2, 4, 3
    for hw in (64, 128):
        feats = vgg(torch.ones(b * t, hw, hw, c))
        assert feats["relu1_1"].shape == (b * t, hw, hw, 64) and feats["relu1_2"].shape == (b * t, hw, hw, 64)
        assert feats["relu2_1"].shape == (b * t, hw // 2, hw // 2, 128)
        assert all((v >= 0).all() for v in feats.values())
    x = torch.ones(b, t, 64, 64, c, requires_grad=True)
    target = torch.ones(b, t, 64, 64, c) * 0.5
    scalar = get_perceptual_loss_fn(vgg)(None, x, target)
    per_sample = get_adversarial_perceptual_loss_fn(vgg)(None, x, target)
    assert scalar.shape == () and torch.isfinite(scalar) and scalar > 0
    assert per_sample.shape == (b,) and torch.allclose(per_sample.mean(), scalar, rtol=1e-5)
    scalar.backward()
    assert x.grad.shape == x.shape and torch.isfinite(x.grad).all() and x.grad.abs().max() > 0
    # corner pixel of relu1_1 for a constant image: only the 2x2 in-image taps contribute (zero padding is applied to
    # the NORMALISED image)
    xn = (torch.ones(3) - torch.tensor(IMAGENET_MEAN)) / torch.tensor(IMAGENET_STD)
    k = vgg.conv1_1_kernel
    want = torch.relu(torch.einsum("hwio,i->o", k[1:, 1:], xn) + vgg.conv1_1_bias)
    got = vgg(torch.ones(1, 8, 8, 3))["relu1_1"][0, 0, 0]
    assert torch.allclose(got, want, atol=1e-5)


def test_oracle_reproduces_rl_step_golden_fixture():
    """tests/golden/rl_step_cfg64_fp32.npz: the rows next to the path in one fixture -- rl_model forward, the RL loss
    with the VGG perceptual term, every gradient norm and one clip+Adam step, as written by make_golden.py."""
    import importlib.util
    import os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "rl_step_cfg64_fp32.npz"))
    now = mg.run_rl_oracle()
    assert sorted(now) == sorted(gold.files)
    for key in gold.files:
        if key == "rl_grad_names":
            assert list(now[key]) == list(gold[key])
        elif key == "rl_selection_mask":
            assert np.array_equal(now[key], gold[key])
        else:
            ref = np.asarray(gold[key], dtype=np.float64)
            got = np.asarray(now[key], dtype=np.float64)
            assert np.allclose(got, ref, rtol=5e-4, atol=5e-5 * max(1e-6, np.abs(ref).max())), key


def test_oracle_reproduces_production_depth_golden_fixture():
    """tests/golden/videovae_prod128_fp32.npz: BASELINE configs[0] (one 16x128x128 clip, fp32) at PRODUCTION depth
    (enc 9 / dec 12, mlp 1536, 8 heads x 64), 12 of 16 frames kept -- loss terms, latent / reconstruction slices and
    all 600+ gradient norms as written by make_golden.py::run_prod_oracle (thread count changes BLAS summation order
    over 21 layers: rtol 1e-3 on gradient norms)."""
    import importlib.util
    import os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "videovae_prod128_fp32.npz"))
    now = mg.run_prod_oracle()
    assert sorted(now) == sorted(gold.files)
    assert len(gold["grad_names"]) > 600
    for key in gold.files:
        if key == "grad_names":
            assert list(now[key]) == list(gold[key])
        elif key == "selection":
            assert np.array_equal(now[key], gold[key])
            assert np.array_equal(now[key].reshape(-1), np.asarray(mg.PROD_KEEP_FRAMES, dtype=np.float32))
        else:
            ref = np.asarray(gold[key], dtype=np.float64)
            got = np.asarray(now[key], dtype=np.float64)
            assert np.allclose(got, ref, rtol=1e-3, atol=1e-4 * max(1e-6, np.abs(ref).max())), key
