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
