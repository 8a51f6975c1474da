"""Weight import/export against a Flax-shaped parameter tree (train/model_loader.py:35-42). CPU only."""
import numpy as np
import pytest
import torch

import video_vae_b200 as V
from video_vae_b200 import checkpoint as ck

CFG = (64, 64, 3, 16, 1, 1, 128, 2, 128, 16, 8, 4)


def _model(seed):
    return V.VideoVAE(*CFG, V.Rngs(seed), dtype=torch.float32, device="cpu")


def test_flax_tree_round_trip_nested_value_wrappers(tmp_path):
    a, b = _model(0), _model(1)
    tree = ck.to_flax_tree(a, wrap_value=True)
    # shaped like nnx.state(model): attribute names, int list indices, {"value": array} leaves, Flax layouts
    assert tree["encoder"]["layers"][0]["TemporalAttention"]["qkv_projection"]["kernel"]["value"].shape == (768, 384)
    assert tree["decoder"]["unet"]["final_conv"]["kernel"]["value"].ndim == 5
    tree["encoder"]["rngs"] = {"default": {"key": {"value": np.zeros(2, np.uint32)}, "count": {"value": np.zeros(())}}}
    loaded = ck.load_flax_tree(b, tree)
    assert len(loaded) == len(a.state_dict())
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    # transport file
    ck.save_checkpoint(a, str(tmp_path / "w.npz"))
    c = _model(2)
    ck.load_checkpoint(c, str(tmp_path / "w.npz"))
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), c.state_dict().values()))


def test_mismatches_are_reported():
    a, b = _model(0), _model(1)
    tree = ck.to_flax_tree(a)
    del tree["fill_token"]
    with pytest.raises(KeyError, match="fill_token"):
        ck.load_flax_tree(b, tree)
    ck.load_flax_tree(b, tree, strict=False)
    tree = ck.to_flax_tree(a)
    tree["decoder"]["unet"]["final_conv"]["kernel"] = tree["decoder"]["unet"]["final_conv"]["kernel"][..., :2]
    with pytest.raises(ValueError, match="final_conv.kernel"):
        ck.load_flax_tree(b, tree)
    with pytest.raises(ImportError, match="orbax"):
        ck.load_checkpoint(b, "/tmp")


def test_oracle_and_product_share_the_tree():
    """The oracle is built on the same names/layouts, so a product export loads into it and vice versa."""
    from oracle import Rngs as ORngs
    from oracle.model import VideoVAE as OV
    o = OV(*CFG, ORngs(5))
    m = _model(0)
    ck.load_flax_tree(m, ck.to_flax_tree(o))
    assert all(torch.equal(x, y) for x, y in zip(o.state_dict().values(), m.state_dict().values()))


def test_adam_moments_import():
    from video_vae_b200.ddp import FlatAdam, FlatParams
    m = _model(0)
    flat = FlatParams(m)
    adam = FlatAdam(flat)
    mu = ck.to_flax_tree(_model(3))
    nu = ck.to_flax_tree(_model(4))
    ck.load_adam_moments(adam, m, mu, nu, count=17)
    assert adam.t == 17
    i = [n for n, _ in m.named_parameters()].index("decoder.unet.final_conv.bias")
    s = flat.offsets[i]
    want = torch.from_numpy(ck.flatten_tree(mu)["decoder.unet.final_conv.bias"])
    assert torch.equal(adam.m[s:s + want.numel()], want)


def test_rl_host_helpers_need_no_host_reads():
    """rl_model.repeat2 == einops 'b ... -> (b 2) ...' (== repeat_interleave), and the capture-safe form of the
    trajectory-probability product used by rl_losses (1 + sum_t(p_t - 1) with p_t = raw/stop_grad(raw)) has the value
    and the gradient of prod_t(p_t) (train/rl_nonadversarial.py:164-171)."""
    from video_vae_b200.rl_model import repeat2
    a = torch.arange(24.0).reshape(3, 4, 2)
    assert torch.equal(repeat2(a), a.repeat_interleave(2, dim=0))
    raw = (torch.rand(2, 2, 5, generator=torch.Generator().manual_seed(0)) * 0.9 + 0.05).requires_grad_()
    mask = torch.tensor([[True, True, True, False, False]]).expand(2, 2, 5)
    one = torch.ones_like(raw)
    w = torch.tensor([[0.7, -0.7], [-1.0, 1.0]])[:, :, None]
    ratio = torch.where(mask, raw / raw.detach(), one)
    (ratio.prod(dim=2, keepdim=True) * w).sum().backward()
    g_prod = raw.grad.clone()
    raw.grad = None
    ratio = torch.where(mask, raw / raw.detach(), one)
    alt = 1.0 + (ratio - 1.0).sum(dim=2, keepdim=True)
    assert torch.equal(alt.detach(), torch.ones_like(alt))
    (alt * w).sum().backward()
    assert torch.allclose(raw.grad, g_prod, rtol=1e-6, atol=0) and float(raw.grad[..., 3:].abs().max()) == 0.0
