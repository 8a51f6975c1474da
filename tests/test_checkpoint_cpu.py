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


def _reference_shaped_checkpoint(model, mu_model, nu_model, count=41):
    """A tree shaped exactly like the reference's ``{"model": nnx.state(model), "optimizer": nnx.state(optimizer)}``
    (train/rl_nonadversarial.py:62-67): ``{"value": ...}`` leaves, int list indices, the RotaryEmbedding tables that
    train/layers.py:101-102 stores as nnx.Variables ([1, max_len, 1, head_dim]), RNG stream state, and the optax chain
    ``(EmptyState(), (ScaleByAdamState(count, mu, nu), ScaleByScheduleState(count)))`` under ``opt_state``."""
    tree = ck.to_flax_tree(model, wrap_value=True)
    for name, mod in model.named_modules():
        if name.endswith("ROPE"):
            node = tree
            for p_ in name.split("."):
                node = node.setdefault(int(p_) if p_.isdigit() else p_, {})
            node["cos_cached"] = {"value": mod.cos_cached.numpy()[None, :, None, :]}
            node["sin_cached"] = {"value": mod.sin_cached.numpy()[None, :, None, :]}
    tree["encoder"]["rngs"] = {"default": {"key": {"value": np.zeros(2, np.uint32)}, "count": {"value": np.zeros((), np.uint32)}}}
    opt = {"step": {"value": np.asarray(count, np.uint32)},
           "opt_state": {0: {},
                         1: {0: {"count": {"value": np.asarray(count, np.int32)},
                                 "mu": ck.to_flax_tree(mu_model, wrap_value=True),
                                 "nu": ck.to_flax_tree(nu_model, wrap_value=True)},
                             1: {"count": {"value": np.asarray(count, np.int32)}}}}}
    return {"model": tree, "optimizer": opt}


def test_reference_shaped_checkpoint_with_rope_tables_and_optimizer_state(tmp_path):
    """VERDICT r1 / ADVICE: a real reference tree carries ROPE.{cos,sin}_cached Variables and an optax chain state."""
    from video_vae_b200.ddp import FlatAdam, FlatParams
    src, mu_m, nu_m = _model(0), _model(3), _model(4)
    state = _reference_shaped_checkpoint(src, mu_m, nu_m, count=41)
    assert "cos_cached" in state["model"]["encoder"]["layers"][0]["TemporalAttention"]["ROPE"]
    dst = _model(1)
    flat = FlatParams(dst)
    adam = FlatAdam(flat)
    loaded = ck.load_flax_tree(dst, state["model"], strict=True)          # strict: RoPE tables / rngs are not "unexpected"
    assert len(loaded) == len(src.state_dict())
    assert all(torch.equal(x, y) for x, y in zip(src.state_dict().values(), dst.state_dict().values()))
    assert ck.rope_tables_match(state["model"], dst) == 8                 # 2 layers x 2 attentions x (cos, sin)
    bad = _reference_shaped_checkpoint(src, mu_m, nu_m)
    bad["model"]["encoder"]["layers"][0]["TemporalAttention"]["ROPE"]["cos_cached"]["value"] *= 1.01
    with pytest.raises(ValueError, match="RoPE"):
        ck.rope_tables_match(bad["model"], dst)
    assert ck.load_optimizer_state(adam, dst, state["optimizer"]) == 41 and adam.t == 41
    names = [n for n, _ in dst.named_parameters()]
    for probe in ("encoder.layers.0.SpatialAttention.qkv_projection.kernel", "fill_token"):
        i = names.index(probe)
        s, n = flat.offsets[i], flat.params[i].numel()
        assert torch.equal(adam.m[s:s + n], dict(mu_m.named_parameters())[probe].detach().reshape(-1))
        assert torch.equal(adam.v[s:s + n], dict(nu_m.named_parameters())[probe].detach().reshape(-1))
    # the flattened transport file of the whole {"model", "optimizer"} state
    ck.save_npz(str(tmp_path / "full.npz"), state)
    again = _model(2)
    flat2 = FlatParams(again)
    adam2 = FlatAdam(flat2)
    ck.load_checkpoint(again, str(tmp_path / "full.npz"), adam=adam2)
    assert all(torch.equal(x, y) for x, y in zip(src.state_dict().values(), again.state_dict().values()))
    assert adam2.t == 41 and torch.equal(adam2.m, adam.m) and torch.equal(adam2.v, adam.v)


def test_two_flat_params_keep_their_own_bf16_shadows():
    """a second FlatParams.enable_bf16_shadow() must not orphan the first model's shadow views."""
    from video_vae_b200 import functional as F_
    from video_vae_b200.ddp import FlatParams
    a, b = _model(0), _model(1)
    fa, fb = FlatParams(a), FlatParams(b)
    fa.shadow = torch.empty(fa.total, dtype=torch.bfloat16)
    fb.shadow = torch.empty(fb.total, dtype=torch.bfloat16)
    import weakref
    for f in (fa, fb):      # what enable_bf16_shadow does, minus the device cast kernel
        for p, o in zip(f.params, f.offsets):
            f._shadow_views[id(p)] = (weakref.ref(p), f.shadow[o:o + p.numel()].view(p.shape))
        F_._flat_shadow_views.update(f._shadow_views)
    pa, pb = a.fill_token, b.fill_token
    assert F_.shadow(pa, torch.bfloat16).data_ptr() == fa._shadow_views[id(pa)][1].data_ptr()
    assert F_.shadow(pb, torch.bfloat16).data_ptr() == fb._shadow_views[id(pb)][1].data_ptr()


REFERENCE = "/root/reference"
WRITE_REFERENCE_CHECKPOINT = r"""
import sys
import numpy as np
import torch
shim, ref_train, out = sys.argv[1:4]
sys.path.insert(0, shim)
sys.path.insert(0, ref_train)
import jax.numpy as jnp
import optax
from flax import nnx
from model import VideoVAE                                         # the reference's train/model.py
from video_vae_b200 import checkpoint as ck
cfg = (64, 64, 3, 16, 1, 1, 128, 2, 128, 16, 8, 4)
model = VideoVAE(*cfg, rngs=nnx.Rngs(2), dtype=jnp.float32, param_dtype=jnp.float32)
optimizer = nnx.Optimizer(model, optax.chain(optax.clip_by_global_norm(1.0), optax.adam(learning_rate=1e-3)))   # rl_nonadversarial.py:248-253
g = torch.Generator().manual_seed(0)
for _ in range(3):
    params = nnx.state(model, nnx.Param)
    grads = nnx.State({p: 0.05 * torch.randn(v.shape, generator=g) for p, v in params.flat.items()})
    optimizer.update(grads)
# save_checkpoint of train/rl_nonadversarial.py:62-67, with .npz as the transport instead of orbax
state = {"model": nnx.state(model).to_pure_dict(), "optimizer": nnx.state(optimizer).to_pure_dict()}
print(len(ck.save_npz(out, state)))
"""


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE + "/train"), reason="/root/reference is only present in the build container")
def test_checkpoint_written_by_the_reference_model_code_loads(tmp_path):
    """{"model": nnx.state(model), "optimizer": nnx.state(optimizer)} as train/rl_nonadversarial.py:62-67 saves it, produced by
    the reference's OWN train/model.py (run on oracle/jaxshim) after three optimizer updates: the key set is the reference's
    -- RotaryEmbedding's cos_cached / sin_cached Variables (train/layers.py:101-102) included, the model's parameters a
    second time under optimizer.model -- and the product loads it strictly, parameters, Adam moments and step count."""
    import os
    import subprocess
    import sys
    from video_vae_b200.ddp import FlatAdam, FlatParams
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = str(tmp_path / "reference_checkpoint.npz")
    r = subprocess.run([sys.executable, "-c", WRITE_REFERENCE_CHECKPOINT, os.path.join(root, "oracle", "jaxshim"),
                        os.path.join(REFERENCE, "train"), path], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONPATH=root), cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = ck.load_npz(path)
    assert any(k.endswith("ROPE.cos_cached") for k in raw) and any(k.startswith("optimizer.opt_state.1.0.mu.") for k in raw)
    assert any(k.startswith("optimizer.model.") for k in raw)
    dst = _model(1)
    flat = FlatParams(dst)
    adam = FlatAdam(flat)
    loaded = ck.load_checkpoint(dst, path, strict=True, adam=adam)
    names = [n for n, _ in dst.named_parameters()]
    assert sorted(loaded) == sorted(names) and adam.t == 3
    assert ck.rope_tables_match(ck._unflatten({k[len("model."):]: v for k, v in raw.items() if k.startswith("model.")}), dst) == 8
    for i, n in enumerate(names):
        s, cnt = flat.offsets[i], flat.params[i].numel()
        assert torch.equal(flat.params[i].detach().reshape(-1), torch.from_numpy(raw["model." + n]).reshape(-1)), n
        assert torch.equal(adam.m[s:s + cnt], torch.from_numpy(raw["optimizer.opt_state.1.0.mu." + n]).reshape(-1)), n
        assert torch.equal(adam.v[s:s + cnt], torch.from_numpy(raw["optimizer.opt_state.1.0.nu." + n]).reshape(-1)), n
    assert float(adam.v.abs().max()) > 0.0
