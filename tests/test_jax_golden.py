"""Consumes fixtures written by tests/golden/make_golden_jax.py: outputs of the REFERENCE'S OWN CODE.

Two kinds of file, same schema:
  * tests/golden/refshim_videovae_*.npz (committed): train/model.py, layers.py, unet.py and the loss_fn of
    train/legacy/training_loop_adversarial.py executed unmodified on oracle/jaxshim (jax / flax.nnx look-alikes on CPU
    torch; see oracle/jaxshim/README.md) -- pins every line the reference wrote; the third-party primitives
    (nnx.Linear / LayerNorm / GroupNorm / Conv / ConvTranspose, jax.nn.dot_product_attention) are restated by the shim;
  * tests/golden/jax_videovae_*.npz (none yet: needs a box with real jax / flax) -- pins those too.

For every such file:
  * CPU: the reference's weights and recorded draws go through the ORACLE and must reproduce the reference's outputs and
    gradients (fp32: rel 1e-4 / 1e-3) -- this is what turns "parity unpinned" into pinned;
  * GPU: the same through the CUDA path (fp32 fixture -> fp32 kernels at 1e-4 / 1e-3; bf16 fixture -> bf16 at 2e-2 on
    loss / mean / logvar, north_star's bar).
One more CPU test checks the consumer itself against a stand-in fixture of the same schema written from the oracle.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
import sys
sys.path.insert(0, GOLDEN)
import weight_recipe  # noqa: E402

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "jax_videovae_*.npz")) + glob.glob(os.path.join(GOLDEN, "refshim_videovae_*.npz")))
RL_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "jax_rlvae_*.npz")) + glob.glob(os.path.join(GOLDEN, "refshim_rlvae_*.npz")))
DIST_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "jax_rldistvae_*.npz")) + glob.glob(os.path.join(GOLDEN, "refshim_rldistvae_*.npz")))
TRAIN_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "jax_*train_*.npz")) + glob.glob(os.path.join(GOLDEN, "refshim_*train_*.npz")))
NO_FIXTURE = "parity unpinned: no tests/golden/{jax,refshim}_videovae_*.npz (tests/golden/make_golden_jax.py [--shim])"


def rel_l2(a, ref):
    a, ref = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(ref)).double()
    return ((a - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def rel_err(a, ref):
    a, ref = torch.as_tensor(np.asarray(a)).float(), torch.as_tensor(np.asarray(ref)).float()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def load_fixture(path):
    z = np.load(path)
    fx = {"cfg": tuple(int(v) for v in z["cfg"]), "dtype": str(z["dtype"]), "hparams": json.loads(str(z["hparams"])),
          "mask": torch.from_numpy(z["mask"]).bool(),
          "gumbel_u": torch.from_numpy(z["gumbel_u"]),
          "noise": torch.from_numpy(z["noise"] if "noise" in z.files else weight_recipe.normal(z["noise_shape"])),
          "latent_stride": int(z["latent_stride"]) if "latent_stride" in z.files else 1,
          "out": {k[4:]: z[k] for k in z.files if k.startswith("out/")}, "recon_stride": 1, "compact": "recipe" in z.files,
          "model": str(z["model"]) if "model" in z.files else "vae"}
    if fx["compact"]:
        # --compact files: weights and clip by name from tests/golden/weight_recipe.py, gradients as norm + max + probe
        assert str(z["recipe"]) == weight_recipe.RECIPE_ID
        fx["video"] = torch.from_numpy(weight_recipe.clip(tuple(int(v) for v in z["video_shape"])))
        fx["params"] = {k[7:]: weight_recipe.param(k[7:], z[k]) for k in z.files if k.startswith("pshape/")}
        fx["grads"] = {k[6:]: {"norm": float(z[k]), "max": float(z["gmax/" + k[6:]]), "probe": z["gprobe/" + k[6:]]}
                       for k in z.files if k.startswith("gnorm/")}
        fx["recon_stride"] = int(z["recon_stride"])
    else:
        fx["video"] = torch.from_numpy(z["video"])
        fx["params"] = {k[6:]: z[k] for k in z.files if k.startswith("param/")}
        fx["grads"] = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
    return fx


def cube_perceptual(vgg_params, reconstruction, target):
    """The closed-form stand-in make_golden_jax.py passes as perceptual_loss_fn (rl_nonadversarial.py:125), in torch."""
    return ((reconstruction.float() - target.float()).abs() ** 3).mean(dim=(1, 2, 3, 4))


def run_impl_rl(fx, impl, dtype):
    """train/rl_model.py + the RL loss of train/rl_nonadversarial.py:100-186 with the fixture's weights and draws."""
    from video_vae_b200 import checkpoint as ck
    cfg = fx["cfg"]
    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    u = fx["gumbel_u"]                                          # the uniform behind jax.random.bernoulli, ((b 2), t, 1, 1)
    if impl == "oracle":
        from oracle import Rngs
        from oracle.losses import expand_mask
        from oracle.rl_losses import loss_fn
        from oracle.rl_model import VideoVAE
        m = VideoVAE(*cfg, Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        loss, aux = loss_fn(m, fx["video"], expand_mask(fx["mask"], hw), fx["mask"], Rngs(0), fx["hparams"],
                            cube_perceptual, None, noise=fx["noise"], bernoulli_u=u)
    else:
        import video_vae_b200 as V
        from video_vae_b200.rl_losses import loss_fn
        from video_vae_b200.rl_model import VideoVAE
        m = VideoVAE(*cfg, V.Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        loss, aux = loss_fn(m, fx["video"].cuda(), fx["mask"][:, None, None, :].cuda(), fx["mask"].cuda(), V.Rngs(0),
                            fx["hparams"], cube_perceptual, None, noise=fx["noise"].cuda(), bernoulli_u=u.cuda())
    loss.backward()
    return loss, aux, m


def check_rl_against_fixture(fx, loss, aux, m, tol, grad_tol):
    out = fx["out"]
    f = lambda v: v.detach().float().cpu().numpy()             # noqa: E731
    assert np.array_equal(f(aux["selection_mask"]).reshape(-1), out["selection_mask"].reshape(-1))
    report = {"selection": rel_err(torch.from_numpy(f(aux["selection"]).reshape(-1)), out["selection"].reshape(-1))}
    st = fx.get("recon_stride", 1)
    report["reconstruction"] = rel_err(aux["reconstruction"].detach().float().cpu()[:, :, ::st, ::st, :], out["reconstruction"])
    for k in ("selection", "reconstruction"):
        assert report[k] < tol, (k, report[k])
    for k in ("MSE", "perceptual_loss", "selection_loss", "kl_loss", "kept_frame_density", "mean_trajectory_prob",
              "per_sample_MAE"):
        report[k] = abs(float(aux[k]) - float(out[k])) / max(abs(float(out[k])), 1e-6)
        assert report[k] <= tol, (k, report[k])
    # rl_loss is mean(probs * disadvantages) with probs == 1 in value and disadvantages centred per pair: zero up to
    # rounding; it matters through its gradient (the selection layers' gradients below)
    assert abs(float(aux["rl_loss"]) - float(out["rl_loss"])) <= 1e-5
    report["loss"] = abs(float(loss) - float(out["loss"])) / abs(float(out["loss"]))
    assert report["loss"] <= tol
    named = dict(m.named_parameters())
    checked, worst = 0, (0.0, None)
    for name, ref in fx["grads"].items():
        assert isinstance(ref, dict), "rl fixtures are written with --compact"
        if ref["max"] == 0.0:
            continue
        got = named[name].grad.detach().float().cpu().numpy()
        gn = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
        e_norm = abs(gn - ref["norm"]) / ref["norm"]
        d_probe = np.abs(weight_recipe.grad_probe(got) - ref["probe"]) / ref["max"]
        assert e_norm < grad_tol, (name, e_norm)
        assert float((d_probe < 5 * grad_tol).mean()) >= 0.94 and float(d_probe.max()) < 2e-2, (name, float(d_probe.max()))
        worst = max(worst, (max(e_norm, float(d_probe.max())), name))
        checked += 1
    assert checked >= 50
    assert any("selection_layer" in n for n, r in fx["grads"].items() if r["max"] > 0.0)      # the RL term's gradient is there
    report.update(grads_checked=checked, worst_grad=worst)
    print("rl fixture parity:", report)
    return report


def dist_loss(video, reconstruction, selection, variance, mean):
    """The closed-form loss make_golden_jax.py --model rl_dist differentiates (the trainer's own loss is a closure inside
    claude_distributed/distributed_train.py's main(); this fixture pins the MODEL copy of claude_distributed/)."""
    v2 = video.repeat_interleave(2, dim=0).to(reconstruction.dtype)
    return (((reconstruction - v2) ** 2).mean() + 0.01 * variance.float().mean() + 0.01 * (mean.float() ** 2).mean()
            + 0.1 * selection.float().mean())


def run_impl_dist(fx, impl, dtype):
    """claude_distributed/rl_model.py (returns the variance) with the fixture's weights and draws."""
    from video_vae_b200 import checkpoint as ck
    cfg = fx["cfg"]
    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    u = fx["gumbel_u"]
    if impl == "oracle":
        from oracle import Rngs
        from oracle.distributed_rl_model import VideoVAE
        from oracle.losses import expand_mask
        m = VideoVAE(*cfg, Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        video = fx["video"]
        outs = m(video, expand_mask(fx["mask"], hw), Rngs(0), train=True, noise=fx["noise"], bernoulli_u=u)
    else:
        import video_vae_b200 as V
        from video_vae_b200.distributed_rl_model import VideoVAE
        m = VideoVAE(*cfg, V.Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        video = fx["video"].cuda()
        outs = m(video, fx["mask"][:, None, None, :].cuda(), V.Rngs(0), train=True, noise=fx["noise"].cuda(), bernoulli_u=u.cuda())
    reconstruction, compressed, selection, selection_mask, variance, mean = outs
    loss = dist_loss(video, reconstruction, selection, variance, mean)
    loss.backward()
    return loss, dict(reconstruction=reconstruction, compressed=compressed, selection=selection,
                      selection_mask=selection_mask, variance=variance, mean=mean), m


def check_dist_against_fixture(fx, loss, aux, m, tol, grad_tol):
    out = fx["out"]
    f = lambda v: v.detach().float().cpu()                     # noqa: E731
    assert np.array_equal(f(aux["selection_mask"]).numpy().reshape(-1), out["selection_mask"].reshape(-1))
    st, lat = fx.get("recon_stride", 1), fx.get("latent_stride", 1)
    report = {"selection": rel_err(f(aux["selection"]).reshape(-1), out["selection"].reshape(-1)),
              "reconstruction": rel_err(f(aux["reconstruction"])[:, :, ::st, ::st, :], out["reconstruction"]),
              # the fixture stores the 5th output of VideoVAE.__call__ under the key "logvar": here it IS the variance
              "variance": rel_err(f(aux["variance"])[:, :, ::lat, :], out["logvar"]),
              "mean": rel_err(f(aux["mean"])[:, :, ::lat, :], out["mean"]),
              "compressed": rel_err(f(aux["compressed"])[:, :, ::lat, :], out["compressed"]),
              "loss": abs(float(loss) - float(out["loss"])) / abs(float(out["loss"]))}
    assert float(out["logvar"].min()) > 0.0                   # a variance, not its logarithm
    for k, v in report.items():
        assert v < tol, (k, v)
    named = dict(m.named_parameters())
    checked, worst = 0, (0.0, None)
    for name, ref in fx["grads"].items():
        if ref["max"] == 0.0:
            continue
        got = named[name].grad.detach().float().cpu().numpy()
        gn = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
        e_norm = abs(gn - ref["norm"]) / ref["norm"]
        d_probe = np.abs(weight_recipe.grad_probe(got) - ref["probe"]) / ref["max"]
        assert e_norm < grad_tol, (name, e_norm)
        assert float((d_probe < 5 * grad_tol).mean()) >= 0.94 and float(d_probe.max()) < 2e-2, (name, float(d_probe.max()))
        worst = max(worst, (max(e_norm, float(d_probe.max())), name))
        checked += 1
    assert checked >= 50
    report.update(grads_checked=checked, worst_grad=worst)
    print("rl_dist fixture parity:", report)
    return report


def run_impl(fx, impl, dtype):
    """Feed the fixture's weights, clip, mask and draws through `impl` ("oracle" on CPU, "cuda"); returns loss, aux, model."""
    from video_vae_b200 import checkpoint as ck
    cfg = fx["cfg"]
    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    b, t = fx["mask"].shape
    u = fx["gumbel_u"].reshape(b, t, 1)
    if impl == "oracle":
        from oracle import Rngs
        from oracle.losses import expand_mask, loss_fn
        from oracle.model import VideoVAE
        m = VideoVAE(*cfg, Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        loss, aux = loss_fn(m, fx["video"], expand_mask(fx["mask"], hw), fx["mask"], Rngs(0), fx["hparams"],
                            noise=fx["noise"], gumbel_u=u)
    else:
        import video_vae_b200 as V
        m = V.VideoVAE(*cfg, V.Rngs(0), dtype=dtype)
        ck.load_flax_tree(m, fx["params"], strict=True)
        loss, aux = V.loss_fn(m, fx["video"].cuda(), fx["mask"][:, None, None, :].cuda(), fx["mask"].cuda(), V.Rngs(0),
                              fx["hparams"], noise=fx["noise"].cuda(), gumbel_u=u.cuda())
    loss.backward()
    return loss, aux, m


def check_against_fixture(fx, loss, aux, m, tol, grad_tol, lowp=False):
    out = fx["out"]
    assert np.array_equal(aux["selection"].detach().float().cpu().numpy().reshape(-1), out["selection"].reshape(-1))
    keys = ("mean", "logvar") if lowp else ("mean", "logvar", "reconstruction", "compressed")
    st = fx.get("recon_stride", 1)
    report = {}
    for k in keys:
        got = aux[k].detach().float().cpu()
        if k == "reconstruction":
            got = got[:, :, ::st, ::st, :]
        else:
            got = got[:, :, ::fx.get("latent_stride", 1), :]
        report[k] = rel_err(got, out[k])
        assert report[k] < tol, (k, report[k])
    for k in (("MSE",) if lowp else ("MSE", "selection_loss", "kl_loss")):
        assert abs(float(aux[k]) - float(out[k])) <= tol * max(abs(float(out[k])), 1e-6), k
    assert abs(float(loss) - float(out["loss"])) <= tol * abs(float(out["loss"]))
    if grad_tol is None:
        report["loss"] = abs(float(loss) - float(out["loss"])) / abs(float(out["loss"]))
        print("fixture parity (low precision):", report)
        return report
    named = dict(m.named_parameters())
    checked = 0
    worst = (0.0, None)
    for name, ref in fx["grads"].items():
        if isinstance(ref, dict):                            # compact fixture: norm, max and a strided probe
            if ref["max"] == 0.0:
                continue
            got = named[name].grad.detach().float().cpu().numpy()
            gn = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
            e_norm = abs(gn - ref["norm"]) / ref["norm"]
            d_probe = np.abs(weight_recipe.grad_probe(got) - ref["probe"]) / ref["max"]
            e_probe = float(d_probe.max())
            assert e_norm < grad_tol, (name, e_norm)
            # The U-Net has max-pools: when the two candidates of one 2x2 window differ by less than the fp32 noise of
            # the forward pass (refshim "small": 1 window of 98304 at encoders.1, values 1e-5 apart), two correct
            # implementations route that window's gradient to different pixels and the few kernel-gradient elements
            # fed by that pixel move by a few 1e-3 of the tensor's maximum while everything else agrees to 1e-6.  So:
            # the norm to grad_tol, >= 94 % of the probe to 5 * grad_tol, every probe element to 2e-2.
            assert float((d_probe < 5 * grad_tol).mean()) >= 0.94, (name, e_probe, float((d_probe < 5 * grad_tol).mean()))
            assert e_probe < 2e-2, (name, e_probe)
            worst = max(worst, (max(e_norm, e_probe), name))
            checked += 1
            continue
        if float(np.abs(ref).max()) == 0.0:
            continue
        got = named[name].grad.detach().float().cpu()
        # rel-L2 per tensor at grad_tol; max-norm gets 5x: conv kernels in front of a GroupNorm have scale-invariant
        # losses, their gradients are residuals of cancelling sums (see tests/test_parity_prod_gpu.py)
        assert rel_l2(got, ref) < grad_tol, name
        assert rel_err(got, ref) < 5 * grad_tol, name
        checked += 1
    assert checked >= 50
    report.update(loss=abs(float(loss) - float(out["loss"])) / abs(float(out["loss"])), grads_checked=checked,
                  worst_grad=worst)
    print("fixture parity:", report)
    return report


# ------------------------------------------------------------------------------------------------ real fixtures
@pytest.mark.skipif(not FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_oracle_reproduces_reference_outputs(path):
    fx = load_fixture(path)
    if fx["dtype"] != "float32":
        pytest.skip("the oracle is an fp32 restatement; bf16 fixtures are consumed by the GPU test")
    loss, aux, m = run_impl(fx, "oracle", torch.float32)
    check_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)


BF16_PAIRS = [(p, p.replace("_bfloat16.npz", "_float32.npz")) for p in FIXTURES
              if p.endswith("_bfloat16.npz") and os.path.exists(p.replace("_bfloat16.npz", "_float32.npz"))]


@pytest.mark.skipif(not BF16_PAIRS, reason="no bf16 / fp32 fixture pair")
@pytest.mark.parametrize("lowp,full", BF16_PAIRS or [(None, None)])
def test_reference_bf16_against_its_own_fp32(lowp, full):
    """The reference's code run with dtype=bfloat16 (flax promote_dtype semantics: bf16 activations and matmul operands,
    fp32 parameters, fp32 LayerNorm / GroupNorm statistics and attention logits) against the same code in fp32, same
    weights and draws: the depth-wise growth the reference documents (train/llm_tests.py:491-502) stays inside
    north_star's 2e-2 at production depth -- the yardstick the CUDA bf16 path is held to."""
    a, b = np.load(lowp), np.load(full)
    assert np.array_equal(a["gumbel_u"], b["gumbel_u"]) and np.array_equal(a["noise_shape"], b["noise_shape"])
    assert np.array_equal(a["out/selection"], b["out/selection"])
    rep = {k: rel_err(torch.from_numpy(a["out/" + k]), b["out/" + k]) for k in ("mean", "logvar", "reconstruction")}
    rep["loss"] = abs(float(a["out/loss"]) - float(b["out/loss"])) / abs(float(b["out/loss"]))
    print("reference bf16 vs its fp32:", os.path.basename(lowp), rep)
    assert rep["mean"] < 2e-2 and rep["logvar"] < 2e-2 and rep["loss"] < 2e-2 and rep["reconstruction"] < 3e-2


@pytest.mark.gpu
@pytest.mark.skipif(not FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_cuda_path_reproduces_reference_outputs(path):
    fx = load_fixture(path)
    if fx["dtype"] == "float32":
        loss, aux, m = run_impl(fx, "cuda", torch.float32)
        check_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)
    else:
        # bf16 fixture = the reference's code run with dtype=bfloat16.  Two bars: against the reference's fp32 outputs for
        # the same weights and draws (the sibling *_float32.npz) north_star's 2e-2 on loss / mean / logvar; against the
        # reference's bf16 outputs 3e-2 (two independent bf16 roundings of a 21-layer network: sqrt(2) x the above)
        loss, aux, m = run_impl(fx, "cuda", torch.bfloat16)
        check_against_fixture(fx, loss, aux, m, 3e-2, None, lowp=True)
        sibling = path.replace("_bfloat16.npz", "_float32.npz")
        if os.path.exists(sibling):
            check_against_fixture(load_fixture(sibling), loss, aux, m, 2e-2, None, lowp=True)


@pytest.mark.skipif(not RL_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", RL_FIXTURES or [None])
def test_oracle_reproduces_reference_rl_outputs(path):
    """train/rl_model.py and the loss_fn of train/rl_nonadversarial.py (the reference's files, executed by
    tests/golden/make_golden_jax.py --model rl) against their oracle restatements (oracle/rl_model.py, rl_losses.py)."""
    fx = load_fixture(path)
    assert fx["model"] == "rl" and fx["dtype"] == "float32"
    loss, aux, m = run_impl_rl(fx, "oracle", torch.float32)
    check_rl_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)


@pytest.mark.gpu
@pytest.mark.skipif(not RL_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", RL_FIXTURES or [None])
def test_cuda_path_reproduces_reference_rl_outputs(path):
    fx = load_fixture(path)
    loss, aux, m = run_impl_rl(fx, "cuda", torch.float32)
    check_rl_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)


@pytest.mark.skipif(not DIST_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", DIST_FIXTURES or [None])
def test_oracle_reproduces_reference_distributed_model_outputs(path):
    """claude_distributed/{rl_model,layers,unet}.py (the data-parallel trainer's copies: variance instead of log-variance,
    the mask expanded inside FactoredAttention) executed by make_golden_jax.py --model rl_dist, against
    oracle/distributed_rl_model.py."""
    fx = load_fixture(path)
    assert fx["model"] == "rl_dist" and fx["dtype"] == "float32"
    loss, aux, m = run_impl_dist(fx, "oracle", torch.float32)
    check_dist_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)


@pytest.mark.gpu
@pytest.mark.skipif(not DIST_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", DIST_FIXTURES or [None])
def test_cuda_path_reproduces_reference_distributed_model_outputs(path):
    fx = load_fixture(path)
    loss, aux, m = run_impl_dist(fx, "cuda", torch.float32)
    check_dist_against_fixture(fx, loss, aux, m, 1e-4, 1e-3)


# ------------------------------------------------------------------------------------------------ training trajectory
def load_train_fixture(path):
    z = np.load(path)
    assert str(z["model"]) in ("rl_train", "vae_train") and str(z["recipe"]) == weight_recipe.RECIPE_ID
    names = [k[len("step0/pnorm/"):] for k in z.files if k.startswith("step0/pnorm/")]
    return z, {"cfg": tuple(int(v) for v in z["cfg"]), "hparams": json.loads(str(z["hparams"])), "rl": str(z["model"]) == "rl_train",
               "schedule": json.loads(str(z["schedule"])), "steps": int(z["steps"]), "names": names,
               "video": torch.from_numpy(weight_recipe.clip(tuple(int(v) for v in z["video_shape"]))),
               "mask": torch.from_numpy(z["mask"]).bool()}


def check_train_step(z, step, loss, aux, named, tol_loss, tol_param):
    """Loss terms of update `step` and the norm of every parameter AFTER it, against the fixture."""
    rep = {"loss": abs(float(loss) - float(z[f"step{step}/loss"])) / abs(float(z[f"step{step}/loss"]))}
    for k in ("MSE", "perceptual_loss", "selection_loss", "kl_loss", "kept_frame_density", "per_sample_MAE"):
        if f"step{step}/{k}" in z.files:              # the plain loss path has no perceptual / per-sample terms
            ref = float(z[f"step{step}/{k}"])
            rep[k] = abs(float(aux[k]) - ref) / max(abs(ref), 1e-6)
    for k, v in rep.items():
        assert v < tol_loss, (step, k, v)
    worst = 0.0
    for n, p in named.items():
        ref = float(z[f"step{step}/pnorm/{n}"])
        worst = max(worst, abs(float(p.detach().double().norm()) - ref) / max(ref, 1e-12))
    assert worst < tol_param, (step, worst)
    rep["worst_param_norm"] = worst
    return rep


def check_train_final(z, fx, named, moments, count, tol):
    """Final parameters (probe) and the optimizer state the reference checkpoints (nnx.state(optimizer)): count, mu, nu."""
    from video_vae_b200 import checkpoint as ck
    worst_p = 0.0
    for n, p in named.items():
        ref = z["final/pprobe/" + n]
        got = weight_recipe.grad_probe(p.detach().float().cpu().numpy())
        worst_p = max(worst_p, float(np.abs(got - ref).max() / max(float(np.abs(ref).max()), 1e-12)))
    assert worst_p < tol, worst_p
    # the key layout of the optimizer state, parsed by the product's own importer (opt_state -> 1 -> 0 -> {count, mu, nu})
    tree = {k[len("opt_probe/"):]: z[k] for k in z.files if k.startswith("opt_probe/")}
    tree.update({k[len("opt/"):]: z[k] for k in z.files if k.startswith("opt/")})
    mu, nu, found = ck._find_adam_state(tree)
    assert found == fx["steps"] == count and set(mu) == set(nu) == set(named)
    worst_m = 0.0
    for n in named:
        m, v = moments(n)
        for got, ref_probe, key in ((m, mu[n], "mu"), (v, nu[n], "nu")):
            scale = max(float(np.abs(ref_probe).max()), 1e-30)
            worst_m = max(worst_m, float(np.abs(weight_recipe.grad_probe(got) - ref_probe).max() / scale))
    assert worst_m < 20 * tol, worst_m          # second moments are squares of gradients: twice their relative error, per step
    return {"final_param_probe": worst_p, "moments_probe": worst_m, "count": found}


def check_eval_step(z, loss, aux, tol):
    """eval_step of training_loop_adversarial.py:139-148 (train=False) on the weights the six updates left."""
    rep = {"loss": abs(float(loss) - float(z["eval/loss"])) / abs(float(z["eval/loss"]))}
    for k in ("MSE", "selection_loss", "kl_loss", "kept_frame_density"):
        rep[k] = abs(float(aux[k]) - float(z["eval/" + k])) / max(abs(float(z["eval/" + k])), 1e-6)
    rep["reconstruction"] = rel_err(aux["reconstruction"].detach().float().cpu()[:, :, ::3, ::3, :], z["eval/reconstruction"])
    for k, v in rep.items():
        assert v < tol, (k, v)
    return rep


@pytest.mark.skipif(not TRAIN_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", TRAIN_FIXTURES or [None])
def test_oracle_reproduces_reference_training_trajectory(path):
    """train_step of train/rl_nonadversarial.py:188-198 + nnx.Optimizer(model, optax.chain(clip_by_global_norm(1.0),
    adam(warmup_cosine_decay_schedule))) (:241-253) over six updates, written by tests/golden/make_golden_train.py, against
    oracle/rl_losses.py + oracle/optim.py: lr = 0 on the first update, the clip active on five of six, a selection-loss
    spike in between -- every loss term and every parameter after every update."""
    from video_vae_b200 import checkpoint as ck
    from oracle import Rngs
    from oracle.losses import expand_mask
    from oracle.optim import ClipAdam, warmup_cosine_decay_schedule
    z, fx = load_train_fixture(path)
    if fx["rl"]:
        from oracle.rl_losses import loss_fn as rl_loss_fn
        from oracle.rl_model import VideoVAE

        def loss_fn(m, video, mask, original_mask, noise, u):
            return rl_loss_fn(m, video, mask, original_mask, Rngs(0), fx["hparams"], cube_perceptual, None, noise=noise, bernoulli_u=u)
    else:                                         # train/model.py + training_loop_adversarial.py:90-136: the step bench.py times
        from oracle.losses import loss_fn as vae_loss_fn
        from oracle.model import VideoVAE

        def loss_fn(m, video, mask, original_mask, noise, u):
            return vae_loss_fn(m, video, mask, original_mask, Rngs(0), fx["hparams"], noise=noise, gumbel_u=u)
    cfg = fx["cfg"]
    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    m = VideoVAE(*cfg, Rngs(0), dtype=torch.float32)
    shapes = {n: p.shape for n, p in m.named_parameters()}
    ck.load_flax_tree(m, {n: weight_recipe.param(n, shapes[n]) for n in fx["names"]}, strict=True)
    named = dict(m.named_parameters())
    params = list(named.values())
    opt = ClipAdam(params, lr=warmup_cosine_decay_schedule(**fx["schedule"]), clip=1.0)
    clipped = 0
    for step in range(fx["steps"]):
        for p in params:
            p.grad = None
        noise = torch.from_numpy(weight_recipe.normal(z[f"step{step}/noise_shape"]))
        loss, aux = loss_fn(m, fx["video"], expand_mask(fx["mask"], hw), fx["mask"], noise,
                            torch.from_numpy(z[f"step{step}/bernoulli_u"]))
        loss.backward()
        clipped += opt.step([p.grad if p.grad is not None else torch.zeros_like(p) for p in params]) >= 1.0
        print(f"train fixture step {step}:", check_train_step(z, step, loss.detach(), aux, named, 1e-5, 1e-5))
    assert clipped > 0 and (clipped < fx["steps"] or not fx["rl"])   # rl fixture: both branches of clip_by_global_norm taken
    idx = {n: i for i, n in enumerate(named)}
    print("train fixture end:", check_train_final(z, fx, named, lambda n: (opt.m[idx[n]].numpy(), opt.v[idx[n]].numpy()),
                                                  opt.count, 1e-4))
    if not fx["rl"]:
        from oracle.losses import eval_step
        loss, aux = eval_step(m, fx["video"], fx["mask"], fx["hparams"], hw, Rngs(7))
        print("train fixture eval_step:", check_eval_step(z, loss, aux, 1e-4))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after this round's GPU budget was spent: never executed on a GPU; non-strict so "
                                        "that its first run cannot stop the parity suite")
@pytest.mark.skipif(not TRAIN_FIXTURES, reason=NO_FIXTURE)
@pytest.mark.parametrize("path", TRAIN_FIXTURES or [None])
def test_cuda_path_reproduces_reference_training_trajectory(path):
    """The same six updates through rl_losses.train_step + ddp.FlatAdam (vvae_sumsq_f32_det + vvae_adam_step)."""
    import video_vae_b200 as V
    from video_vae_b200 import checkpoint as ck
    from video_vae_b200.ddp import FlatAdam, FlatParams
    from video_vae_b200.optim import warmup_cosine_decay_schedule
    z, fx = load_train_fixture(path)
    if fx["rl"]:
        from video_vae_b200.rl_losses import train_step as rl_train_step
        from video_vae_b200.rl_model import VideoVAE

        def train_step(m, video, mask, noise, u):
            return rl_train_step(m, video, mask, fx["hparams"], V.Rngs(0), cube_perceptual, None, noise=noise, bernoulli_u=u)
    else:
        from video_vae_b200.losses import loss_fn as vae_loss_fn
        from video_vae_b200.model import VideoVAE

        def train_step(m, video, mask, noise, u):
            loss, aux = vae_loss_fn(m, video, mask[:, None, None, :], mask, V.Rngs(0), fx["hparams"], noise=noise, gumbel_u=u)
            loss.backward()
            return loss, aux
    m = VideoVAE(*fx["cfg"], V.Rngs(0), dtype=torch.float32)
    shapes = {n: p.shape for n, p in m.named_parameters()}
    ck.load_flax_tree(m, {n: weight_recipe.param(n, shapes[n]) for n in fx["names"]}, strict=True)
    flat = FlatParams(m)
    adam = FlatAdam(flat, lr=warmup_cosine_decay_schedule(**fx["schedule"]), clip=1.0)
    named = dict(m.named_parameters())
    video, mask = fx["video"].cuda(), fx["mask"].cuda()
    for step in range(fx["steps"]):
        flat.zero_grad()
        noise = torch.from_numpy(weight_recipe.normal(z[f"step{step}/noise_shape"])).cuda()
        loss, aux = train_step(m, video, mask, noise, torch.from_numpy(z[f"step{step}/bernoulli_u"]).cuda())
        adam.step()
        torch.cuda.synchronize()
        print(f"train fixture (cuda) step {step}:", check_train_step(z, step, loss.detach(), aux, named, 1e-4, 1e-4))
    off = {id(p): o for p, o in zip(flat.params, flat.offsets)}

    def moments(n):
        p = named[n]
        o = off[id(p)]
        return adam.m[o:o + p.numel()].cpu().numpy(), adam.v[o:o + p.numel()].cpu().numpy()
    print("train fixture (cuda) end:", check_train_final(z, fx, named, moments, adam.t, 1e-3))
    if not fx["rl"]:
        from video_vae_b200.losses import eval_step
        loss, aux = eval_step(m, video, mask, fx["hparams"], V.Rngs(7))
        print("train fixture (cuda) eval_step:", check_eval_step(z, loss, aux, 1e-3))


# ------------------------------------------------------------------------------------------------ consumer self-check
def _write_stand_in(path):
    """A file with make_golden_jax.py's schema, produced by the ORACLE (so it pins nothing): exercises the loader, the
    weight import by Flax names and the comparison code that a real fixture will go through."""
    from oracle import Rngs
    from oracle.losses import DEFAULT_HPARAMS, expand_mask, loss_fn
    from oracle.model import VideoVAE
    from video_vae_b200 import checkpoint as ck
    cfg = (32, 32, 3, 16, 1, 1, 64, 2, 32, 16, 8, 4)
    o = VideoVAE(*cfg, Rngs(2))
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        o.decoder.unet.final_conv.kernel.copy_(torch.randn(o.decoder.unet.final_conv.kernel.shape, generator=g) * 0.05)
    b, t, hw = 2, 4, 4
    video = torch.rand(b, t, 32, 32, 3, generator=g)
    mask = torch.tensor([[True] * 4, [True] * 3 + [False]])
    noise = torch.randn(b, t, hw, 96, generator=g)
    u = torch.sigmoid((torch.tensor([[1., 0, 1, 1], [0, 1, 1, 0]]) * 2 - 1) * 6.0).reshape(b, t, 1)
    loss, aux = loss_fn(o, video, expand_mask(mask, hw), mask, Rngs(0), DEFAULT_HPARAMS, noise=noise, gumbel_u=u)
    loss.backward()
    out = {"cfg": np.asarray(cfg, np.int64), "dtype": np.asarray("float32"), "hparams": np.asarray(json.dumps(DEFAULT_HPARAMS)),
           "video": video.numpy(), "mask": mask.numpy(), "gumbel_u": u.numpy(), "noise": noise.numpy()}
    for k, v in ck.flatten_tree(ck.to_flax_tree(o)).items():
        out["param/" + k] = v
    for n, p in o.named_parameters():
        out["grad/" + n] = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    for k in ("MSE", "selection_loss", "kl_loss", "kept_frame_density", "reconstruction", "compressed", "selection", "logvar",
              "mean"):
        out["out/" + k] = aux[k].detach().numpy()
    out["out/loss"] = loss.detach().numpy()
    np.savez_compressed(path, **out)


def test_fixture_consumer_on_stand_in(tmp_path):
    path = str(tmp_path / "jax_videovae_standin_float32.npz")
    _write_stand_in(path)
    fx = load_fixture(path)
    loss, aux, m = run_impl(fx, "oracle", torch.float32)
    check_against_fixture(fx, loss, aux, m, 1e-5, 1e-4)
    fx["out"]["mean"] = fx["out"]["mean"] * 1.01              # and the check has teeth
    with pytest.raises(AssertionError):
        check_against_fixture(fx, loss, aux, m, 1e-5, 1e-4)


@pytest.mark.gpu
def test_fixture_consumer_on_stand_in_cuda(tmp_path):
    path = str(tmp_path / "jax_videovae_standin_float32.npz")
    _write_stand_in(path)
    fx = load_fixture(path)
    loss, aux, m = run_impl(fx, "cuda", torch.float32)
    # 32x32 clips: the U-Net's deepest maps are 4x4, its GroupNorm-preceded conv kernels see a few hundred voxels and
    # their (scale-invariant => cancelling) gradients carry fp32 summation-order noise of ~1e-3 relative L2 (r02b: 1.2e-3)
    check_against_fixture(fx, loss, aux, m, 1e-4, 3e-3)
