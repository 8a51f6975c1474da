#!/usr/bin/env python
"""Generate golden vectors from the REAL reference (JAX / Flax-NNX) -- run this on a box that has the reference's
environment (jax 0.9, flax 0.12, einops; claude_distributed/requirements.txt) and a checkout of floatingtrees/video-VAE:

    python tests/golden/make_golden_jax.py --reference /path/to/video-VAE [--cfg small|prod128] [--dtype float32]

Real JAX does not exist in the build container of this repository (no network).  There the same script runs with
`--shim`: the reference's files execute unmodified on oracle/jaxshim (jax / flax.nnx / jaxtyping / beartype look-alikes on
CPU torch, see oracle/jaxshim/README.md) and the file is named refshim_videovae_<cfg>_<dtype>.npz; that pins every line
the reference itself wrote, while the third-party primitives stay restated.  The file it writes,
tests/golden/{jax,refshim}_videovae_<cfg>_<dtype>.npz, is consumed by
tests/test_jax_golden.py: the CPU test loads the reference's weights into the oracle and compares the oracle with the
reference's outputs (this is what PINS the oracle); the GPU test does the same for the CUDA path.  Nothing else in the
repository reads the reference at run time.

What runs here is the reference's own code, unmodified: train/model.py::VideoVAE (imported from --reference/train) and
the functions `magnify_negatives` / `loss_fn` of train/legacy/training_loop_adversarial.py:66-124, extracted from that
file's AST and executed as they stand (importing the module itself would pull in its dataloader / wandb / orbax / VGG
dependencies, which the hot path does not need).  The random draws the model makes inside the step (the Gumbel gate's
uniform, train/layers.py:246, and the reparameterisation normal, train/model.py:126) are recorded by wrapping
jax.random.uniform / jax.random.normal for the duration of the call, and stored, so that the oracle and the CUDA path
can be fed the identical draws (`noise=`, `gumbel_u=`).

Stored keys: cfg (12 ints), dtype, hparams (json), video, mask, gumbel_u, noise, param/<dotted.name>, out/{loss, MSE,
selection_loss, kl_loss, kept_frame_density, reconstruction, compressed, selection, logvar, mean}, grad/<dotted.name>.
With --compact (implied by --shim): recipe, video_shape, pshape/<name> instead of video / param/ (the consumer rebuilds
them with tests/golden/weight_recipe.py) and gnorm/ gsum/ gmax/ gprobe/<name> instead of grad/.
Names are Flax attribute paths with list indices as integers -- exactly the names video_vae_b200 and the oracle use.
"""
import argparse
import ast
import json
import os
import sys

import numpy as np

CFGS = {
    # the reference's own CPU test configuration (claude_distributed/test_rl_model.py), a ~5 M parameter file
    "small": dict(cfg=(64, 64, 3, 16, 2, 2, 256, 4, 128, 32, 8, 4), batch=2, frames=6, keep=(6, 4)),
    # production hyper-parameters with 64-wide heads at a commit-sized width is impossible (170 M parameters): this
    # one is for local use only (680 MB of fp32 weights in the .npz)
    "prod128": dict(cfg=(128, 128, 3, 16, 9, 12, 1536, 8, 512, 64, 8, 4), batch=1, frames=16, keep=(12,)),
    # 64-wide heads (the tcgen05 attention / GEMM path of the CUDA build) at a small depth
    # BASELINE configs[4] shapes at depth 1: 512x512 frames (hw = 1024: the streaming L = 1024 spatial attention, the
    # full-resolution U-Net maps) and 64-frame clips (temporal attention at L = 64 with a ragged prefix mask)
    "hw1024": dict(cfg=(512, 512, 3, 16, 1, 1, 1536, 8, 512, 16, 8, 4), batch=1, frames=4, keep=(3,)),
    "t64": dict(cfg=(64, 64, 3, 16, 1, 1, 1536, 8, 512, 64, 8, 4), batch=2, frames=64, keep=(64, 41)),
    "hd64": dict(cfg=(64, 64, 3, 16, 2, 2, 256, 2, 128, 32, 8, 4), batch=2, frames=8, keep=(8, 5)),
}
HPARAMS = {"gamma1": 0.05, "gamma2": 0.001, "max_compression_rate": 2, "magnify_negatives_rate": 100}
# train/rl_nonadversarial.py:47-57,255-263
RL_HPARAMS = {"gamma1": 0.2, "gamma2": 0.001, "gamma3": 0.1, "gamma4": 0.05, "max_compression_rate": 2,
              "magnify_negatives_rate": 100, "rl_loss_weight": 0.01}


def _path_str(path):
    """jax key path -> 'a.b.0.c' (DictKey.key / GetAttrKey.name / SequenceKey.idx / FlattenedIndexKey.key)."""
    parts = []
    for k in path:
        for attr in ("key", "name", "idx"):
            if hasattr(k, attr):
                parts.append(str(getattr(k, attr)))
                break
        else:
            parts.append(str(k))
    if parts and parts[-1] in ("value", "raw_value"):
        parts = parts[:-1]
    return ".".join(parts)


def flatten_state(state):
    import jax
    if hasattr(state, "to_pure_dict"):
        state = state.to_pure_dict()
    leaves = jax.tree_util.tree_leaves_with_path(state)
    return {_path_str(p): np.asarray(v) for p, v in leaves}


def reference_functions(path, names, want):
    """The functions `names` exactly as written in the reference file `path` (extracted from its AST and executed
    unmodified; the rest of the file -- training-loop glue, wandb / orbax imports -- is not run).  Returns `want`."""
    import jax
    import jax.numpy as jnp
    from einops import rearrange, reduce, repeat
    from flax import nnx
    from jaxtyping import Array, Float
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in keep} == set(names), "reference layout changed"
    ns = {"jax": jax, "jnp": jnp, "nnx": nnx, "rearrange": rearrange, "reduce": reduce, "repeat": repeat,
          "Float": Float, "Array": Array}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)      # noqa: S102 (the reference's own code)
    return ns[want]


def reference_loss_fn(reference_root):
    """`magnify_negatives` and `loss_fn` exactly as written in train/legacy/training_loop_adversarial.py."""
    return reference_functions(os.path.join(reference_root, "train", "legacy", "training_loop_adversarial.py"),
                               ("magnify_negatives", "loss_fn"), "loss_fn")


def reference_rl_loss_fn(reference_root):
    """`per_sample_mean`, `magnify_negatives` and `loss_fn` exactly as written in train/rl_nonadversarial.py:59-186."""
    return reference_functions(os.path.join(reference_root, "train", "rl_nonadversarial.py"),
                               ("per_sample_mean", "magnify_negatives", "loss_fn"), "loss_fn")


class DrawRecorder:
    """Records every jax.random.uniform / normal result produced while active."""

    def __init__(self, normal_fn=None):
        self.uniform, self.normal, self.normal_fn = [], [], normal_fn

    def __enter__(self):
        import jax
        self._u, self._n, self._b = jax.random.uniform, jax.random.normal, jax.random.bernoulli

        def uniform(key, shape=(), *a, **k):
            out = self._u(key, shape, *a, **k)
            self.uniform.append(np.asarray(out))
            return out

        def bernoulli(key, p=0.5, shape=None):
            # jax.random.bernoulli IS `uniform(key, shape) < p` (jax/_src/random.py::_bernoulli); spelled out so that the
            # uniform draw behind it is recorded
            import jax.numpy as jnp
            p = jnp.asarray(p)
            return uniform(key, tuple(p.shape) if shape is None else tuple(shape)) < p

        def normal(key, shape=(), *a, **k):
            if self.normal_fn is not None:                       # --compact: the draw is a function of the shape
                import jax.numpy as jnp
                out = jnp.asarray(self.normal_fn(shape))
            else:
                out = self._n(key, shape, *a, **k)
            self.normal.append(np.asarray(out))
            return out
        jax.random.uniform, jax.random.normal, jax.random.bernoulli = uniform, normal, bernoulli
        return self

    def __exit__(self, *exc):
        import jax
        jax.random.uniform, jax.random.normal, jax.random.bernoulli = self._u, self._n, self._b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--cfg", default="small", choices=sorted(CFGS))
    ap.add_argument("--dtype", default="float32", choices=["float32", "bfloat16"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--model", default="vae", choices=["vae", "rl", "rl_dist"],
                    help="vae: train/model.py + the loss_fn of train/legacy/training_loop_adversarial.py; rl: "
                         "train/rl_model.py + the loss_fn of train/rl_nonadversarial.py (writes *_rlvae_*.npz); rl_dist: the "
                         "data-parallel trainer's copy claude_distributed/{rl_model,layers,unet}.py, which returns the "
                         "VARIANCE (rl_model.py:55-60,147), under a closed-form loss (writes *_rldistvae_*.npz)")
    ap.add_argument("--shim", action="store_true",
                    help="no JAX here: run the reference's files on oracle/jaxshim (jax / flax.nnx look-alikes on CPU torch) "
                         "and write refshim_videovae_<cfg>_<dtype>.npz")
    ap.add_argument("--compact", action="store_true",
                    help="weights from tests/golden/weight_recipe.py (by name), gradients as norm + sum + probe: a "
                         "commit-sized file at any width (implied by --shim)")
    args = ap.parse_args()
    args.compact = args.compact or args.shim
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import weight_recipe
    if args.shim:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                        "oracle", "jaxshim"))
    sys.path.insert(0, os.path.join(args.reference, "claude_distributed" if args.model == "rl_dist" else "train"))
    import jax
    import jax.numpy as jnp
    from einops import rearrange, repeat
    from flax import nnx
    if args.model == "vae":
        from model import VideoVAE                 # the reference's train/model.py
    else:
        from rl_model import VideoVAE              # the reference's train/rl_model.py (claude_distributed/ for rl_dist)
        import rl_model as _rm
        assert os.path.dirname(os.path.abspath(_rm.__file__)).endswith("claude_distributed" if args.model == "rl_dist" else "train")

    spec = CFGS[args.cfg]
    cfg, b, t = spec["cfg"], spec["batch"], spec["frames"]
    dtype = jnp.float32 if args.dtype == "float32" else jnp.bfloat16
    model = VideoVAE(*cfg, rngs=nnx.Rngs(2), dtype=dtype, param_dtype=jnp.float32)
    # the reference zero-initialises final_conv (train/unet.py:144-153), which switches the U-Net's gradients off:
    # randomise it so the fixture exercises the whole path
    k = model.decoder.unet.final_conv.kernel
    k.value = 0.05 * jax.random.normal(jax.random.key(5), k.value.shape, k.value.dtype)

    if args.compact:
        state = nnx.state(model, nnx.Param)
        names = flatten_state(state)
        pure = state.to_pure_dict()

        def fill(d, path):
            for kk, vv in d.items():
                if isinstance(vv, dict):
                    fill(vv, path + (str(kk),))
                else:
                    name = ".".join(path + (str(kk),))
                    assert name in names, name
                    d[kk] = jnp.asarray(weight_recipe.param(name, vv.shape))
        fill(pure, ())
        if hasattr(nnx, "replace_by_pure_dict"):           # real flax (untested here): State <- pure dict, then update
            nnx.replace_by_pure_dict(state, pure)
            nnx.update(model, state)
        else:
            nnx.update(model, pure)

    hw = (cfg[0] // cfg[3]) * (cfg[1] // cfg[3])
    key = jax.random.key(11)
    kv, = jax.random.split(key, 1)
    video = jax.random.uniform(kv, (b, t, cfg[0], cfg[1], cfg[2]), jnp.float32)
    if args.compact:
        video = jnp.asarray(weight_recipe.clip((b, t, cfg[0], cfg[1], cfg[2])))
    original_mask = jnp.arange(t)[None, :] < jnp.asarray(spec["keep"])[:, None]            # prefix masks (dataloader.py:232-234)
    mask = rearrange(original_mask, "b time -> b 1 1 time")                                # train_step, :126-130
    mask = repeat(mask, "b 1 1 time -> b hw 1 1 time", hw=hw)
    mask = rearrange(mask, "b hw 1 1 time -> (b hw) 1 1 time")
    if args.model == "rl_dist":       # claude_distributed/layers.py:213-214 expands the (b, 1, 1, t) mask itself
        mask = rearrange(original_mask, "b time -> b 1 1 time")

    rl = args.model in ("rl", "rl_dist")
    hparams = RL_HPARAMS if args.model == "rl" else HPARAMS
    if args.model == "rl_dist":
        # the trainer's loss is a closure inside distributed_train.py's main(); the model is what this fixture pins, under
        # a closed-form loss that reaches every output (tests/test_jax_golden.py::dist_loss is the same in torch)
        hparams = {}

        def loss_fn(model, video, mask, rngs):
            recon, compressed, selection, selection_mask, variance, mean = model(video, mask, rngs, train=True)
            loss = (jnp.mean(jnp.square(recon - repeat(video, "b ... -> (b 2) ..."))) + 0.01 * jnp.mean(variance)
                    + 0.01 * jnp.mean(jnp.square(mean)) + 0.1 * jnp.mean(selection))
            return loss, {"reconstruction": recon}
        grad_fn = nnx.value_and_grad(loss_fn, has_aux=True)
        with DrawRecorder(weight_recipe.normal if args.compact else None) as rec:
            (loss, aux), grads = grad_fn(model, video.astype(dtype), mask, nnx.Rngs(3))
        recon = aux["reconstruction"]
    elif rl:
        # the VGG term of rl_nonadversarial.py:125 is the caller's function (flaxmodels + downloaded weights in the
        # reference); a closed-form stand-in with the same signature keeps the gamma3 path and its gradient alive
        def perceptual(vgg_params, reconstruction, target):
            return jnp.mean(jnp.abs(reconstruction - target) ** 3, axis=(1, 2, 3, 4))
        loss_fn = reference_rl_loss_fn(args.reference)
        grad_fn = nnx.value_and_grad(loss_fn, has_aux=True)
        with DrawRecorder(weight_recipe.normal if args.compact else None) as rec:
            (loss, aux), grads = grad_fn(model, video.astype(dtype), mask, original_mask, nnx.Rngs(3), hparams, perceptual, None)
        recon = aux["reconstruction"]
    else:
        loss_fn = reference_loss_fn(args.reference)
        grad_fn = nnx.value_and_grad(loss_fn, has_aux=True)
        with DrawRecorder(weight_recipe.normal if args.compact else None) as rec:
            (loss, (mse, sel_loss, kl, recon, density)), grads = grad_fn(
                model, video.astype(dtype), mask, original_mask, nnx.Rngs(3), hparams)
    assert len(rec.uniform) == 1 and len(rec.normal) == 1, (len(rec.uniform), len(rec.normal))
    gumbel_u, noise = rec.uniform[0], rec.normal[0]              # rl: the uniform behind jax.random.bernoulli

    # second, non-differentiated call with the SAME draws to export the remaining outputs of VideoVAE.__call__
    class Replay:
        def __init__(self):
            self.u, self.n = [gumbel_u], [noise]

        def __enter__(self):
            self._u, self._n, self._b = jax.random.uniform, jax.random.normal, jax.random.bernoulli
            jax.random.uniform = lambda key, shape=(), *a, **k: jnp.asarray(self.u.pop(0))
            jax.random.normal = lambda key, shape=(), *a, **k: jnp.asarray(self.n.pop(0))
            jax.random.bernoulli = lambda key, p=0.5, shape=None: jnp.asarray(self.u.pop(0)) < p
            return self

        def __exit__(self, *exc):
            jax.random.uniform, jax.random.normal, jax.random.bernoulli = self._u, self._n, self._b
    with Replay():
        if rl:
            recon2, compressed, selection, selection_mask, logvar, mean = model(video.astype(dtype), mask, nnx.Rngs(3), train=True)
        else:
            recon2, compressed, selection, logvar, mean = model(video.astype(dtype), mask, nnx.Rngs(3), train=True)
    assert np.allclose(np.asarray(recon2, np.float32), np.asarray(recon, np.float32), rtol=1e-5, atol=1e-6)

    out = {"cfg": np.asarray(cfg, np.int64), "dtype": np.asarray(args.dtype), "hparams": np.asarray(json.dumps(hparams)),
           "video": np.asarray(video, np.float32), "mask": np.asarray(original_mask), "gumbel_u": gumbel_u.astype(np.float32),
           "noise": noise.astype(np.float32), "model": np.asarray(args.model)}
    if args.compact:
        out["recipe"] = np.asarray(weight_recipe.RECIPE_ID)
        del out["video"]                                         # weight_recipe.clip(shape)
        out["video_shape"] = np.asarray(video.shape, np.int64)
        assert np.array_equal(out["noise"], weight_recipe.normal(out["noise"].shape))
        out["noise_shape"] = np.asarray(out["noise"].shape, np.int64)
        del out["noise"]                                         # weight_recipe.normal(shape)
        for name, v in flatten_state(nnx.state(model, nnx.Param)).items():
            assert np.array_equal(v.astype(np.float32), weight_recipe.param(name, v.shape)), name
            out["pshape/" + name] = np.asarray(v.shape, np.int64)
        for name, v in flatten_state(grads).items():
            g = v.astype(np.float64)
            out["gnorm/" + name] = np.asarray(np.sqrt((g * g).sum()), np.float64)
            out["gsum/" + name] = np.asarray(g.sum(), np.float64)
            out["gmax/" + name] = np.asarray(np.abs(g).max(), np.float64)
            out["gprobe/" + name] = weight_recipe.grad_probe(v)
    else:
        for name, v in flatten_state(nnx.state(model, nnx.Param)).items():
            out["param/" + name] = v.astype(np.float32)
        for name, v in flatten_state(grads).items():
            out["grad/" + name] = v.astype(np.float32)
    f32 = lambda x: np.asarray(x, np.float32)      # noqa: E731
    if rl:
        if args.model == "rl":
            out.update({"out/" + k: f32(aux[k]) for k in ("MSE", "perceptual_loss", "selection_loss", "kl_loss",
                                                          "kept_frame_density", "mean_trajectory_prob", "rl_loss", "per_sample_MAE")})
        out.update({"out/loss": f32(loss), "out/reconstruction": f32(recon), "out/compressed": f32(compressed),
                    "out/selection": f32(selection), "out/selection_mask": f32(selection_mask), "out/logvar": f32(logvar),
                    "out/mean": f32(mean)})
    else:
        out.update({"out/loss": f32(loss), "out/MSE": f32(mse), "out/selection_loss": f32(sel_loss), "out/kl_loss": f32(kl),
                    "out/kept_frame_density": f32(density), "out/reconstruction": f32(recon), "out/compressed": f32(compressed),
                    "out/selection": f32(selection), "out/logvar": f32(logvar), "out/mean": f32(mean)})
    if args.compact:
        import math
        stride = max(1, math.ceil(math.sqrt(out["out/reconstruction"].size / 200_000)))  # keep the file commit-sized
        out["recon_stride"] = np.asarray(stride, np.int64)
        out["out/reconstruction"] = np.ascontiguousarray(out["out/reconstruction"][:, :, ::stride, ::stride, :])
        lat = 1 if out["out/mean"].size <= 200_000 else -(-out["out/mean"].size // 200_000)   # patches kept: every lat-th
        out["latent_stride"] = np.asarray(lat, np.int64)
        for k in ("out/mean", "out/logvar", "out/compressed"):
            out[k] = np.ascontiguousarray(out[k][:, :, ::lat, :])
    shim = bool(getattr(jax, "IS_SHIM", False))
    assert shim == args.shim, "a jax look-alike is on sys.path without --shim (or --shim found the real jax first)"
    out["generator"] = np.asarray(("reference files on oracle/jaxshim (CPU torch), jax " if shim else "reference files on jax ")
                                  + jax.__version__)
    prefix = "refshim" if shim else "jax"
    path = args.out or os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{prefix}_{ {'vae': 'videovae', 'rl': 'rlvae', 'rl_dist': 'rldistvae'}[args.model] }_{args.cfg}_{args.dtype}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss =", float(loss), "; jax", jax.__version__)


if __name__ == "__main__":
    main()
